#ifndef PINB_REFFULL_FFTW_H
#define PINB_REFFULL_FFTW_H
typedef double fftw_complex[2];
int fftw_init_threads(void);
void fftw_mpi_init(void);
void fftw_plan_with_nthreads(int);
void fftw_cleanup_threads(void);
void fftw_mpi_cleanup(void);
#endif
