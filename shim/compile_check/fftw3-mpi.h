/* syntax-check stub, see mpi.h */
#ifndef PINB_STUB_FFTW_H
#define PINB_STUB_FFTW_H
typedef double fftw_complex[2];
#endif
