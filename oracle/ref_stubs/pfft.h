/* Declaration-only stand-in for PFFT 1.0.8 (absent from this image): the handful of entry
 * points src/fmax-pfft.c and src/allocations.c name.  oracle/ref_fft.c implements them for ONE
 * task with an in-repo FFT (slab = whole box, non-transposed layout). */
#ifndef PINB_REFSTUB_PFFT_H
#define PINB_REFSTUB_PFFT_H
#include <stddef.h>
#include <mpi.h>
typedef double pfft_complex[2];
typedef struct pfft_plan_s* pfft_plan;
#define PFFT_FORWARD (-1)
#define PFFT_BACKWARD (+1)
#define PFFT_MEASURE (0U)
#define PFFT_TUNE (1U << 5)
#define PFFT_TRANSPOSED_IN (1U << 8)
#define PFFT_TRANSPOSED_OUT (1U << 9)
void pfft_init(void);
void pfft_cleanup(void);
ptrdiff_t pfft_local_size_dft_r2c_3d(const ptrdiff_t* n, MPI_Comm comm, unsigned flags, ptrdiff_t* local_ni, ptrdiff_t* local_i_start,
                                     ptrdiff_t* local_no, ptrdiff_t* local_o_start);
pfft_plan pfft_plan_dft_r2c_3d(const ptrdiff_t* n, double* in, pfft_complex* out, MPI_Comm comm, int sign, unsigned flags);
pfft_plan pfft_plan_dft_c2r_3d(const ptrdiff_t* n, pfft_complex* in, double* out, MPI_Comm comm, int sign, unsigned flags);
void pfft_execute(const pfft_plan plan);
void pfft_destroy_plan(pfft_plan plan);
double* pfft_alloc_real(size_t n);
pfft_complex* pfft_alloc_complex(size_t n);
void pfft_free(void* p);
#endif
