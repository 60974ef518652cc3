/* One-task PFFT stand-in (see oracle/ref_stubs/pfft.h; implemented by oracle/ref_fft.c) plus the
 * process-mesh calls of src/initialization.c:176-203. */
#ifndef PINB_REFFULL_PFFT_H
#define PINB_REFFULL_PFFT_H
#include "../../ref_stubs/pfft.h"
int pfft_create_procmesh(int rnk, MPI_Comm comm, const int* np, MPI_Comm* comm_cart);
void pfft_plan_with_nthreads(int n);
#define pfft_fprintf(comm, stream, ...) fprintf(stream, __VA_ARGS__)
#define pfft_printf(comm, ...) printf(__VA_ARGS__)
#endif
