/* syntax-check stub, see mpi.h */
#ifndef PINB_STUB_PFFT_H
#define PINB_STUB_PFFT_H
#include <stddef.h>
typedef double pfft_complex[2];
typedef struct pfft_plan_s* pfft_plan;
#endif
