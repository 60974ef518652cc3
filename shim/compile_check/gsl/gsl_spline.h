/* syntax-check stub: the public layout of gsl_spline (GSL 2.x interpolation/gsl_spline.h) */
#ifndef PINB_STUB_GSL_SPLINE_H
#define PINB_STUB_GSL_SPLINE_H
#include <stddef.h>
typedef struct gsl_interp_s gsl_interp;
typedef struct gsl_interp_accel_s gsl_interp_accel;
typedef struct { gsl_interp* interp; double* x; double* y; size_t size; } gsl_spline;
#endif
