#ifndef PINB_REFFULL_GSL_SPLINE2D_H
#define PINB_REFFULL_GSL_SPLINE2D_H
#include "gsl_interp2d.h"
#include "gsl_spline.h"
typedef struct gsl_spline2d_s gsl_spline2d;
gsl_spline2d* gsl_spline2d_alloc(const gsl_interp2d_type*, size_t nx, size_t ny);
int gsl_spline2d_init(gsl_spline2d*, const double* x, const double* y, const double* z, size_t nx, size_t ny);
double gsl_spline2d_eval(const gsl_spline2d*, double x, double y, gsl_interp_accel*, gsl_interp_accel*);
void gsl_spline2d_free(gsl_spline2d*);
#endif
