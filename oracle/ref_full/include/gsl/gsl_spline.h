#ifndef PINB_REFFULL_GSL_SPLINE_H
#define PINB_REFFULL_GSL_SPLINE_H
#include <stddef.h>
typedef struct gsl_interp_type_s gsl_interp_type;
typedef struct gsl_interp_s gsl_interp;
typedef struct gsl_interp_accel_s gsl_interp_accel;
/* public layout of GSL 2.x interpolation/gsl_spline.h */
typedef struct { gsl_interp* interp; double* x; double* y; size_t size; } gsl_spline;
extern const gsl_interp_type* gsl_interp_linear;
extern const gsl_interp_type* gsl_interp_cspline;
gsl_interp_accel* gsl_interp_accel_alloc(void);
void gsl_interp_accel_free(gsl_interp_accel*);
int gsl_interp_accel_reset(gsl_interp_accel*);
gsl_interp* gsl_interp_alloc(const gsl_interp_type*, size_t n);
int gsl_interp_init(gsl_interp*, const double* x, const double* y, size_t n);
double gsl_interp_eval(const gsl_interp*, const double* x, const double* y, double xq, gsl_interp_accel*);
void gsl_interp_free(gsl_interp*);
gsl_spline* gsl_spline_alloc(const gsl_interp_type*, size_t n);
int gsl_spline_init(gsl_spline*, const double* x, const double* y, size_t n);
double gsl_spline_eval(const gsl_spline*, double xq, gsl_interp_accel*);
double gsl_spline_eval_deriv(const gsl_spline*, double xq, gsl_interp_accel*);
void gsl_spline_free(gsl_spline*);
#endif
