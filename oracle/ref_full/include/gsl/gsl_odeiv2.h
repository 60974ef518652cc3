/* Declaration-only stand-in, see gsl_errno.h.  ell_sng() (src/collapse_times.c:315-400) is
 * compiled but never called with -DELL_CLASSIC; oracle/ref_harness.c defines these as traps. */
#ifndef PINB_REFSTUB_GSL_ODEIV2_H
#define PINB_REFSTUB_GSL_ODEIV2_H
#include <stddef.h>
#include "gsl_errno.h"
typedef struct gsl_odeiv2_step_type_s gsl_odeiv2_step_type;
typedef struct gsl_odeiv2_step_s gsl_odeiv2_step;
typedef struct gsl_odeiv2_control_s gsl_odeiv2_control;
typedef struct gsl_odeiv2_evolve_s gsl_odeiv2_evolve;
typedef struct {
  int (*function)(double t, const double y[], double dydt[], void* params);
  int (*jacobian)(double t, const double y[], double* dfdy, double dfdt[], void* params);
  size_t dimension;
  void* params;
} gsl_odeiv2_system;
extern const gsl_odeiv2_step_type* gsl_odeiv2_step_rkf45;
gsl_odeiv2_step* gsl_odeiv2_step_alloc(const gsl_odeiv2_step_type*, size_t);
gsl_odeiv2_control* gsl_odeiv2_control_standard_new(double, double, double, double);
gsl_odeiv2_evolve* gsl_odeiv2_evolve_alloc(size_t);
int gsl_odeiv2_evolve_apply(gsl_odeiv2_evolve*, gsl_odeiv2_control*, gsl_odeiv2_step*, const gsl_odeiv2_system*, double* t,
                            double t1, double* h, double y[]);
void gsl_odeiv2_evolve_free(gsl_odeiv2_evolve*);
void gsl_odeiv2_control_free(gsl_odeiv2_control*);
void gsl_odeiv2_step_free(gsl_odeiv2_step*);
#endif
