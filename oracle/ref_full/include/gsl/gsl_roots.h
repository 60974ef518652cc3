#ifndef PINB_REFFULL_GSL_ROOTS_H
#define PINB_REFFULL_GSL_ROOTS_H
#include "gsl_math.h"
typedef struct gsl_root_fsolver_type_s gsl_root_fsolver_type;
typedef struct gsl_root_fsolver_s gsl_root_fsolver;
extern const gsl_root_fsolver_type* gsl_root_fsolver_brent;
gsl_root_fsolver* gsl_root_fsolver_alloc(const gsl_root_fsolver_type*);
void gsl_root_fsolver_free(gsl_root_fsolver*);
int gsl_root_fsolver_set(gsl_root_fsolver*, gsl_function*, double x_lower, double x_upper);
int gsl_root_fsolver_iterate(gsl_root_fsolver*);
double gsl_root_fsolver_root(const gsl_root_fsolver*);
double gsl_root_fsolver_x_lower(const gsl_root_fsolver*);
double gsl_root_fsolver_x_upper(const gsl_root_fsolver*);
int gsl_root_test_interval(double x_lower, double x_upper, double epsabs, double epsrel);
#endif
