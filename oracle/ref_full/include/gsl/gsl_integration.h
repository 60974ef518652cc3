#ifndef PINB_REFFULL_GSL_INT_H
#define PINB_REFFULL_GSL_INT_H
#include <stddef.h>
#include "gsl_math.h"
typedef struct gsl_integration_workspace_s gsl_integration_workspace;
gsl_integration_workspace* gsl_integration_workspace_alloc(size_t n);
void gsl_integration_workspace_free(gsl_integration_workspace*);
int gsl_integration_qags(const gsl_function* f, double a, double b, double epsabs, double epsrel, size_t limit,
                         gsl_integration_workspace* w, double* result, double* abserr);
#endif
