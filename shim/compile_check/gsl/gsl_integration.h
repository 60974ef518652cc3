/* syntax-check stub */
#ifndef PINB_STUB_GSL_INT_H
#define PINB_STUB_GSL_INT_H
typedef struct gsl_integration_workspace_s gsl_integration_workspace;
#endif
