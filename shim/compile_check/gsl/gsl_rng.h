/* syntax-check stub */
#ifndef PINB_STUB_GSL_RNG_H
#define PINB_STUB_GSL_RNG_H
typedef struct gsl_rng_s gsl_rng;
#endif
