/* syntax-check stub */
#ifndef PINB_STUB_GSL_RNG_H
#define PINB_STUB_GSL_RNG_H
typedef struct gsl_rng_s gsl_rng;
void gsl_rng_set(const gsl_rng* r, unsigned long int seed);
double gsl_rng_uniform(const gsl_rng* r);
#endif
