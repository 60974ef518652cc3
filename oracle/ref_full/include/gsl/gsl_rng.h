/* Declarations of the part of GSL's rng interface that src/GenIC.c uses (GSL is absent from this
 * image); oracle/ref_gsl_rng.c restates the two generators.  TEST INFRASTRUCTURE. */
#ifndef PINB_REFSTUB_GSL_RNG_H
#define PINB_REFSTUB_GSL_RNG_H
typedef struct gsl_rng_type_s gsl_rng_type;
typedef struct gsl_rng_s gsl_rng;
extern const gsl_rng_type* gsl_rng_ranlxd1;
extern const gsl_rng_type* gsl_rng_mt19937;
gsl_rng* gsl_rng_alloc(const gsl_rng_type* T);
void gsl_rng_set(const gsl_rng* r, unsigned long int seed);
double gsl_rng_uniform(const gsl_rng* r);
unsigned long int gsl_rng_get(const gsl_rng* r);
void gsl_rng_free(gsl_rng* r);
#endif
