#ifndef PINB_REFFULL_GSL_MATH_H
#define PINB_REFFULL_GSL_MATH_H
#include <math.h>
typedef struct { double (*function)(double x, void* params); void* params; } gsl_function;
#define GSL_FN_EVAL(F, x) (*((F)->function))(x, (F)->params)
#endif
