#ifndef PINB_REFFULL_GSL_INTERP2D_H
#define PINB_REFFULL_GSL_INTERP2D_H
#include <stddef.h>
typedef struct gsl_interp2d_type_s gsl_interp2d_type;
typedef struct gsl_interp2d_s gsl_interp2d;
extern const gsl_interp2d_type* gsl_interp2d_bicubic;
#endif
