#ifndef PINB_REFFULL_GSL_ERRNO_H
#define PINB_REFFULL_GSL_ERRNO_H
#define GSL_SUCCESS 0
#define GSL_CONTINUE (-2)
#define GSL_FAILURE (-1)
#define GSL_EMAXITER 11
typedef void gsl_error_handler_t(const char* reason, const char* file, int line, int gsl_errno);
gsl_error_handler_t* gsl_set_error_handler_off(void);
const char* gsl_strerror(int);
#endif
