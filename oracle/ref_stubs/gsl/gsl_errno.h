/* Declaration-only stand-in for GSL (absent from this image), used to compile the reference's
 * collapse_times.c verbatim for oracle/_ref.  Only what that translation unit names. */
#ifndef PINB_REFSTUB_GSL_ERRNO_H
#define PINB_REFSTUB_GSL_ERRNO_H
#define GSL_SUCCESS 0
#endif
