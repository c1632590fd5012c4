/*
 * Stand-in for <fftw3.h> on boxes without FFTW: only what the reference driver
 * (run-fft.c) and offt.h need to compile - type names and planner-flag macros.
 * The B200 library never calls FFTW; where a real fftw3.h is installed, drop
 * include/compat from the include path.
 */
#ifndef OFFTB_COMPAT_FFTW3_H
#define OFFTB_COMPAT_FFTW3_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
typedef double fftw_complex[2];
typedef struct offtb_fftw_plan_s *fftw_plan;
#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)
void fftw_execute(const fftw_plan p);
void fftw_destroy_plan(fftw_plan p);
void fftw_print_plan(const fftw_plan p);
#ifdef __cplusplus
}
#endif
#endif
