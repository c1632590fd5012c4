/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Stand-in for the FFTW3 subset the reference calls:
 * FFTW 3.3.x itself is a third-party dependency that is neither vendored under
 * /root/reference nor installed here.  The arithmetic behind these entry
 * points is the oracle's own double-precision DFT (shim_fftw.c), which follows
 * FFTW's published conventions: forward transform, exponent sign -1,
 * unnormalised.
 */
#ifndef OFFT_ORACLE_SHIM_FFTW3_H
#define OFFT_ORACLE_SHIM_FFTW3_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef double fftw_complex[2];
typedef struct shim_fftw_plan_s *fftw_plan;
typedef struct { int n; int is; int os; } fftw_iodim;

#define FFTW_FORWARD (-1)
#define FFTW_BACKWARD (+1)
#define FFTW_MEASURE (0U)
#define FFTW_EXHAUSTIVE (1U << 3)
#define FFTW_PATIENT (1U << 5)
#define FFTW_ESTIMATE (1U << 6)

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags);
fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed, int istride, int idist,
                             fftw_complex *out, const int *onembed, int ostride, int odist,
                             int sign, unsigned flags);
fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags);
fftw_plan fftw_plan_guru_dft(int rank, const fftw_iodim *dims, int howmany_rank,
                             const fftw_iodim *howmany_dims, fftw_complex *in, fftw_complex *out,
                             int sign, unsigned flags);
void fftw_execute(const fftw_plan p);
void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out);
void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out);
void fftw_destroy_plan(fftw_plan p);
void fftw_print_plan(const fftw_plan p);
void *fftw_malloc(size_t n);
void fftw_free(void *p);

#ifdef __cplusplus
}
#endif
#endif
