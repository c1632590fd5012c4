/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Stand-in for the FFTW3 entry points the reference calls (see fftw3.h in this
 * directory).  FFTW itself is absent from /root/reference and from this image,
 * so the arithmetic here is the oracle's own: a recursive mixed-radix
 * decimation-in-time DFT in IEEE double with a twiddle table generated in long
 * double.  Conventions follow FFTW's manual: X[k] = sum_n x[n] exp(sign*2*pi*i*n*k/N),
 * sign = -1 for FFTW_FORWARD, no normalisation; rank-0 guru plans are pure
 * strided copies (used by the reference as in-place 3-D permutations,
 * offt-compute.c:637).
 */
#include "fftw3.h"
#include "fftw3-mpi.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../oracle_dft.h"
typedef odft_cplx cplx;

enum { PLAN_C2C = 1, PLAN_R2C = 2, PLAN_COPY = 3 };

struct shim_fftw_plan_s {
  int kind;
  int n;
  int sign;
  int howmany, istride, idist, ostride, odist;
  odft_plan dft;
  cplx *a, *b;    /* contiguous work rows */
  fftw_iodim hd[3];
  void *plan_in, *plan_out;
};

static struct shim_fftw_plan_s *plan_new(int kind, int n, int sign) {
  struct shim_fftw_plan_s *p = (struct shim_fftw_plan_s *)calloc(1, sizeof(*p));
  p->kind = kind; p->n = n; p->sign = sign;
  p->howmany = 1; p->istride = p->ostride = 1; p->idist = p->odist = 0;
  if (kind != PLAN_COPY) {
    odft_init(&p->dft, n, sign);
    p->a = (cplx *)malloc(sizeof(cplx) * (size_t)n);
    p->b = (cplx *)malloc(sizeof(cplx) * (size_t)n);
  }
  return p;
}

fftw_plan fftw_plan_dft_1d(int n, fftw_complex *in, fftw_complex *out, int sign, unsigned flags) {
  (void)flags;
  struct shim_fftw_plan_s *p = plan_new(PLAN_C2C, n, sign);
  p->plan_in = in; p->plan_out = out;
  return p;
}

fftw_plan fftw_plan_many_dft(int rank, const int *n, int howmany,
                             fftw_complex *in, const int *inembed, int istride, int idist,
                             fftw_complex *out, const int *onembed, int ostride, int odist,
                             int sign, unsigned flags) {
  (void)inembed; (void)onembed; (void)flags;
  if (rank != 1) { fprintf(stderr, "shim-fftw: only rank-1 plan_many_dft\n"); abort(); }
  struct shim_fftw_plan_s *p = plan_new(PLAN_C2C, n[0], sign);
  p->howmany = howmany; p->istride = istride; p->idist = idist; p->ostride = ostride; p->odist = odist;
  p->plan_in = in; p->plan_out = out;
  return p;
}

fftw_plan fftw_plan_dft_r2c_1d(int n, double *in, fftw_complex *out, unsigned flags) {
  (void)flags;
  struct shim_fftw_plan_s *p = plan_new(PLAN_R2C, n, FFTW_FORWARD);
  p->plan_in = in; p->plan_out = out;
  return p;
}

fftw_plan fftw_plan_guru_dft(int rank, const fftw_iodim *dims, int howmany_rank,
                             const fftw_iodim *howmany_dims, fftw_complex *in, fftw_complex *out,
                             int sign, unsigned flags) {
  (void)dims; (void)sign; (void)flags;
  if (rank != 0 || howmany_rank != 3) { fprintf(stderr, "shim-fftw: only rank-0 guru copies with 3 loops\n"); abort(); }
  struct shim_fftw_plan_s *p = plan_new(PLAN_COPY, 0, 0);
  memcpy(p->hd, howmany_dims, 3 * sizeof(fftw_iodim));
  p->plan_in = in; p->plan_out = out;
  return p;
}

static void exec_c2c(const struct shim_fftw_plan_s *p, cplx *in, cplx *out) {
  int h, j, n = p->n;
  for (h = 0; h < p->howmany; h++) {
    const cplx *src = in + (size_t)h * p->idist;
    cplx *dst = out + (size_t)h * p->odist;
    if (p->ostride == 1 && (const cplx *)dst != src) {
      odft_exec(&p->dft, dst, src, p->istride);
    } else {
      odft_exec(&p->dft, p->b, src, p->istride);
      if (p->ostride == 1) memcpy(dst, p->b, sizeof(cplx) * (size_t)n);
      else for (j = 0; j < n; j++) dst[(size_t)j * p->ostride] = p->b[j];
    }
  }
}

static void exec_copy(const struct shim_fftw_plan_s *p, cplx *in, cplx *out) {
  size_t ext = 1;
  int d;
  for (d = 0; d < 3; d++) ext += (size_t)(p->hd[d].n - 1) * (size_t)p->hd[d].is;
  const cplx *src = in;
  cplx *tmp = NULL;
  if (in == out) {
    tmp = (cplx *)malloc(sizeof(cplx) * ext);
    memcpy(tmp, in, sizeof(cplx) * ext);
    src = tmp;
  }
  int i0, i1, i2;
  for (i0 = 0; i0 < p->hd[0].n; i0++)
    for (i1 = 0; i1 < p->hd[1].n; i1++)
      for (i2 = 0; i2 < p->hd[2].n; i2++)
        out[(size_t)i0 * p->hd[0].os + (size_t)i1 * p->hd[1].os + (size_t)i2 * p->hd[2].os] =
            src[(size_t)i0 * p->hd[0].is + (size_t)i1 * p->hd[1].is + (size_t)i2 * p->hd[2].is];
  free(tmp);
}

void fftw_execute_dft(const fftw_plan p, fftw_complex *in, fftw_complex *out) {
  if (p->kind == PLAN_C2C) exec_c2c(p, (cplx *)in, (cplx *)out);
  else if (p->kind == PLAN_COPY) exec_copy(p, (cplx *)in, (cplx *)out);
  else { fprintf(stderr, "shim-fftw: execute_dft on an r2c plan\n"); abort(); }
}

void fftw_execute_dft_r2c(const fftw_plan p, double *in, fftw_complex *out) {
  int j, n = p->n;
  cplx *o = (cplx *)out;
  for (j = 0; j < n; j++) { p->a[j].re = in[j]; p->a[j].im = 0.0; }
  odft_exec(&p->dft, p->b, p->a, 1);
  for (j = 0; j <= n / 2; j++) o[j] = p->b[j];
}

void fftw_execute(const fftw_plan p) {
  if (p->kind == PLAN_R2C) fftw_execute_dft_r2c(p, (double *)p->plan_in, (fftw_complex *)p->plan_out);
  else fftw_execute_dft(p, (fftw_complex *)p->plan_in, (fftw_complex *)p->plan_out);
}

void fftw_destroy_plan(fftw_plan p) {
  if (!p) return;
  if (p->kind != PLAN_COPY) odft_free(&p->dft);
  free(p->a); free(p->b); free(p);
}

void fftw_print_plan(const fftw_plan p) {
  printf("(shim-dft kind %d n %d)", p->kind, p->n);
}

void *fftw_malloc(size_t n) { return malloc(n); }
void fftw_free(void *p) { free(p); }

/* ---- FFTW-MPI comparator path of run-fft.c (-a 1): out of scope, link stubs only ---- */
void fftw_mpi_init(void) {}
void fftw_mpi_cleanup(void) {}
fftw_plan fftw_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in,
                               fftw_complex *out, MPI_Comm comm, int sign, unsigned flags) {
  (void)n0; (void)n1; (void)n2; (void)in; (void)out; (void)comm; (void)sign; (void)flags;
  fprintf(stderr, "shim-fftw: the FFTW-MPI comparator path (-a 1) is not provided\n");
  abort();
  return NULL;
}
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in,
                                   fftw_complex *out, MPI_Comm comm, unsigned flags) {
  (void)n0; (void)n1; (void)n2; (void)in; (void)out; (void)comm; (void)flags;
  fprintf(stderr, "shim-fftw: the FFTW-MPI comparator path (-a 1) is not provided\n");
  abort();
  return NULL;
}
