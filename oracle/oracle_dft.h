/*
 * TEST INFRASTRUCTURE - not product code.
 *
 * The oracle's 1-D DFT: recursive mixed-radix decimation in time in IEEE double,
 * twiddle table generated in long double.  X[k] = sum_n x[n] exp(sign*2*pi*i*n*k/N),
 * unnormalised (FFTW's convention for fftw_plan_dft_1d, which is what the
 * reference calls at offt-compute.c:338, 416, 421, 443, 446).  Shared by the
 * FFTW stand-in (shim/shim_fftw.c) and the restatement (offt_oracle.c) so both
 * produce bit-identical rows.
 */
#ifndef OFFT_ORACLE_DFT_H
#define OFFT_ORACLE_DFT_H

#include <math.h>
#include <stdlib.h>

typedef struct { double re, im; } odft_cplx;

typedef struct {
  int n;
  int sign;
  int nfac;
  int fac[64];
  odft_cplx *tw; /* tw[j] = exp(sign*2*pi*i*j/n), j < n */
} odft_plan;

static void odft_factorize(odft_plan *p) {
  int n = p->n, nf = 0;
  while (n % 4 == 0) { p->fac[nf++] = 4; n /= 4; }
  while (n % 2 == 0) { p->fac[nf++] = 2; n /= 2; }
  int f;
  for (f = 3; f * f <= n; f += 2)
    while (n % f == 0) { p->fac[nf++] = f; n /= f; }
  if (n > 1) p->fac[nf++] = n;
  if (nf == 0) p->fac[nf++] = 1;
  p->nfac = nf;
}

/* out[0..n) = DFT of in[0], in[is], ...; twiddles taken from the size-N table with step ts = N/n */
static void odft_rec(const odft_plan *P, odft_cplx *out, const odft_cplx *in, int n, int is, int ts, int fi) {
  int p = P->fac[fi];
  int m = n / p;
  int q, k;
  if (m == 1) {
    for (q = 0; q < p; q++) out[q] = in[(size_t)q * is];
  } else {
    for (q = 0; q < p; q++) odft_rec(P, out + (size_t)q * m, in + (size_t)q * is, m, is * p, ts * p, fi + 1);
  }
  const odft_cplx *tw = P->tw;
  if (p == 2) {
    for (k = 0; k < m; k++) {
      odft_cplx w = tw[(size_t)k * ts];
      odft_cplx a = out[k], b = out[k + m];
      double br = b.re * w.re - b.im * w.im, bi = b.re * w.im + b.im * w.re;
      out[k].re = a.re + br; out[k].im = a.im + bi;
      out[k + m].re = a.re - br; out[k + m].im = a.im - bi;
    }
  } else if (p == 4) {
    /* exp(sign*i*pi/2) = sign*i */
    double sg = (double)P->sign;
    for (k = 0; k < m; k++) {
      odft_cplx w1 = tw[(size_t)k * ts], w2 = tw[(size_t)2 * k * ts], w3 = tw[(size_t)3 * k * ts];
      odft_cplx a = out[k], b = out[k + m], c = out[k + 2 * m], d = out[k + 3 * m];
      double br = b.re * w1.re - b.im * w1.im, bi = b.re * w1.im + b.im * w1.re;
      double cr = c.re * w2.re - c.im * w2.im, ci = c.re * w2.im + c.im * w2.re;
      double dr = d.re * w3.re - d.im * w3.im, di = d.re * w3.im + d.im * w3.re;
      double s0r = a.re + cr, s0i = a.im + ci, s1r = a.re - cr, s1i = a.im - ci;
      double s2r = br + dr, s2i = bi + di, s3r = br - dr, s3i = bi - di;
      /* (sign*i)*(s3) = sign*(-s3i, s3r) */
      double jr = -sg * s3i, ji = sg * s3r;
      out[k].re = s0r + s2r; out[k].im = s0i + s2i;
      out[k + m].re = s1r + jr; out[k + m].im = s1i + ji;
      out[k + 2 * m].re = s0r - s2r; out[k + 2 * m].im = s0i - s2i;
      out[k + 3 * m].re = s1r - jr; out[k + 3 * m].im = s1i - ji;
    }
  } else {
    odft_cplx t[p];
    int r;
    int N = P->n;
    for (k = 0; k < m; k++) {
      for (q = 0; q < p; q++) {
        odft_cplx w = tw[((size_t)q * k * ts) % N];
        odft_cplx v = out[k + (size_t)q * m];
        t[q].re = v.re * w.re - v.im * w.im;
        t[q].im = v.re * w.im + v.im * w.re;
      }
      for (r = 0; r < p; r++) {
        double sr = 0.0, si = 0.0;
        for (q = 0; q < p; q++) {
          /* w_p^{qr} = tw[(q*r mod p) * N/p] */
          odft_cplx w = tw[(size_t)((q * r) % p) * (N / p)];
          sr += t[q].re * w.re - t[q].im * w.im;
          si += t[q].re * w.im + t[q].im * w.re;
        }
        out[k + (size_t)r * m].re = sr; out[k + (size_t)r * m].im = si;
      }
    }
  }
}


static void odft_init(odft_plan *p, int n, int sign) {
  int j;
  p->n = n; p->sign = sign;
  odft_factorize(p);
  p->tw = (odft_cplx *)malloc(sizeof(odft_cplx) * (size_t)(n > 0 ? n : 1));
  for (j = 0; j < n; j++) {
    long double ang = (long double)sign * 2.0L * 3.14159265358979323846264338327950288L * (long double)j / (long double)n;
    p->tw[j].re = (double)cosl(ang);
    p->tw[j].im = (double)sinl(ang);
  }
}

static void odft_free(odft_plan *p) { free(p->tw); p->tw = NULL; }

/* out[0..n) (contiguous, must not alias in) = DFT of in[0], in[istride], ... */
static void odft_exec(const odft_plan *p, odft_cplx *out, const odft_cplx *in, int istride) {
  odft_rec(p, out, in, p->n, istride, 1, 0);
}

#endif
