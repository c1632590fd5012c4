// Real-to-complex z pass on the register-butterfly kernels (local z passes of is_r2c plans, offt-compute.c:960-961,
// 3973-3974, where FFTW runs its r2c plan).
//
// A row of N reals is read as H = N/2 complex numbers z[j] = x[2j] + i x[2j+1] - the in-place r2c layout already puts
// them there (run-fft.c:53-55) - and transformed with the ordinary H-point kernel; this file holds the O(N) step that
// turns that spectrum into the first H+1 points of the N-point one, and its inverse:
//     Xe[k] = (Z[k] + conj Z[H-k]) / 2          even samples' spectrum
//     Xo[k] = (Z[k] - conj Z[H-k]) / (2i)       odd samples' spectrum
//     X[k]  = Xe[k] + w^k Xo[k],   w = exp(-2 pi i / N),   k = 0..H   (Z[H] = Z[0])
// backward (unnormalised, so that forward then backward returns N x):
//     Z'[k] = (X[k] + conj X[H-k]) + i conj(w^k) (X[k] - conj X[H-k]),   k < H,   then the H-point backward transform.
// One thread owns the pair (k, H-k) of one row: it reads both, writes both, so the step is safe in place; rows are
// addressed through the batch digits of an FftMap (no split: local z rows are whole).
#include "fft_launch.h"

#include <algorithm>

namespace offtb {

namespace {

template <typename T>
__global__ void r2c_post_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, const cx<T> *__restrict__ w, FftMap im, FftMap om,
                                int H, long long nbatch) {
  const int pairs = H / 2 + 1;
  const long long total = nbatch * pairs;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / pairs;
    const int k = (int)(e - b * pairs);
    const cx<T> *src = in + map_b(im, (unsigned)b);
    cx<T> *dst = out + map_b(om, (unsigned)b);
    if (k == 0) {
      const cx<T> z0 = src[0];
      dst[0] = cx<T>{z0.x + z0.y, (T)0};
      dst[H] = cx<T>{z0.x - z0.y, (T)0};
      continue;
    }
    const cx<T> a = src[k], bz = src[H - k];
    // Xe = (a + conj b)/2, Xo = (a - conj b)/(2i)
    const cx<T> xe = {(a.x + bz.x) * (T)0.5, (a.y - bz.y) * (T)0.5};
    const cx<T> xo = {(a.y + bz.y) * (T)0.5, (bz.x - a.x) * (T)0.5};
    const cx<T> wk = w[k];
    const cx<T> t = cmul(xo, wk);
    const cx<T> xk = cadd(xe, t);                        // X[k]
    // X[H-k] = conj(Xe[k]) + w^(H-k) conj(Xo[k]) = conj(Xe[k] - w^k Xo[k])
    const cx<T> d = csub(xe, t);
    dst[k] = xk;
    dst[H - k] = cx<T>{d.x, -d.y};
  }
}

template <typename T>
__global__ void c2r_pre_kernel(const cx<T> *__restrict__ in, cx<T> *__restrict__ out, const cx<T> *__restrict__ w, FftMap im, FftMap om,
                               int H, long long nbatch) {
  const int pairs = H / 2 + 1;
  const long long total = nbatch * pairs;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const long long b = e / pairs;
    const int k = (int)(e - b * pairs);
    const cx<T> *src = in + map_b(im, (unsigned)b);
    cx<T> *dst = out + map_b(om, (unsigned)b);
    if (k == 0) {
      const cx<T> x0 = src[0], xh = src[H];
      // Z'[0] = (X0 + conj XH) + i (X0 - conj XH)
      const cx<T> s = {x0.x + xh.x, x0.y - xh.y}, d = {x0.x - xh.x, x0.y + xh.y};
      dst[0] = cx<T>{s.x - d.y, s.y + d.x};
      continue;
    }
    const cx<T> a = src[k], bx = src[H - k];
    const cx<T> wk = w[k];
    const cx<T> wc = {wk.x, -wk.y};
    // k:   Z'[k]   = (a + conj b) + i conj(w^k) (a - conj b)
    const cx<T> s = {a.x + bx.x, a.y - bx.y}, d = {a.x - bx.x, a.y + bx.y};
    const cx<T> t = cmul(d, wc);
    // H-k: Z'[H-k] = (b + conj a) + i conj(w^(H-k)) (b - conj a) = conj(s) + i (-w^k) (-conj d) = conj(s) + i w^k conj(d)
    const cx<T> dc = {d.x, -d.y};
    const cx<T> u = cmul(dc, wk);
    dst[k] = cx<T>{s.x - t.y, s.y + t.x};
    if (2 * k != H) dst[H - k] = cx<T>{s.x - u.y, -s.y + u.x};
  }
}

}  // namespace

// w: exp(-2 pi i k / N), k <= H/2, on the device.  forward: post-step of the r2c pass (rows of H complex -> H+1 complex);
// backward: pre-step of the c2r pass (rows of H+1 complex -> H complex)
cudaError_t r2c_step_launch(int prec, bool backward, const void *in, void *out, const void *w, const FftMap &im, const FftMap &om, int H,
                            long long nbatch, cudaStream_t stream) {
  if (nbatch <= 0) return cudaSuccess;
  const long long total = nbatch * (H / 2 + 1);
  const int threads = 256;
  const unsigned grid = (unsigned)std::min<long long>((total + threads - 1) / threads, 148LL * 16);
  if (prec == PREC_F64) {
    if (backward) c2r_pre_kernel<double><<<grid, threads, 0, stream>>>((const cx<double> *)in, (cx<double> *)out, (const cx<double> *)w, im, om, H, nbatch);
    else r2c_post_kernel<double><<<grid, threads, 0, stream>>>((const cx<double> *)in, (cx<double> *)out, (const cx<double> *)w, im, om, H, nbatch);
  } else {
    if (backward) c2r_pre_kernel<float><<<grid, threads, 0, stream>>>((const cx<float> *)in, (cx<float> *)out, (const cx<float> *)w, im, om, H, nbatch);
    else r2c_post_kernel<float><<<grid, threads, 0, stream>>>((const cx<float> *)in, (cx<float> *)out, (const cx<float> *)w, im, om, H, nbatch);
  }
  return cudaGetLastError();
}

}  // namespace offtb
