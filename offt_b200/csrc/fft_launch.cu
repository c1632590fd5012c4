// Length dispatch for the batched 1-D FFT kernels.
#include "fft_configs.h"
#include "fft_launch.h"

#include <algorithm>
#include <cmath>

namespace offtb {

#define X(N, ...)                                                                                  \
  cudaError_t fft_launch_##N(int prec, const FftArgs &args, long long nbatch, cudaStream_t stream, FftShape *shape_only); \
  void fft_info_##N(int prec, FftKernelInfo *info);
OFFTB_FFT_CONFIGS(X)
#undef X

bool fft_kernel_info(int N, int prec, FftKernelInfo *info) {
  switch (N) {
#define X(N, ...) case N: fft_info_##N(prec, info); return true;
    OFFTB_FFT_CONFIGS(X)
#undef X
    default: return false;
  }
}

int fft_twiddle_table(int N, int prec, long double *out) {
  FftKernelInfo info;
  if (!fft_kernel_info(N, prec, &info)) return -1;
  const long double two_pi = 6.283185307179586476925286766559005768L;
  int count = 0;
  long long P = 1;
  for (int s = 0; s + 1 < info.ns; ++s) {
    const long long sub = N / P;              // R_s * M_s: length of the sub-problems of this stage
    const long long M = sub / info.radix[s];
    for (long long np = 0; np < M; ++np) {
      const long double ang = two_pi * (long double)np / (long double)sub;
      out[2 * count] = cosl(ang);
      out[2 * count + 1] = -sinl(ang);
      ++count;
    }
    P *= info.radix[s];
  }
  return count;
}

// Columns per CTA (measured on B200, tools/kbench.py; DESIGN.md has the table).
// Contiguous rows ("n-fast"): tiles of about 16 KB and at least two warps - many small CTAs per SM interleave
// their phases best (512-point complex128 rows: 6.3 TB/s at 2 columns, 5.8 at 4; 1024-point: 6.5 TB/s at 1).
// Strided axes ("c-fast"): 64 contiguous bytes per transform index, inside one run of the lowest batch digit;
// 128 bytes when consecutive transform indices are a megabyte or more apart, where every row is its own page
// and the cost is per row touched (512^3 x pass: 4.6 TB/s at 8 columns, 2.3 at 4).
int fft_pick_c_log(const FftKernelInfo &info, int prec, bool cfast, unsigned B0, long long nbatch, long long in_stride_elems,
                   long long out_stride_elems) {
  const long long esz = prec == PREC_F64 ? 16 : 8;
  long long cmax = std::max(1, info.maxt / info.T);
  cmax = std::min<long long>(cmax, std::max<long long>(1, 220 * 1024 / ((long long)info.colsize * esz)));
  long long c = 1;
  if (cfast) {
    // 128 bytes per index when a side strides by a megabyte or more - if that still leaves room for a ring of two
    // slots (512 points), or if both sides do (1024 points: 3.3 vs 3.0 TB/s); otherwise 64 bytes keeps two CTAs per SM
    const bool big_in = in_stride_elems * esz >= (1 << 20), big_out = out_stride_elems * esz >= (1 << 20);
    const bool ring_fits = 2 * (128 / esz) * (long long)info.colsize * esz <= 220 * 1024;
    const long long want = (((big_in || big_out) && ring_fits) || (big_in && big_out) ? 128 : 64) / esz;
    while (c * 2 <= want && c * 2 <= cmax && B0 % (unsigned)(c * 2) == 0) c *= 2;
  } else {
    const long long want = std::max<long long>(std::max(1, 64 / info.T), 16384 / ((long long)info.N * esz));
    while (c * 2 <= want && c * 2 <= cmax && nbatch % (c * 2) == 0) c *= 2;
  }
  int lg = 0;
  while ((1LL << lg) < c) ++lg;
  return lg;
}

static cudaError_t dispatch(int N, int prec, const FftArgs &args, long long nbatch, cudaStream_t stream, FftShape *shape_only) {
  if (nbatch & ((1LL << args.c_log) - 1)) return cudaErrorInvalidValue;
  switch (N) {
#define X(N, ...) case N: return fft_launch_##N(prec, args, nbatch, stream, shape_only);
    OFFTB_FFT_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t fft_launch(int N, int prec, const FftArgs &args, long long nbatch, cudaStream_t stream) {
  return dispatch(N, prec, args, nbatch, stream, nullptr);
}

cudaError_t fft_shape(int N, int prec, const FftArgs &args, long long nbatch, FftShape *shape) {
  return dispatch(N, prec, args, nbatch, nullptr, shape);
}

}  // namespace offtb
