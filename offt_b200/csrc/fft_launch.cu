// Length dispatch for the batched 1-D FFT kernels.
#include "fft_configs.h"
#include "fft_launch.h"

#include <cmath>

namespace offtb {

#define X(N, ...)                                                                                  \
  cudaError_t fft_launch_##N(int prec, const FftArgs &args, long long nbatch, cudaStream_t stream); \
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

cudaError_t fft_launch(int N, int prec, const FftArgs &args, long long nbatch, cudaStream_t stream) {
  if (nbatch & ((1LL << args.c_log) - 1)) return cudaErrorInvalidValue;
  switch (N) {
#define X(N, ...) case N: return fft_launch_##N(prec, args, nbatch, stream);
    OFFTB_FFT_CONFIGS(X)
#undef X
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace offtb
