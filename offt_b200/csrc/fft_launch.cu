// Length dispatch for the batched 1-D FFT kernels.
#include "fft_configs.h"
#include "fft_launch.h"

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
