// Host-side entry to the batched 1-D FFT kernels (fft_kernels.cuh).
#pragma once
#include <cuda_runtime.h>
#include "fft_kernels.cuh"

namespace offtb {

enum Precision { PREC_F64 = 64, PREC_F32 = 32 };

struct FftKernelInfo {
  int N, E, T, colsize, maxt;   // colsize: shared-memory elements per column
};

// Fills `info` for length N; returns false if N is not a supported length.
bool fft_kernel_info(int N, int prec, FftKernelInfo *info);

// Launches ceil(nbatch / 2^c_log) CTAs.  nbatch must be a multiple of 2^c_log.
// Returns cudaSuccess or the launch error.
cudaError_t fft_launch(int N, int prec, const FftArgs &args, long long nbatch, cudaStream_t stream);

}  // namespace offtb
