// Host-side entry to the batched 1-D FFT kernels (fft_kernels.cuh).
#pragma once
#include <cuda_runtime.h>
#include "fft_kernels.cuh"

namespace offtb {

enum Precision { PREC_F64 = 64, PREC_F32 = 32 };

struct FftKernelInfo {
  int N, E, T, colsize, maxt;   // colsize: shared-memory elements per column
  int ns, radix[4];             // stage count and radices
};

// Host copy of the kernel's twiddle tables for length N (interleaved re, im in long double):
// for every stage s but the last, M_s entries exp(-2*pi*i*n'/(R_s*M_s)), concatenated.
// Returns the number of complex entries written (< N); out must hold 2*N long doubles.
int fft_twiddle_table(int N, int prec, long double *out);

// what one launch occupies: resolved ring depth, grid, and the per-CTA resources
struct FftShape {
  int threads = 0, regs = 0, depth = 0, occ = 0, sm_count = 0;
  size_t smem = 0;   // dynamic shared memory of the launch = depth ring slots
  unsigned grid = 0;
};

// Fills `info` for length N; returns false if N is not a supported length.
bool fft_kernel_info(int N, int prec, FftKernelInfo *info);

// log2 of the columns one CTA transforms (see fft_launch.cu)
int fft_pick_c_log(const FftKernelInfo &info, int prec, bool cfast, unsigned B0, long long nbatch, long long in_stride_elems,
                   long long out_stride_elems);

// Launches ceil(nbatch / 2^c_log) CTAs.  nbatch must be a multiple of 2^c_log.
// Returns cudaSuccess or the launch error.
cudaError_t fft_launch(int N, int prec, const FftArgs &args, long long nbatch, cudaStream_t stream);
// the same decisions without launching
cudaError_t fft_shape(int N, int prec, const FftArgs &args, long long nbatch, FftShape *shape);

// ---- any length, general block split (fft_generic.cu) ---------------------------------------------------------
#define OFFTB_GEN_MAX_STAGES 24
// n / d for n < 2^31 by one multiply-high and a shift (d fixed per launch; the generic kernel divides every element
// index of every stage by the same few numbers)
struct FastDiv {
  unsigned m, s, d;
};
inline FastDiv fastdiv_make(unsigned d) {
  FastDiv f;
  f.d = d;
  unsigned s = 0;
  while ((1ULL << s) < d) ++s;
  f.s = s;
  f.m = (unsigned)(((1ULL << 32) * ((1ULL << s) - d)) / d + 1);
  return f;
}
struct GenArgs {
  FftArgs a;          // maps (with the general split), peer table, flags; a.tw = full table exp(-2*pi*i*k/N), k < N
  int N, ns, cols;    // length, Stockham stages, columns per CTA
  int tw_in_smem;     // the table is copied behind the two buffers at kernel start
  FastDiv dN, dNh, dC, dNs[OFFTB_GEN_MAX_STAGES], dNsR[OFFTB_GEN_MAX_STAGES], dR[OFFTB_GEN_MAX_STAGES];   // divisors N, N/2+1, columns per CTA; Ns, Ns*R, R per stage
  int radix[OFFTB_GEN_MAX_STAGES];
  long long nbatch;
};
// factors of N in stage order; returns the stage count or -1
int fft_generic_factor(int N, int *radix, int max_stages);
// longest transform the shared-memory ping-pong holds
size_t fft_generic_max_n(int prec);
// N entries exp(-2*pi*i*k/N) as interleaved long double; returns N
int fft_generic_twiddle_table(int N, long double *out);
// launches (or, with shape_only, only sizes) the generic kernel; nbatch may be 0 (the launch then only waits and signals)
cudaError_t fft_generic_launch(int N, int prec, const FftArgs &args, long long nbatch, cudaStream_t stream, FftShape *shape_only);

// ---- O(N) step between an H-point complex transform and the N = 2H point real one (r2c_pass.cu) -----------------------
cudaError_t r2c_step_launch(int prec, bool backward, const void *in, void *out, const void *w, const FftMap &im, const FftMap &om, int H,
                            long long nbatch, cudaStream_t stream);

}  // namespace offtb
