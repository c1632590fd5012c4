// One translation unit per transform length (compiled with -DOFFTB_INST_N=<N>), so the
// lengths build in parallel.  Defines offtb::fft_launch_<N> and offtb::fft_info_<N>.
#include "fft_configs.h"
#include "fft_launch.h"

#ifndef OFFTB_INST_N
#error "compile with -DOFFTB_INST_N=<length>"
#endif

namespace offtb {

#define OFFTB_CAT2(a, b) a##b
#define OFFTB_CAT(a, b) OFFTB_CAT2(a, b)

template <typename T, class CFG>
static cudaError_t launch_one(const FftArgs &args, long long nbatch, cudaStream_t stream) {
  const int C = 1 << args.c_log;
  const size_t smem = (size_t)C * CFG::colsize() * sizeof(cx<T>);
  static size_t configured = 0;   // per instantiation
  static bool carved = false;
  if (!carved) {
    // these kernels live on shared memory, not on L1: take the largest shared carve-out
    cudaFuncSetAttribute(fft_kernel<T, CFG>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    carved = true;
  }
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(fft_kernel<T, CFG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = smem;
  }
  const long long grid = nbatch >> args.c_log;
  if (grid <= 0) return cudaSuccess;
  if (grid > 2147483647LL) return cudaErrorInvalidConfiguration;
  fft_kernel<T, CFG><<<(unsigned)grid, CFG::T * C, smem, stream>>>(args);
  return cudaGetLastError();
}

#define X(N, E, R0, R1, R2, R3, PD0, PD1, PD2, PF0, PF1, PF2, MAXT, MINB)                                   \
  OFFTB_IF_##N(                                                                                             \
      using CfgD = FftCfg<N, E, R0, R1, R2, R3, PD0, PD1, PD2, MAXT, MINB>;                                 \
      using CfgF = FftCfg<N, E, R0, R1, R2, R3, PF0, PF1, PF2, MAXT, MINB>;)
// expand OFFTB_IF_<N>(body) to body only for N == OFFTB_INST_N
#define OFFTB_IF_2(...)
#define OFFTB_IF_4(...)
#define OFFTB_IF_8(...)
#define OFFTB_IF_16(...)
#define OFFTB_IF_32(...)
#define OFFTB_IF_64(...)
#define OFFTB_IF_128(...)
#define OFFTB_IF_256(...)
#define OFFTB_IF_512(...)
#define OFFTB_IF_1024(...)
#define OFFTB_IF_2048(...)
#define OFFTB_IF_4096(...)
#define OFFTB_IF_8192(...)
#if OFFTB_INST_N == 2
#undef OFFTB_IF_2
#define OFFTB_IF_2(...) __VA_ARGS__
#elif OFFTB_INST_N == 4
#undef OFFTB_IF_4
#define OFFTB_IF_4(...) __VA_ARGS__
#elif OFFTB_INST_N == 8
#undef OFFTB_IF_8
#define OFFTB_IF_8(...) __VA_ARGS__
#elif OFFTB_INST_N == 16
#undef OFFTB_IF_16
#define OFFTB_IF_16(...) __VA_ARGS__
#elif OFFTB_INST_N == 32
#undef OFFTB_IF_32
#define OFFTB_IF_32(...) __VA_ARGS__
#elif OFFTB_INST_N == 64
#undef OFFTB_IF_64
#define OFFTB_IF_64(...) __VA_ARGS__
#elif OFFTB_INST_N == 128
#undef OFFTB_IF_128
#define OFFTB_IF_128(...) __VA_ARGS__
#elif OFFTB_INST_N == 256
#undef OFFTB_IF_256
#define OFFTB_IF_256(...) __VA_ARGS__
#elif OFFTB_INST_N == 512
#undef OFFTB_IF_512
#define OFFTB_IF_512(...) __VA_ARGS__
#elif OFFTB_INST_N == 1024
#undef OFFTB_IF_1024
#define OFFTB_IF_1024(...) __VA_ARGS__
#elif OFFTB_INST_N == 2048
#undef OFFTB_IF_2048
#define OFFTB_IF_2048(...) __VA_ARGS__
#elif OFFTB_INST_N == 4096
#undef OFFTB_IF_4096
#define OFFTB_IF_4096(...) __VA_ARGS__
#elif OFFTB_INST_N == 8192
#undef OFFTB_IF_8192
#define OFFTB_IF_8192(...) __VA_ARGS__
#else
#error "unsupported OFFTB_INST_N"
#endif

OFFTB_FFT_CONFIGS(X)
#undef X

cudaError_t OFFTB_CAT(fft_launch_, OFFTB_INST_N)(int prec, const FftArgs &args, long long nbatch, cudaStream_t stream) {
  if (prec == PREC_F64) return launch_one<double, CfgD>(args, nbatch, stream);
  return launch_one<float, CfgF>(args, nbatch, stream);
}

void OFFTB_CAT(fft_info_, OFFTB_INST_N)(int prec, FftKernelInfo *info) {
  info->N = CfgD::N; info->E = CfgD::E; info->T = CfgD::T; info->maxt = CfgD::MAXT;
  info->colsize = prec == PREC_F64 ? CfgD::colsize() : CfgF::colsize();
  info->ns = CfgD::NS;
  for (int s = 0; s < 4; ++s) info->radix[s] = CfgD::radix(s);
}

}  // namespace offtb
