// One translation unit per transform length (compiled with -DOFFTB_INST_N=<N>), so the
// lengths build in parallel.  Defines offtb::fft_launch_<N> and offtb::fft_info_<N>.
#include "fft_configs.h"
#include "fft_launch.h"

#include <algorithm>
#include <cstdlib>

#ifndef OFFTB_INST_N
#error "compile with -DOFFTB_INST_N=<length>"
#endif

namespace offtb {

#define OFFTB_CAT2(a, b) a##b
#define OFFTB_CAT(a, b) OFFTB_CAT2(a, b)

// Ring depth and grid of one launch.  The ring wants two tiles in flight behind the one being
// transformed (depth 3) when shared memory allows; the grid is one wave of resident CTAs, each
// walking its share of the tiles.  OFFTB_DEPTH / OFFTB_CTAS_PER_SM override both for experiments.
template <typename T, class CFG, bool BULK>
static cudaError_t plan_one(FftArgs &args, long long nbatch, FftShape *shape) {
  const int C = 1 << args.c_log;
  const int threads = CFG::T * C;
  const size_t slot = (size_t)C * CFG::colsize() * sizeof(cx<T>);
  static int sm_count = 0, smem_optin = 0, env_depth = -1, env_ctas = -1, regs = 0, cfg_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!sm_count || dev != cfg_dev) {   // function attributes are per device
    cfg_dev = dev;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const char *e = getenv("OFFTB_DEPTH");
    env_depth = e ? atoi(e) : 0;
    e = getenv("OFFTB_CTAS_PER_SM");
    env_ctas = e ? atoi(e) : 0;
    // these kernels live on shared memory, not on L1: take the largest shared carve-out
    cudaFuncSetAttribute(fft_kernel<T, CFG, BULK>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, fft_kernel<T, CFG, BULK>);
    regs = fa.numRegs;
    smem_optin -= (int)fa.sharedSizeBytes;   // the kernel's static shared memory counts against the same limit
    cudaError_t ea = cudaFuncSetAttribute(fft_kernel<T, CFG, BULK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (ea != cudaSuccess) { sm_count = 0; return ea; }
  }
  const long long ntiles = nbatch >> args.c_log;
  shape->grid = 0;
  if (ntiles <= 0) return cudaSuccess;
  if (ntiles > 2147483647LL) return cudaErrorInvalidConfiguration;
  if (slot > (size_t)smem_optin) return cudaErrorInvalidConfiguration;
  // deepest ring that still leaves 16 warps resident per SM (the butterflies need them); else no ring
  const long long per_cta = (ntiles + sm_count - 1) / sm_count;   // tiles a CTA will see at least
  int depth = args.depth > 0 ? args.depth : env_depth, occ = 0;
  if (args.bulk_store) {
    // staged bulk stores need a slot that drains while the next tile is worked on: three slots (landing, butterflies,
    // draining) where shared memory allows, two otherwise; a tile that fills the SM alone keeps the direct stores
    const int fit = (int)std::min<size_t>(3, (size_t)smem_optin / slot);
    if (fit >= 2 && (env_depth <= 0 || env_depth >= 2)) {
      depth = env_depth >= 2 ? std::min(env_depth, fit) : fit;
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_kernel<T, CFG, BULK>, threads, slot * depth);
      if (e != cudaSuccess) return e;
    } else {
      args.bulk_store = 0;
    }
  }
  if (args.bulk_store) {
  } else if (depth > 0) {
    depth = std::min<int>(depth, std::min<int>(4, (int)(smem_optin / slot)));
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_kernel<T, CFG, BULK>, threads, slot * depth);
    if (e != cudaSuccess) return e;
  } else {
    for (depth = args.load_cfast ? 3 : 2; depth >= 1; --depth) {   // contiguous rows: 2 slots measured best
      if (slot * depth > (size_t)smem_optin || (depth > 1 && per_cta < depth)) continue;
      cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_kernel<T, CFG, BULK>, threads, slot * depth);
      if (e != cudaSuccess) return e;
      if (occ * threads >= 512 || depth == 1) break;
    }
  }
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  args.depth = depth;
  args.ntiles = (unsigned)ntiles;
  if (env_ctas > 0) occ = std::min(occ, env_ctas);
  long long grid = std::min<long long>(ntiles, (long long)occ * sm_count);
  if (args.grid_cap > 0) grid = std::min<long long>(grid, args.grid_cap);
  shape->threads = threads; shape->regs = regs; shape->smem = slot * depth; shape->depth = depth; shape->occ = occ;
  shape->grid = (unsigned)grid; shape->sm_count = sm_count;
  return cudaSuccess;
}

template <typename T, class CFG, bool BULK>
static cudaError_t launch_one(const FftArgs &args_in, long long nbatch, cudaStream_t stream, FftShape *shape_only) {
  FftArgs args = args_in;
  FftShape shape;
  cudaError_t e = plan_one<T, CFG, BULK>(args, nbatch, &shape);
  if constexpr (BULK) {
    // the tile does not leave room for a draining slot: the plain instantiation with direct stores takes over
    if (e == cudaSuccess && !args.bulk_store) {
      FftArgs plain = args_in;
      plain.bulk_store = 0;
      return launch_one<T, CFG, false>(plain, nbatch, stream, shape_only);
    }
  }
  if (shape_only) *shape_only = shape;
  if (e != cudaSuccess || shape_only || shape.grid == 0) return e;
  if (args.pdl & 2) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(shape.grid); cfg.blockDim = dim3(shape.threads); cfg.dynamicSmemBytes = shape.smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fft_kernel<T, CFG, BULK>, args);
  }
  fft_kernel<T, CFG, BULK><<<shape.grid, shape.threads, shape.smem, stream>>>(args);
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

cudaError_t OFFTB_CAT(fft_launch_, OFFTB_INST_N)(int prec, const FftArgs &args, long long nbatch, cudaStream_t stream, FftShape *shape_only) {
  if (args.bulk_store) {
    if (prec == PREC_F64) return launch_one<double, CfgD, true>(args, nbatch, stream, shape_only);
    return launch_one<float, CfgF, true>(args, nbatch, stream, shape_only);
  }
  if (prec == PREC_F64) return launch_one<double, CfgD, false>(args, nbatch, stream, shape_only);
  return launch_one<float, CfgF, false>(args, nbatch, stream, shape_only);
}

void OFFTB_CAT(fft_info_, OFFTB_INST_N)(int prec, FftKernelInfo *info) {
  info->N = CfgD::N; info->E = CfgD::E; info->T = CfgD::T; info->maxt = CfgD::MAXT;
  info->colsize = prec == PREC_F64 ? CfgD::colsize() : CfgF::colsize();
  info->ns = CfgD::NS;
  for (int s = 0; s < 4; ++s) info->radix[s] = CfgD::radix(s);
}

}  // namespace offtb
