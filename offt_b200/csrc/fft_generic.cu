// Batched 1-D FFT of ANY length with the general (uneven) block split - the path for what the power-of-two kernels of
// fft_kernels.cuh do not take: transform lengths with odd factors (the reference hands any N to FFTW,
// offt-compute.c:338, 416, 421, 443) and process grids that do not divide the grid evenly (the F/b/m bookkeeping of
// offt-compute.c:127-144, 1000-1013).
//
// One CTA holds `cols` columns of length N in shared memory and runs a Stockham autosort over the factors of N,
// ping-ponging between two buffers.  Stage with radix R after sub-transforms of length Ns: output
//   o = j_hi*Ns*R + t*Ns + k   (k < Ns, t < R)   is   sum_{q<R} x[j_hi*Ns + k + q*N/R] * w_N^{ q*(k + t*Ns)*N/(Ns*R) }
// - one thread per output, R complex multiply-adds from a full table w_N^k = exp(-2*pi*i*k/N) built in long double.
// That is O(N * sum of the factors) instead of O(N log N) butterflies with shared sub-expressions, and any prime factor
// works; this path is about coverage and exact parity, the bandwidth-bound headline configurations never take it.
// Loads and stores go through the same two-level address maps as the fast kernels (FftMap, with the general split),
// the same peer-slot table, flag waits and last-CTA signal, so every schedule of plan.cu runs unchanged.
#include "fft_launch.h"

#include <algorithm>
#include <cstdlib>

namespace offtb {

namespace {

__device__ __forceinline__ void split_n(const FftMap &m, int n, int &blk, int &lo) {
  if (m.gg > 0) {
    const int small = m.gF * (m.gg - m.gb);
    if (n < small) { blk = n / m.gF; lo = n - blk * m.gF; }
    else { const int r = n - small, q = r / (m.gF + 1); blk = (m.gg - m.gb) + q; lo = r - q * (m.gF + 1); }
  } else {
    blk = n >> m.n_lg;
    lo = n & ((1 << m.n_lg) - 1);
  }
}

__device__ __forceinline__ unsigned fdiv(unsigned n, const FastDiv &f) { return (__umulhi(n, f.m) + n) >> f.s; }

template <typename T>
__global__ void __launch_bounds__(256) fft_generic_kernel(const __grid_constant__ GenArgs g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const FftArgs &a = g.a;
  const int N = g.N, C = g.cols;
  const int tid = threadIdx.x, nthreads = blockDim.x;
  cx<T> *buf0 = reinterpret_cast<cx<T> *>(smem_raw);
  cx<T> *buf1 = buf0 + (size_t)C * N;
  // the twiddle table sits behind the two buffers when it fits (every multiply-add of every stage gathers from it)
  const cx<T> *tw = (const cx<T> *)a.tw;
  if (g.tw_in_smem) {
    cx<T> *stw = buf1 + (size_t)C * N;
    for (int k = tid; k < N; k += nthreads) stw[k] = tw[k];
    tw = stw;
  }
  const T cj = a.conj ? (T)-1 : (T)1;

  __shared__ void *s_tab[OFFTB_MAX_GROUP];
  for (int j = tid; j < OFFTB_MAX_GROUP; j += nthreads) s_tab[j] = a.out_split ? a.out_tab[j] : (void *)((cx<T> *)a.out + (long long)j * a.om.n_hi);
  if (a.wait_count > 0 && flag_wait(a.wait_flags, a.wait_count, a.wait_value, a.wait_timeout_ns, a.error_word, tid, nthreads)) return;
  __syncthreads();

  const long long ntiles = (g.nbatch + C - 1) / C;
  for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const long long b0 = tile * C;
    const int ncol = (int)(g.nbatch - b0 < C ? g.nbatch - b0 : C);
    const int total = ncol * N;
    // ---- load [column][n]
    const int Nh = N / 2 + 1;
    if (a.real_mode == 1) {
      // rows of N real numbers in the in-place r2c layout: real n of the row that starts at complex offset o is double 2*o + n
      for (int e = tid; e < total; e += nthreads) {
        const int c = (int)fdiv((unsigned)e, g.dN), n = e - c * N;
        const T re = ((const T *)a.in)[2 * map_b(a.im, (unsigned)(b0 + c)) + n];
        buf0[c * N + n] = cx<T>{re, (T)0};
      }
    } else if (a.real_mode == 2) {
      // N/2+1 complex points per row, completed to the Hermitian row X[N-n] = conj(X[n])
      for (int e = tid; e < ncol * Nh; e += nthreads) {
        int c, n;
        if (a.load_cfast) { c = e % ncol; n = e / ncol; } else { c = (int)fdiv((unsigned)e, g.dNh); n = e - c * Nh; }
        int blk, lo;
        split_n(a.im, n, blk, lo);
        cx<T> v = ((const cx<T> *)a.in)[map_b(a.im, (unsigned)(b0 + c)) + (long long)blk * a.im.n_hi + (long long)lo * a.im.n_lo];
        v.y *= cj;
        buf0[c * N + n] = v;
        if (n > 0 && 2 * n != N) buf0[c * N + N - n] = cx<T>{v.x, -v.y};
      }
    } else {
      for (int e = tid; e < total; e += nthreads) {
        int c, n;
        if (a.load_cfast) { n = ncol == C ? (int)fdiv((unsigned)e, g.dC) : e / ncol; c = e - n * ncol; } else { c = (int)fdiv((unsigned)e, g.dN); n = e - c * N; }
        int blk, lo;
        split_n(a.im, n, blk, lo);
        cx<T> v = ((const cx<T> *)a.in)[map_b(a.im, (unsigned)(b0 + c)) + (long long)blk * a.im.n_hi + (long long)lo * a.im.n_lo];
        v.y *= cj;
        buf0[c * N + n] = v;
      }
    }
    __syncthreads();
    // ---- Stockham stages
    cx<T> *src = buf0, *dst = buf1;
    int Ns = 1;
    for (int s = 0; s < g.ns; ++s) {
      const int R = g.radix[s];
      const int NR = N / R, step = N / (Ns * R);
      for (int e = tid; e < total; e += nthreads) {
        const int c = (int)fdiv((unsigned)e, g.dN), o = e - c * N;
        // Ry rule (offt-compute.c:1484, 1708): columns outside the window are moved, not transformed
        bool transform = true;
        if (a.ry_level >= 0) {
          const int r = (a.ry_x0 + (int)digit_b(a.im, (unsigned)(b0 + c), a.ry_level)) % 10;
          transform = a.ry_lo <= r && r < a.ry_hi;
        }
        if (!transform) { dst[e] = src[e]; continue; }
        const int oq = (int)fdiv((unsigned)o, g.dNs[s]), k = o - oq * Ns;            // o = (jhi*R + t)*Ns + k
        const int jhi = (int)fdiv((unsigned)oq, g.dR[s]), t = oq - jhi * R;
        const cx<T> *x = src + c * N + jhi * Ns + k;
        const int twb = (k + t * Ns) * step;   // < N
        cx<T> acc = x[0];
        int ti = 0;
        for (int q = 1; q < R; ++q) {
          ti += twb;
          if (ti >= N) ti -= N;
          acc = cadd(acc, cmul(x[q * NR], tw[ti]));
        }
        dst[e] = acc;
      }
      __syncthreads();
      cx<T> *tmp = src; src = dst; dst = tmp;
      Ns *= R;
    }
    // ---- store
    if (a.real_mode == 2) {
      for (int e = tid; e < total; e += nthreads) {
        const int c = (int)fdiv((unsigned)e, g.dN), n = e - c * N;
        ((T *)s_tab[0])[2 * map_b(a.om, (unsigned)(b0 + c)) + n] = src[c * N + n].x;
      }
    } else {
      const int Nst = a.real_mode == 1 ? Nh : N;   // a real row's spectrum is stored up to the Nyquist point only
      for (int e = tid; e < ncol * Nst; e += nthreads) {
        int c, n;
        if (a.store_cfast) { n = ncol == C ? (int)fdiv((unsigned)e, g.dC) : e / ncol; c = e - n * ncol; }
        else { c = (int)fdiv((unsigned)e, a.real_mode == 1 ? g.dNh : g.dN); n = e - c * Nst; }
        int blk, lo;
        split_n(a.om, n, blk, lo);
        cx<T> v = src[c * N + n];
        v.y *= cj;
        ((cx<T> *)s_tab[blk])[map_b(a.om, (unsigned)(b0 + c)) + (long long)lo * a.om.n_lo] = v;
      }
    }
    __syncthreads();
  }
  if (a.signal_count > 0) {
    __threadfence_system();
    __syncthreads();
    if (tid == 0) {
      const unsigned prev = atomicAdd(a.done_counter, 1u);
      if (prev + 1 == gridDim.x) {
        *a.done_counter = 0;
        __threadfence_system();
        for (int j = 0; j < a.signal_count; ++j) *(volatile unsigned *)a.signal_ptrs[j] = a.signal_value;
      }
    }
  }
}

}  // namespace

// radices of the Stockham stages: 4s first, then 2, 3, 5, 7 and whatever primes are left
int fft_generic_factor(int N, int *radix, int max_stages) {
  int n = N, ns = 0;
  auto push = [&](int r) { if (ns < max_stages) radix[ns] = r; ++ns; };
  while (n % 4 == 0) { push(4); n /= 4; }
  for (int r : {2, 3, 5, 7})
    while (n % r == 0) { push(r); n /= r; }
  for (int r = 11; (long long)r * r <= n; r += 2)
    while (n % r == 0) { push(r); n /= r; }
  if (n > 1) push(n);
  return ns <= max_stages ? ns : -1;
}

size_t fft_generic_max_n(int prec) {
  const size_t esz = prec == PREC_F64 ? 16 : 8;
  return (size_t)200 * 1024 / (2 * esz);
}

cudaError_t fft_generic_launch(int N, int prec, const FftArgs &args, long long nbatch, cudaStream_t stream, FftShape *shape_only) {
  GenArgs g;
  g.a = args;
  g.N = N;
  g.nbatch = nbatch;
  g.ns = fft_generic_factor(N, g.radix, OFFTB_GEN_MAX_STAGES);
  if (g.ns < 0 || N < 1 || (size_t)N > fft_generic_max_n(prec)) return cudaErrorInvalidValue;
  const size_t esz = prec == PREC_F64 ? 16 : 8;
  static int sm_count = 0, smem_optin = 0, cfg_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (!sm_count || dev != cfg_dev) {
    cfg_dev = dev;
    cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    smem_optin -= 1024;
    cudaError_t e1 = cudaFuncSetAttribute(fft_generic_kernel<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    cudaError_t e2 = cudaFuncSetAttribute(fft_generic_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_optin);
    if (e1 != cudaSuccess || e2 != cudaSuccess) { sm_count = 0; return e1 != cudaSuccess ? e1 : e2; }
  }
  // columns per CTA: about 2048 points per tile, at least 64 contiguous bytes per transform index on strided sides
  long long cols = std::max<long long>(1, 2048 / N);
  if (args.load_cfast || args.store_cfast) cols = std::max<long long>(cols, (long long)(64 / esz));
  while (cols > 1 && 2 * (size_t)cols * N * esz > (size_t)smem_optin / 2) --cols;
  cols = std::min<long long>(cols, std::max<long long>(nbatch, 1));
  if (2 * (size_t)cols * N * esz > (size_t)smem_optin) return cudaErrorInvalidConfiguration;
  g.cols = (int)cols;
  g.dN = fastdiv_make((unsigned)N); g.dNh = fastdiv_make((unsigned)(N / 2 + 1)); g.dC = fastdiv_make((unsigned)cols);
  for (int s = 0, Ns = 1; s < g.ns; ++s) {
    g.dNs[s] = fastdiv_make((unsigned)Ns); g.dR[s] = fastdiv_make((unsigned)g.radix[s]); g.dNsR[s] = fastdiv_make((unsigned)(Ns * g.radix[s]));
    Ns *= g.radix[s];
  }
  g.tw_in_smem = (2 * (size_t)cols + 1) * N * esz <= (size_t)smem_optin ? 1 : 0;
  const size_t smem = (2 * (size_t)cols + (g.tw_in_smem ? 1 : 0)) * N * esz;
  const long long ntiles = (nbatch + cols - 1) / cols;
  int occ = 0;
  cudaError_t e = prec == PREC_F64 ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_generic_kernel<double>, 256, smem)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fft_generic_kernel<float>, 256, smem);
  if (e != cudaSuccess) return e;
  if (occ < 1) return cudaErrorLaunchOutOfResources;
  long long grid = std::max<long long>(1, std::min<long long>(ntiles, (long long)occ * sm_count));   // a launch with nothing to move still waits and signals
  if (args.grid_cap > 0) grid = std::min<long long>(grid, args.grid_cap);
  if (shape_only) {
    shape_only->threads = 256; shape_only->regs = 64; shape_only->smem = smem; shape_only->depth = 1; shape_only->occ = occ;
    shape_only->grid = (unsigned)grid; shape_only->sm_count = sm_count;
    return cudaSuccess;
  }
  if (prec == PREC_F64) fft_generic_kernel<double><<<(unsigned)grid, 256, smem, stream>>>(g);
  else fft_generic_kernel<float><<<(unsigned)grid, 256, smem, stream>>>(g);
  return cudaGetLastError();
}

int fft_generic_twiddle_table(int N, long double *out) {
  const long double two_pi = 6.283185307179586476925286766559005768L;
  for (int k = 0; k < N; ++k) {
    const long double ang = two_pi * (long double)k / (long double)N;
    out[2 * k] = cosl(ang);
    out[2 * k + 1] = -sinl(ang);
  }
  return N;
}

}  // namespace offtb
