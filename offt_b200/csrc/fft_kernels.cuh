// Batched 1-D Stockham-style FFT kernels for sm_100a (power-of-two lengths).
//
// One CTA transforms C "columns" (independent length-N transforms).  Every
// thread keeps E complex points in registers; a transform of length
// N = R0*R1*R2*R3 runs as up to four register-butterfly stages (radix 2..32)
// with a shared-memory exchange between consecutive stages.  The first stage
// loads straight from HBM into registers and the last stage stores straight
// from registers, so each point crosses HBM exactly once in each direction.
//
// The kernel is persistent: a CTA walks tiles blockIdx.x, blockIdx.x + gridDim.x, ... and keeps
// the inputs of the next `depth - 1` tiles in flight with cp.async (LDGSTS) into a ring of
// shared-memory slots while it transforms the current one.  Every thread copies exactly the points
// it will consume into private slots ([point][thread], conflict-free), so landing needs only
// cp.async.wait_group, no barrier; once the points are in registers the slot becomes the tile's
// exchange buffer.  HBM reads therefore overlap the butterflies, the exchanges and the stores
// of the previous tiles without costing registers.
//
// Index algebra (decimation in frequency, digits k_s of the output index):
//   stage s works on N/R_s butterflies beta = n' + M_s*K, n' < M_s = N/(R_0..R_s),
//   K = k_0 + R_0*k_1 + ... (digits produced so far); it reads the R_s points
//   n' + M_s*i of sub-problem K, multiplies output k_s by w_N^(P_s*n'*k_s),
//   P_s = R_0..R_{s-1}, and files it for butterfly beta' = n'' + M_{s+1}*(K + P_s*k_s)
//   of the next stage as its input i = n' / M_{s+1} (n'' = n' % M_{s+1}).
//   The shared layout is [i][beta'] with row pitch N/R_{s+1} + PAD_s, so the
//   reading side is always unit-stride in beta'.  After the last stage the
//   output index is k = beta + (N/R_last)*k_last: unit-stride in beta again.
//
// Addressing is a two-level affine map on both sides (see FftMap), which is what
// fuses the reference's pack / unpack loops (offt-compute.c:1015-1032, 1100-1116,
// 1307-1311, 1382-1385, 1773-1776, 2055-2058, 2447-2450, 2686-2689) into the
// transform's own loads and stores.  Lanes run either along the transform index
// ("n-fast", contiguous rows) or along the column index ("c-fast", strided
// axis); when the two sides differ the result is turned through shared memory.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <utility>

#include "roots32.h"

namespace offtb {

template <typename T> struct cx;
template <> struct __align__(16) cx<double> { double x, y; };
template <> struct __align__(8) cx<float> { float x, y; };

// element (n, b) of a batch lives at
//   off + (n >> n_lg)*n_hi + (n & ((1<<n_lg)-1))*n_lo + b0*s0 + b1*s1 + b2*s2,
//   b = b0 + B0*(b1 + B1*b2),  all in complex elements.  B0, B1 need not be powers of two
//   (ragged tiles), the split of n is (blocks of an even division always are).
struct FftMap {
  long long off;
  long long n_hi, n_lo;
  long long s0, s1, s2;
  int n_lg;
  unsigned B0, B1;
  // General split of the transform index (fft_generic.cu only; the power-of-two kernels need gg == 0): when gg > 0 the
  // index belongs to one of gg blocks of which the first gg-gb hold gF items and the last gb hold gF+1 - the
  // ownership rule of an uneven division (offt-compute.c:141-144, 1000-1013) - and element n lives at
  // block(n)*n_hi + (n - first(block))*n_lo.
  int gF, gb, gg;
};

#define OFFTB_MAX_GROUP 16   // ranks in one exchange group (one NVSwitch box holds 8)

struct FftArgs {
  const void *in;
  void *out;
  const void *tw;  // cx<T>[<N]: the per-stage tables of FftCfg::twoff, see fft_twiddle_table()
  FftMap im, om;
  int c_log;       // log2(columns per CTA)
  int depth;       // shared-memory ring slots (1: no prefetch)
  int grid_cap;    // > 0: at most this many CTAs in the grid (launches that share the SMs with another kernel)
  unsigned ntiles; // batch / columns per CTA
  int load_cfast, store_cfast;
  int conj;        // 1: backward transform via conj(FFT(conj(x)))
  // Ry rule of the reference (offt-compute.c:1484, 1708): transform a column only if
  // lo <= (x % 10) < hi, x = ry_x0 + batch digit `ry_level`; ry_level < 0: always
  int ry_level, ry_x0, ry_lo, ry_hi;
  // Fused exchange (plan.cu): when out_split is set, block a = n >> om.n_lg of the output does not live at
  // out + a*om.n_hi but in out_tab[a] - the receive slot of peer a, mapped over NVLink (or this rank's own).
  // Before touching memory every CTA waits until wait_flags[j] >= wait_value for all j < wait_count (peers
  // have released the slots); the last CTA to finish publishes signal_value to signal_ptrs[0..signal_count).
  int out_split;
  void *out_tab[OFFTB_MAX_GROUP];
  const unsigned *wait_flags;
  int wait_count;
  unsigned wait_value;
  unsigned *signal_ptrs[OFFTB_MAX_GROUP];
  int signal_count;
  unsigned signal_value;
  unsigned *done_counter;   // zeroed device word, returns to zero after the launch
  // A flag that does not arrive within wait_timeout_ns (0: wait for ever) makes the CTA store a code into *error_word
  // (host-visible) and leave without touching data or signalling; the host reports it after the stream has drained.
  unsigned long long wait_timeout_ns;
  unsigned *error_word;
  // dependent-launch chain (plan.cu).  Bit 0: once this launch's CTAs are past their flag wait the next launch of
  // the stream may start filling the SMs this one leaves; bit 1 (host side): this launch itself is allowed to start
  // before its predecessor has drained
  int pdl;
  // Staged bulk stores (writer launches of the fused exchange).  Remote st.global back-pressures the warps that
  // issue it and everything queued behind them in the SM's load/store path, so a CTA that scatters its tile over
  // NVLink cannot load or transform its next tile meanwhile (measured: 580-600 GB/s per direction against 717 for
  // a kernel that does nothing but store).  With bulk_store the last stage files its results in the tile's
  // shared-memory slot in destination order - block a of the output is one contiguous run there - and a few
  // threads hand each run to the TMA engine (cp.async.bulk shared -> global); the slot drains on its own while
  // the CTA works on the next tile in the next slot.  Needs a ring of >= 2 slots; 3 keeps one tile landing, one
  // in the butterflies and one draining.  Set by the host when the store map allows it (fft_inst.cu, plan_one).
  int bulk_store;
  // z pass of a real-to-complex plan (fft_generic.cu only).  1: the input rows are N real numbers (the in-place r2c
  // layout, run-fft.c:53-55: double index 2*row_start + n) and only outputs 0..N/2 are stored; 2: the backward
  // transform of that - N/2+1 complex points in, completed to the Hermitian row, N real numbers out.
  int real_mode;
};

constexpr int ilog2(int n) { return n <= 1 ? 0 : 1 + ilog2(n >> 1); }
constexpr int brev(int v, int radix) {
  int r = 0;
  for (int m = radix >> 1; m > 0; m >>= 1) { r = (r << 1) | (v & 1); v >>= 1; }
  return r;
}

template <int N_, int E_, int R0_, int R1_, int R2_, int R3_, int PAD0_, int PAD1_, int PAD2_, int MAXT_, int MINB_>
struct FftCfg {
  static constexpr int N = N_, E = E_, T = N_ / E_, MAXT = MAXT_, MINB = MINB_;
  static constexpr int NS = R1_ == 1 ? 1 : (R2_ == 1 ? 2 : (R3_ == 1 ? 3 : 4));
  static constexpr int radix(int s) { return s == 0 ? R0_ : s == 1 ? R1_ : s == 2 ? R2_ : R3_; }
  static constexpr int pad(int s) { return s == 0 ? PAD0_ : s == 1 ? PAD1_ : PAD2_; }
  static constexpr int P(int s) { return s == 0 ? 1 : P(s - 1) * radix(s - 1); }   // digits already produced
  static constexpr int M(int s) { return N_ / (P(s) * radix(s)); }                 // sub-problem length left
  static constexpr int pitch(int s) { return N_ / radix(s + 1) + pad(s); }         // exchange s -> s+1
  static constexpr int xsize(int s) { return radix(s + 1) * pitch(s); }
  // compact per-stage twiddle tables: stage s holds exp(-2*pi*i*n'/(R_s*M_s)), n' < M_s, at twoff(s)
  static constexpr int twoff(int s) { return s == 0 ? 0 : twoff(s - 1) + M(s - 1); }
  // twiddles a thread keeps in registers: one per butterfly of every stage but the last
  static constexpr int twregs(int s) { return s <= 0 ? 0 : twregs(s - 1) + E_ / radix(s - 1); }
  static constexpr int colsize() {
    int m = N_ + 1;  // the turn buffer of transposing launches
    for (int s = 0; s + 1 < NS; ++s) m = xsize(s) > m ? xsize(s) : m;
    return m;
  }
  static_assert(R0_ * R1_ * R2_ * R3_ == N_, "radices must multiply to N");
  static_assert(E_ % R0_ == 0 && E_ % R1_ == 0 && E_ % R2_ == 0 && E_ % R3_ == 0, "E must be a multiple of every radix");
};

__device__ __forceinline__ cx<double> ldg_cx(const cx<double> *p) {
  double2 d = __ldg(reinterpret_cast<const double2 *>(p));
  return {d.x, d.y};
}
__device__ __forceinline__ cx<float> ldg_cx(const cx<float> *p) {
  float2 d = __ldg(reinterpret_cast<const float2 *>(p));
  return {d.x, d.y};
}
// data stores: the pointer comes out of a table, so tell the compiler it is global memory (STG instead of generic ST)
template <typename T> __device__ __forceinline__ void st_global(cx<T> *p, cx<T> v) {
#ifdef OFFTB_ASSUME_GLOBAL_ST
  // STG instead of the generic ST: measured slower on B200 (512^3 y pass 5.57 -> 4.83 TB/s, transposing x pass
  // 4.43 -> 3.75, profiles/r02_kernel_ab.md), so the generic store stays
  __builtin_assume(__isGlobal(p));
#endif
  *p = v;
}
// Experiment switches (tools/variants.sh; the shipped build leaves both at 0, see profiles/r02_kernel_ab.md):
// OFFTB_CPASYNC_L2 = 128 / 256 adds the L2 prefetch-size hint to the ring fills, OFFTB_TILE_CHUNK_LOG = q makes a CTA
// walk runs of 2^q adjacent tiles (columns next to each other in memory) instead of one tile per round of the grid.
#ifndef OFFTB_CPASYNC_L2
#define OFFTB_CPASYNC_L2 0
#endif
#ifndef OFFTB_TILE_CHUNK_LOG
#define OFFTB_TILE_CHUNK_LOG 0
#endif
#if OFFTB_CPASYNC_L2 == 256
#define OFFTB_CPASYNC_HINT ".L2::256B"
#elif OFFTB_CPASYNC_L2 == 128
#define OFFTB_CPASYNC_HINT ".L2::128B"
#else
#define OFFTB_CPASYNC_HINT ""
#endif
__device__ __forceinline__ void cp_async_cx(cx<double> *smem_dst, const cx<double> *gsrc) {
  asm volatile("cp.async.cg.shared.global" OFFTB_CPASYNC_HINT " [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_cx(cx<float> *smem_dst, const cx<float> *gsrc) {
  asm volatile("cp.async.ca.shared.global" OFFTB_CPASYNC_HINT " [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of this thread's groups are still in flight (pending < 4)
__device__ __forceinline__ void cp_async_wait(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
  }
}
__device__ __forceinline__ void bulk_store_issue(void *gdst, const void *smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(__cvta_generic_to_global(gdst)),
               "r"((unsigned)__cvta_generic_to_shared(smem_src)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most one of this thread's bulk groups still reads its shared-memory source
__device__ __forceinline__ void bulk_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ unsigned long long global_timer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
template <typename T> __device__ __forceinline__ cx<T> cadd(cx<T> a, cx<T> b) { return {a.x + b.x, a.y + b.y}; }
template <typename T> __device__ __forceinline__ cx<T> csub(cx<T> a, cx<T> b) { return {a.x - b.x, a.y - b.y}; }
template <typename T> __device__ __forceinline__ cx<T> cmul(cx<T> a, cx<T> w) {
  return {a.x * w.x - a.y * w.y, a.x * w.y + a.y * w.x};
}

// a * exp(-2*pi*i*K32/32), K32 known at compile time
template <typename T, int K32> __device__ __forceinline__ cx<T> mul_root(cx<T> a) {
  if constexpr (K32 == 0) return a;
  else if constexpr (K32 == 8) return {a.y, -a.x};
  else if constexpr (K32 == 4) {
    constexpr T h = (T)Root32<4>::re;
    return {(a.x + a.y) * h, (a.y - a.x) * h};
  } else if constexpr (K32 == 12) {
    constexpr T h = (T)Root32<4>::re;
    return {(a.y - a.x) * h, -(a.x + a.y) * h};
  } else {
    constexpr T c = (T)Root32<K32>::re, s = (T)Root32<K32>::im;
    return {a.x * c - a.y * s, a.x * s + a.y * c};
  }
}

// in-place radix-R DFT of v[OFF..OFF+R) by radix-2 DIF splitting; result k ends at OFF + brev(k, R)
template <typename T, int R, int OFF, int J> __device__ __forceinline__ void bfly_pair(cx<T> *v) {
  cx<T> a = v[OFF + J], b = v[OFF + J + R / 2];
  v[OFF + J] = cadd(a, b);
  v[OFF + J + R / 2] = mul_root<T, J * (32 / R)>(csub(a, b));
}
template <typename T, int R, int OFF, int... J> __device__ __forceinline__ void bfly_level(cx<T> *v, std::integer_sequence<int, J...>) {
  (bfly_pair<T, R, OFF, J>(v), ...);
}
template <typename T, int R, int OFF> __device__ __forceinline__ void bfly(cx<T> *v) {
  if constexpr (R > 1) {
    bfly_level<T, R, OFF>(v, std::make_integer_sequence<int, R / 2>{});
    bfly<T, R / 2, OFF>(v);
    bfly<T, R / 2, OFF + R / 2>(v);
  }
}
template <typename T, int R, int... U> __device__ __forceinline__ void bfly_all(cx<T> *v, std::integer_sequence<int, U...>) {
  (bfly<T, R, U * R>(v), ...);
}

__device__ __forceinline__ long long map_n(const FftMap &m, int n) {
  return (long long)(n >> m.n_lg) * m.n_hi + (long long)(n & ((1 << m.n_lg) - 1)) * m.n_lo;
}
__device__ __forceinline__ long long map_b(const FftMap &m, unsigned b) {
  const unsigned b0 = b % m.B0, r = b / m.B0;
  const unsigned b1 = r % m.B1, b2 = r / m.B1;
  return m.off + (long long)b0 * m.s0 + (long long)b1 * m.s1 + (long long)b2 * m.s2;
}
__device__ __forceinline__ unsigned digit_b(const FftMap &m, unsigned b, int level) {
  if (level == 0) return b % m.B0;
  const unsigned r = b / m.B0;
  return level == 1 ? r % m.B1 : r / m.B1;
}

// address of output point k of the column whose batch offset is bofs (elements): block k >> n_lg of the output
// starts at s_tab[block] - the peers' slots in a fused exchange, out + block*n_hi otherwise
template <typename T>
__device__ __forceinline__ cx<T> *out_ptr(const FftArgs &a, void *const *s_tab, long long bofs, int k) {
  return (cx<T> *)s_tab[k >> a.om.n_lg] + (bofs + (long long)(k & ((1 << a.om.n_lg) - 1)) * a.om.n_lo);
}

// Long transforms run without a ring (one slot per CTA): the slot is the tile's exchange buffer until the last stage has
// read its inputs, and from then on - through the last butterflies and all the stores - it is idle.  `early` is called at
// that point and starts the next tile's fill (kernel body); ON = false compiles it away for the lengths that have a ring.
template <bool ON, class F> struct EarlyFill {
  static constexpr bool on = ON;
  F &f;
  __device__ __forceinline__ void operator()() const { f(); }
};
// Measured (profiles/r02_kernel_ab.md): in-place y passes gain (1024 points +7 %, 2048 points +29 %), but the x pass of a
// 1024^3 grid on one GPU - stores at a 16 MiB plane stride over a 16 GiB array - drops from 4.3 to 3.0 TB/s with the fills
// queued ahead of its stores, and the launches of the fused exchange were not re-measured with it.  Compile-time
// experiment (-DOFFTB_EARLY_MIN_N=1024 -DOFFTB_EARLY_MAX_N=2048), off in the shipped build.
#ifndef OFFTB_EARLY_MIN_N
#define OFFTB_EARLY_MIN_N (1 << 30)
#endif
#ifndef OFFTB_EARLY_MAX_N
#define OFFTB_EARLY_MAX_N 2048
#endif

template <typename T, class CFG, bool BULK, int S, class EARLY>
__device__ __forceinline__ void fft_stage(cx<T> (&v)[CFG::E], const cx<T> *wreg, const FftArgs &a, cx<T> *sm,
                                          void *const *s_tab, long long bofs, unsigned bblock, int t, int s_mul, int s_base, T cj,
                                          const EARLY &early) {
  constexpr int N = CFG::N, E = CFG::E, TT = CFG::T, NS = CFG::NS;
  constexpr int R = CFG::radix(S), NU = E / R, P = CFG::P(S), M = CFG::M(S);

  // ---- inputs (stage 0 arrives in registers)
  if constexpr (S > 0) {
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int beta = t + TT * u;
#pragma unroll
      for (int i = 0; i < R; ++i) v[u * R + i] = sm[(i * CFG::pitch(S - 1) + beta) * s_mul + s_base];
    }
    if constexpr (EARLY::on && S == NS - 1 && !BULK) {
      if (a.load_cfast == a.store_cfast) early();   // direct stores: nothing reads the slot any more
    }
  }
  // ---- butterflies
  bfly_all<T, R>(v, std::make_integer_sequence<int, NU>{});

  if constexpr (S < NS - 1) {
    constexpr int Rn = CFG::radix(S + 1), Mn = M / Rn;
    if constexpr (S > 0) __syncthreads();  // everyone has read this stage's inputs
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int beta = t + TT * u;
      const int np = beta & (M - 1), K = beta / M;
      const int npp = np & (Mn - 1), inext = np / Mn;
      const int abase = inext * CFG::pitch(S) + npp + Mn * K;
      // w^k for k = 1..R-1 by repeated multiplication from the thread's resident twiddle
      const cx<T> w1 = wreg[CFG::twregs(S) + u];
      cx<T> w = w1;
      sm[abase * s_mul + s_base] = v[u * R];
#pragma unroll
      for (int k = 1; k < R; ++k) {
        const cx<T> e = cmul(v[u * R + brev(k, R)], w);
        if (k + 1 < R) w = cmul(w, w1);
        sm[(abase + Mn * P * k) * s_mul + s_base] = e;
      }
    }
    __syncthreads();
    fft_stage<T, CFG, BULK, S + 1>(v, wreg, a, sm, s_tab, bofs, bblock, t, s_mul, s_base, cj, early);
  } else {
    // ---- last stage: output index beta + (N/R)*k
    if constexpr (BULK) {
      // file the results in the slot in destination order ([k][column] for strided launches, [column][k] for rows),
      // then one warp hands every block's run to the TMA engine
      if constexpr (NS > 1) __syncthreads();   // everyone has read this stage's inputs out of the slot
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int beta = t + TT * u;
#pragma unroll
        for (int pos = 0; pos < R; ++pos) {
          cx<T> e = v[u * R + pos];
          e.y *= cj;
          sm[(beta + (N / R) * brev(pos, R)) * s_mul + s_base] = e;
        }
      }
      fence_proxy_async();   // the generic-proxy writes above are ordered before the async-proxy reads below
      __syncthreads();
      if (threadIdx.x < 32) {
        const int C = 1 << a.c_log;
        const int nlo = 1 << a.om.n_lg, nblk = N >> a.om.n_lg;
        const int issuers = (TT << a.c_log) < 32 ? (TT << a.c_log) : 32;   // short transforms run CTAs of fewer than 32 threads
        if (a.store_cfast) {
          const long long b0ofs = map_b(a.om, bblock);
          for (int blk = threadIdx.x; blk < nblk; blk += issuers)
            bulk_store_issue((cx<T> *)s_tab[blk] + b0ofs, sm + (size_t)blk * nlo * C, (unsigned)(nlo * C * sizeof(cx<T>)));
        } else {
          for (int j = threadIdx.x; j < nblk * C; j += issuers) {
            const int cc = j / nblk, blk = j - cc * nblk;
            bulk_store_issue((cx<T> *)s_tab[blk] + map_b(a.om, bblock + cc), sm + (size_t)cc * CFG::colsize() + (size_t)blk * nlo,
                             (unsigned)(nlo * sizeof(cx<T>)));
          }
        }
        bulk_store_commit();
      }
    } else if (a.load_cfast == a.store_cfast) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int beta = t + TT * u;
#pragma unroll
        for (int pos = 0; pos < R; ++pos) {
          cx<T> e = v[u * R + pos];
          e.y *= cj;
          st_global(out_ptr<T>(a, s_tab, bofs, beta + (N / R) * brev(pos, R)), e);
        }
      }
    } else {
      // turn through shared memory: [column][k] with odd pitch N+1
      if constexpr (NS > 1) __syncthreads();
      const int C = 1 << a.c_log;
      const int c = a.load_cfast ? (int)(threadIdx.x & (C - 1)) : (int)(threadIdx.x / TT);
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int beta = t + TT * u;
#pragma unroll
        for (int pos = 0; pos < R; ++pos) sm[c * (N + 1) + beta + (N / R) * brev(pos, R)] = v[u * R + pos];
      }
      __syncthreads();
      const int nthreads = TT << a.c_log;
#pragma unroll
      for (int e = 0; e < E; ++e) {
        const int flat = threadIdx.x + e * nthreads;
        int cc, k;
        if (a.store_cfast) { cc = flat & (C - 1); k = flat >> a.c_log; }
        else { k = flat & (N - 1); cc = flat / N; }
        cx<T> val = sm[cc * (N + 1) + k];
        val.y *= cj;
        st_global(out_ptr<T>(a, s_tab, map_b(a.om, bblock + cc), k), val);
      }
    }
  }
}

// the thread's resident twiddles: entry twregs(s) + u is exp(-2*pi*i*n'/(R_s*M_s)) for its butterfly u of stage s
template <typename T, class CFG, int S>
__device__ __forceinline__ void load_twiddles(cx<T> *wreg, const cx<T> *__restrict__ tw, int t) {
  if constexpr (S < CFG::NS - 1) {
    constexpr int NU = CFG::E / CFG::radix(S), M = CFG::M(S);
#pragma unroll
    for (int u = 0; u < NU; ++u) wreg[CFG::twregs(S) + u] = ldg_cx(&tw[CFG::twoff(S) + ((t + CFG::T * u) & (M - 1))]);
    load_twiddles<T, CFG, S + 1>(wreg, tw, t);
  }
}

// Peers release the slots a launch writes (or fill the ones it reads) with system-scope stores to this rank's flag
// words; every CTA waits for all of them before it touches memory.  A peer that never answers (a rank died,
// mismatched plans) must neither hang the GPU nor kill the context: after wait_timeout_ns (0: wait for ever, like the
// reference's MPI_Wait) the failure is recorded where the host finds it and the CTA leaves.  Kept out of line so that
// the transform's register allocation does not depend on it.  Returns true when the CTA must give up.
static __device__ __noinline__ bool flag_wait(const unsigned *wait_flags, int wait_count, unsigned wait_value,
                                              unsigned long long wait_timeout_ns, unsigned *error_word, int tid, int nthreads) {
  int gave_up = 0;
  for (int j = tid; j < wait_count; j += nthreads) {
    const volatile unsigned *f = wait_flags + j;
    const unsigned long long t0 = global_timer_ns();
    while ((int)(*f - wait_value) < 0) {
      __nanosleep(100);
      if (wait_timeout_ns && global_timer_ns() - t0 > wait_timeout_ns) {
        if (error_word) *(volatile unsigned *)error_word = 1u + (unsigned)j;
        gave_up = 1;
        break;
      }
    }
  }
  __threadfence_system();
  return __syncthreads_or(gave_up) != 0;
}

template <typename T, class CFG, bool BULK>
__global__ void __launch_bounds__(CFG::MAXT, CFG::MINB) fft_kernel(const __grid_constant__ FftArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  cx<T> *sm_all = reinterpret_cast<cx<T> *>(smem_raw);
  constexpr int E = CFG::E, TT = CFG::T, N = CFG::N, R0 = CFG::radix(0), NU0 = E / R0;
  const int C = 1 << a.c_log;
  const int tid = threadIdx.x;
  const int nthreads = TT << a.c_log;
  int t, c;
  if (a.load_cfast) { c = tid & (C - 1); t = tid >> a.c_log; }
  else { t = tid & (TT - 1); c = tid / TT; }
  const int slot_elems = C * CFG::colsize();
  const int s_mul = a.load_cfast ? C : 1;
  const int s_base = a.load_cfast ? c : c * CFG::colsize();
  const T cj = a.conj ? (T)-1 : (T)1;
  const int depth = a.depth;

  __shared__ void *s_tab[OFFTB_MAX_GROUP];
  const unsigned ntiles = a.ntiles;
  for (int j = tid; j < OFFTB_MAX_GROUP; j += nthreads)
    s_tab[j] = a.out_split ? a.out_tab[j] : (void *)((cx<T> *)a.out + (long long)j * a.om.n_hi);
  if (a.wait_count > 0 && flag_wait(a.wait_flags, a.wait_count, a.wait_value, a.wait_timeout_ns, a.error_word, tid, nthreads)) return;   // timed out: leave without touching data or signalling
  // dependent launch: the next kernel of the stream does not consume this one's output (flags order them), let it in
  if (a.pdl & 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  __syncthreads();

  constexpr int NW = CFG::twregs(CFG::NS - 1) > 0 ? CFG::twregs(CFG::NS - 1) : 1;
  cx<T> wreg[NW];
  load_twiddles<T, CFG, 0>(wreg, (const cx<T> *)a.tw, t);

  // this thread's E points of `tile` -> its private places in ring slot `slot`
  auto prefetch = [&](unsigned tile, int slot) {
    if (tile < ntiles) {
      const cx<T> *gin = (const cx<T> *)a.in + map_b(a.im, (tile << a.c_log) + c);
      cx<T> *dst = sm_all + slot * slot_elems + tid;
#pragma unroll
      for (int u = 0; u < NU0; ++u)
#pragma unroll
        for (int i = 0; i < R0; ++i)
          cp_async_cx(dst + (u * R0 + i) * nthreads, gin + map_n(a.im, t + TT * u + (N / R0) * i));
    }
    cp_async_commit();
  };

  // ring-less launches (depth 1): the next tile's fill starts as soon as the last stage has its inputs (EarlyFill)
  constexpr bool EARLY_ON = !BULK && CFG::N >= OFFTB_EARLY_MIN_N && CFG::N <= OFFTB_EARLY_MAX_N && CFG::NS > 1;
  unsigned next_tile = 0;
  bool primed = false;
  auto early_fn = [&]() {
    if (depth == 1) {
      __syncthreads();   // every thread has read the last exchange out of the slot
      prefetch(next_tile, 0);
      primed = true;
    }
  };
  const EarlyFill<EARLY_ON, decltype(early_fn)> early{early_fn};

  // One tile: its E points are in v, the slot `sm` is free to serve as exchange (and, for bulk launches, staging) buffer
  auto transform_tile = [&](cx<T> (&v)[E], cx<T> *sm, unsigned tile) {
    const unsigned bblock = tile << a.c_log;
    const long long bofs = map_b(a.om, bblock + c);

    // Ry rule: a tile-uniform choice between transforming and merely moving its columns
    bool transform = true;
    if (a.ry_level >= 0) {
      const int r = (a.ry_x0 + (int)digit_b(a.im, bblock, a.ry_level)) % 10;
      transform = a.ry_lo <= r && r < a.ry_hi;
    }
    if (transform) {
      fft_stage<T, CFG, BULK, 0>(v, wreg, a, sm, s_tab, bofs, bblock, t, s_mul, s_base, cj, early);
    } else {
#pragma unroll
      for (int u = 0; u < NU0; ++u)
#pragma unroll
        for (int i = 0; i < R0; ++i) {
          cx<T> e = v[u * R0 + i];
          e.y *= cj;   // undo the conjugation of the load
          st_global(out_ptr<T>(a, s_tab, bofs, t + TT * u + (N / R0) * i), e);
        }
      if constexpr (BULK) { if (tid < 32) bulk_store_commit(); }   // an empty group keeps the per-tile group count in step
    }
  };

  if constexpr (!BULK) {
    // Plain launches: a slot is free again as soon as its tile's exchanges are over, so the ring keeps depth-1 tiles
    // in flight behind the current one.
#if OFFTB_TILE_CHUNK_LOG > 0
    // the CTA's k-th tile: runs of 2^q adjacent tiles, the runs dealt round robin over the grid
    auto tile_at = [&](unsigned k) {
      return (((k >> OFFTB_TILE_CHUNK_LOG) * gridDim.x + blockIdx.x) << OFFTB_TILE_CHUNK_LOG) | (k & ((1u << OFFTB_TILE_CHUNK_LOG) - 1));
    };
    unsigned kth = 0;
    unsigned tile = tile_at(0);
    for (int d = 0; d + 1 < depth; ++d) prefetch(tile_at(d), d);
    int slot = 0;
    for (; tile < ntiles; tile = tile_at(++kth)) {
#define OFFTB_TILE_AHEAD(n) tile_at(kth + (unsigned)(n))
#else
    unsigned tile = blockIdx.x;
    for (int d = 0; d + 1 < depth; ++d) prefetch(tile + d * gridDim.x, d);
    int slot = 0;
    for (; tile < ntiles; tile += gridDim.x) {
#define OFFTB_TILE_AHEAD(n) (tile + (unsigned)(n) * gridDim.x)
#endif
      cx<T> *sm = sm_all + slot * slot_elems;
      if (depth == 1) {
        if (!(EARLY_ON && primed)) {
          __syncthreads();         // the previous tile's exchange data has been consumed
          prefetch(tile, 0);
        }
        primed = false;
        cp_async_wait(0);
      } else {
        cp_async_wait(depth - 2);  // this tile has landed (the younger groups may still fly)
      }
      cx<T> v[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        cx<T> x = sm[e * nthreads + tid];
        x.y *= cj;
        v[e] = x;
      }
      if (depth > 1 || CFG::NS > 1 || a.load_cfast != a.store_cfast) __syncthreads();   // the slot now serves as exchange buffer
      if (depth > 1) {
        const int ahead = slot == 0 ? depth - 1 : slot - 1;   // the slot the previous tile has just released
        prefetch(OFFTB_TILE_AHEAD(depth - 1), ahead);
      }
      if constexpr (EARLY_ON) next_tile = OFFTB_TILE_AHEAD(1);
      transform_tile(v, sm, tile);
      slot = slot + 1 == depth ? 0 : slot + 1;
    }
#undef OFFTB_TILE_AHEAD
  } else {
    // Bulk-store launches: the slot of the previous tile is still being drained by the TMA engine, so one slot
    // fewer is available for prefetching (depth-2 tiles ahead) and a slot is re-filled only after warp 0 has seen
    // its bulk group finish reading.
    const int ahead_tiles = depth - 2;
    unsigned tile = blockIdx.x;
    for (int d = 0; d < ahead_tiles; ++d) prefetch(tile + d * gridDim.x, d);
    int slot = 0;
    for (; tile < ntiles; tile += gridDim.x) {
      cx<T> *sm = sm_all + slot * slot_elems;
      if (ahead_tiles == 0) {
        if (tid < 32) bulk_store_wait_read1();   // the tile that used this slot two tiles ago has drained
        __syncthreads();
        prefetch(tile, slot);
        cp_async_wait(0);
      } else {
        cp_async_wait(ahead_tiles - 1);
      }
      cx<T> v[E];
#pragma unroll
      for (int e = 0; e < E; ++e) {
        cx<T> x = sm[e * nthreads + tid];
        x.y *= cj;
        v[e] = x;
      }
      if (ahead_tiles > 0 && tid < 32) bulk_store_wait_read1();
      __syncthreads();   // the slot now serves as exchange and staging buffer
      if (ahead_tiles > 0) {
        int ahead = slot + ahead_tiles;   // the slot of the tile before the previous one
        if (ahead >= depth) ahead -= depth;
        prefetch(tile + (unsigned)ahead_tiles * gridDim.x, ahead);
      }
      transform_tile(v, sm, tile);
      slot = slot + 1 == depth ? 0 : slot + 1;
    }
    if (tid < 32) bulk_store_wait_all();   // every run has reached its destination
  }
  cp_async_wait(0);
  // A launch that was allowed to start before its predecessor had drained must not FINISH before it either: whatever
  // follows the chain on the stream (an event, the next phase, the caller's own kernels) waits only for the last launch.
  if ((a.pdl & 2) && tid == 0) asm volatile("griddepcontrol.wait;" ::: "memory");
  if (a.signal_count > 0) {
    // every store of this CTA is ordered before the counter; the last CTA tells the peers
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

}  // namespace offtb
