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
};

#define OFFTB_MAX_RING 8     // ring slots a phase launch can address (window W <= 7; deeper rings use per-tile launches)

// Phase launch: ONE grid walks phase_tiles consecutive tiles of a phase (each of FftArgs::ntiles items) instead of
// one launch per tile.  Phase tile pt uses ring slot (slot0 + pt) % ring_depth, reads at in_slot[slot] (or FftArgs::in)
// + pt*in_step and writes at + pt*out_step (through tab[slot] when FftArgs::out_split).  The flags of FftArgs are those
// of phase tile 0 in flag row 0; tile pt waits for wait_value + pt in row slot*flag_stride and publishes
// signal_value + pt there.  Readers wait before loading a tile, writers before storing it.
struct PhaseArgs {
  int phase_tiles, ring_depth, slot0, ry_step, flag_stride, wait_at_load;
  long long in_step, out_step;
  unsigned *done;                       // [phase_tiles] zeroed device words, zero again after the launch
  const void *in_slot[OFFTB_MAX_RING];  // non-null: input base of that ring slot
  void *tab[OFFTB_MAX_RING][OFFTB_MAX_GROUP];
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
__device__ __forceinline__ void cp_async_cx(cx<double> *smem_dst, const cx<double> *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_cx(cx<float> *smem_dst, const cx<float> *gsrc) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
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

template <typename T, class CFG, int S>
__device__ __forceinline__ void fft_stage(cx<T> (&v)[CFG::E], const cx<T> *wreg, const FftArgs &a, cx<T> *sm,
                                          void *const *s_tab, long long bofs, long long ooff, unsigned bblock, int t, int s_mul, int s_base, T cj) {
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
    fft_stage<T, CFG, S + 1>(v, wreg, a, sm, s_tab, bofs, ooff, bblock, t, s_mul, s_base, cj);
  } else {
    // ---- last stage: output index beta + (N/R)*k
    if (a.load_cfast == a.store_cfast) {
#pragma unroll
      for (int u = 0; u < NU; ++u) {
        const int beta = t + TT * u;
#pragma unroll
        for (int pos = 0; pos < R; ++pos) {
          cx<T> e = v[u * R + pos];
          e.y *= cj;
          *out_ptr<T>(a, s_tab, bofs, beta + (N / R) * brev(pos, R)) = e;
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
        *out_ptr<T>(a, s_tab, map_b(a.om, bblock + cc) + ooff, k) = val;
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

template <typename T, class CFG>
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
  for (int j = tid; j < OFFTB_MAX_GROUP; j += nthreads)
    s_tab[j] = a.out_split ? a.out_tab[j] : (void *)((cx<T> *)a.out + (long long)j * a.om.n_hi);
  if (a.wait_count > 0) {
    // peers release the slots this launch writes (or fill the ones it reads) with a system-scope store
    for (int j = tid; j < a.wait_count; j += nthreads) {
      const volatile unsigned *f = a.wait_flags + j;
      const long long t0 = clock64();
      while ((int)(*f - a.wait_value) < 0) {
        __nanosleep(100);
        // a peer that never answers (a rank died, mismatched plans) must not hang the GPU: give up after ~10 s
        if (clock64() - t0 > 20000000000LL) __trap();
      }
    }
    __threadfence_system();
  }
  __syncthreads();

  constexpr int NW = CFG::twregs(CFG::NS - 1) > 0 ? CFG::twregs(CFG::NS - 1) : 1;
  cx<T> wreg[NW];
  load_twiddles<T, CFG, 0>(wreg, (const cx<T> *)a.tw, t);

  // this thread's E points of `tile` -> its private places in ring slot `slot`
  auto prefetch = [&](unsigned tile, int slot) {
    if (tile < a.ntiles) {
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

  unsigned tile = blockIdx.x;
  for (int d = 0; d + 1 < depth; ++d) prefetch(tile + d * gridDim.x, d);
  int slot = 0;
  for (; tile < a.ntiles; tile += gridDim.x) {
    cx<T> *sm = sm_all + slot * slot_elems;
    if (depth == 1) {
      __syncthreads();           // the previous tile's exchange data has been consumed
      prefetch(tile, 0);
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
      prefetch(tile + (unsigned)(depth - 1) * gridDim.x, ahead);
    }
    const unsigned bblock = tile << a.c_log;
    const long long bofs = map_b(a.om, bblock + c);

    // Ry rule: a tile-uniform choice between transforming and merely moving its columns
    bool transform = true;
    if (a.ry_level >= 0) {
      const int r = (a.ry_x0 + (int)digit_b(a.im, bblock, a.ry_level)) % 10;
      transform = a.ry_lo <= r && r < a.ry_hi;
    }
    if (transform) {
      fft_stage<T, CFG, 0>(v, wreg, a, sm, s_tab, bofs, 0LL, bblock, t, s_mul, s_base, cj);
    } else {
#pragma unroll
      for (int u = 0; u < NU0; ++u)
#pragma unroll
        for (int i = 0; i < R0; ++i) {
          cx<T> e = v[u * R0 + i];
          e.y *= cj;   // undo the conjugation of the load
          *out_ptr<T>(a, s_tab, bofs, t + TT * u + (N / R0) * i) = e;
        }
    }
    slot = slot + 1 == depth ? 0 : slot + 1;
  }
  cp_async_wait(0);
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

// ---- phase launch ------------------------------------------------------------------------------------------------

// all threads of the CTA: wait until `count` flags reach `want` (system-scope stores of the peers)
static __device__ __noinline__ void phase_wait(const volatile unsigned *flags, int count, unsigned want, int tid, int nthreads) {
  for (int j = tid; j < count; j += nthreads) {
    const long long t0 = clock64();
    while ((int)(flags[j] - want) < 0) {
      __nanosleep(100);
      if (clock64() - t0 > 20000000000LL) __trap();   // ~10 s: a peer died or the plans do not match
    }
  }
  __threadfence_system();
  __syncthreads();
}

// all threads of the CTA: `cnt` items of a phase tile of `ntiles` items are done; whoever completes the tile publishes it
static __device__ __noinline__ void phase_finish(unsigned *done_word, unsigned cnt, unsigned ntiles, unsigned *const *signal_ptrs, int signal_count,
                                          int row_off, unsigned value, int tid) {
  __threadfence_system();   // every store of this CTA is ordered before the counter
  __syncthreads();
  if (tid == 0) {
    const unsigned prev = atomicAdd(done_word, cnt);
    if (prev + cnt == ntiles) {
      *done_word = 0;
      __threadfence_system();
      for (int j = 0; j < signal_count; ++j) *(volatile unsigned *)(signal_ptrs[j] + row_off) = value;
    }
  }
}

template <typename T, class CFG>
__global__ void __launch_bounds__(CFG::MAXT, CFG::MINB) fft_phase_kernel(const __grid_constant__ FftArgs a, const __grid_constant__ PhaseArgs ph) {
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

  __shared__ void *s_tab[OFFTB_MAX_RING][OFFTB_MAX_GROUP];
  __shared__ const cx<T> *s_in[OFFTB_MAX_RING];
  for (int j = tid; j < ph.ring_depth * OFFTB_MAX_GROUP; j += nthreads) {
    const int rs = j / OFFTB_MAX_GROUP, g = j % OFFTB_MAX_GROUP;
    s_tab[rs][g] = a.out_split ? ph.tab[rs][g] : (void *)((cx<T> *)a.out + (long long)g * a.om.n_hi);
    if (g == 0) s_in[rs] = (const cx<T> *)(ph.in_slot[rs] ? ph.in_slot[rs] : a.in);
  }
  __syncthreads();

  constexpr int NW = CFG::twregs(CFG::NS - 1) > 0 ? CFG::twregs(CFG::NS - 1) : 1;
  cx<T> wreg[NW];
  load_twiddles<T, CFG, 0>(wreg, (const cx<T> *)a.tw, t);

  // a cursor over the work items of this CTA: (phase tile, tile in it) with the quantities that follow the phase tile
  struct Cursor { int pt, rs; unsigned tile; long long in_off, out_off; int ryx; };
  auto advance = [&](Cursor &q) {
    q.tile += gridDim.x;
    while (q.tile >= a.ntiles && q.pt < ph.phase_tiles) {
      q.tile -= a.ntiles; ++q.pt;
      q.rs = q.rs + 1 == ph.ring_depth ? 0 : q.rs + 1;
      q.in_off += ph.in_step; q.out_off += ph.out_step; q.ryx += ph.ry_step;
    }
  };
  Cursor cur = {0, ph.slot0 % ph.ring_depth, blockIdx.x, 0, 0, a.ry_x0};
  cur.tile -= gridDim.x;
  advance(cur);              // normalises a grid wider than one phase tile
  Cursor pf = cur;
  int seen_pt = -1;          // phase tiles this CTA has already waited for

  auto prefetch = [&](const Cursor &q, int slot) {
    if (q.pt < ph.phase_tiles) {
      if (ph.wait_at_load && q.pt > seen_pt) {   // reader: the tile must have arrived
        phase_wait(a.wait_flags + q.rs * ph.flag_stride, a.wait_count, a.wait_value + (unsigned)q.pt, tid, nthreads);
        seen_pt = q.pt;
      }
      const cx<T> *gin = s_in[q.rs] + (q.in_off + map_b(a.im, (q.tile << a.c_log) + c));
      cx<T> *dst = sm_all + slot * slot_elems + tid;
#pragma unroll
      for (int u = 0; u < NU0; ++u)
#pragma unroll
        for (int i = 0; i < R0; ++i)
          cp_async_cx(dst + (u * R0 + i) * nthreads, gin + map_n(a.im, t + TT * u + (N / R0) * i));
    }
    cp_async_commit();
  };

  for (int d = 0; d + 1 < depth; ++d) { prefetch(pf, d); advance(pf); }
  int slot = 0;
  unsigned cnt = 0;
  for (; cur.pt < ph.phase_tiles; advance(cur)) {
    cx<T> *sm = sm_all + slot * slot_elems;
    if (depth == 1) {
      __syncthreads();           // the previous tile's exchange data has been consumed
      prefetch(cur, 0);
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
    __syncthreads();             // the slot now serves as exchange buffer
    if (depth > 1) {
      const int ahead = slot == 0 ? depth - 1 : slot - 1;   // the slot the previous tile has just released
      prefetch(pf, ahead);
      advance(pf);
    }
    if (!ph.wait_at_load && cur.pt > seen_pt) {   // writer: the landing slots must have been released
      phase_wait(a.wait_flags + cur.rs * ph.flag_stride, a.wait_count, a.wait_value + (unsigned)cur.pt, tid, nthreads);
      seen_pt = cur.pt;
    }
    const unsigned bblock = cur.tile << a.c_log;
    const long long bofs = map_b(a.om, bblock + c) + cur.out_off;
    void *const *tab = s_tab[cur.rs];

    bool transform = true;
    if (a.ry_level >= 0) {
      const int r = (cur.ryx + (int)digit_b(a.im, bblock, a.ry_level)) % 10;
      transform = a.ry_lo <= r && r < a.ry_hi;
    }
    if (transform) {
      fft_stage<T, CFG, 0>(v, wreg, a, sm, tab, bofs, cur.out_off, bblock, t, s_mul, s_base, cj);
    } else {
#pragma unroll
      for (int u = 0; u < NU0; ++u)
#pragma unroll
        for (int i = 0; i < R0; ++i) {
          cx<T> e = v[u * R0 + i];
          e.y *= cj;   // undo the conjugation of the load
          *out_ptr<T>(a, tab, bofs, t + TT * u + (N / R0) * i) = e;
        }
    }
    ++cnt;
    if (cur.tile + gridDim.x >= a.ntiles) {   // this CTA's last item of the phase tile: the peers are waiting for it
      if (a.signal_count > 0)
        phase_finish(ph.done + cur.pt, cnt, a.ntiles, a.signal_ptrs, a.signal_count, cur.rs * ph.flag_stride, a.signal_value + (unsigned)cur.pt, tid);
      cnt = 0;
    }
    slot = slot + 1 == depth ? 0 : slot + 1;
  }
  cp_async_wait(0);
}

}  // namespace offtb
