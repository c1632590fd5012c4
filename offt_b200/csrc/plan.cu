// Plan construction and the execute schedules of the B200 3-D FFT.
//
// What the reference does per rank (offt-compute.c:3864-4048):
//     [FFTz + pack1] -> all-to-all in the row group -> [unpack1 + FFTy]      (phase 1, x tiles)
//     mid transpose
//     [FFTy + pack2] -> all-to-all in the column group -> [unpack2 + FFTx]   (phase 2, z tiles)
// with tile i's exchange overlapping the compute of tiles i-W .. i+W.
//
// Here every bracket is ONE launch of the batched 1-D kernel (fft_kernels.cuh) whose load and
// store maps do the (un)packing; z stays the contiguous axis of every intermediate array, so
// no transpose pass exists and the output layout the caller asked for (_S_, is_equalxy,
// offt-compute.c:282-313) is produced by the store map of the last launch.  The exchange is
// fused into the launches on either side of it: the packing launch stores each destination's
// block into that peer's ring slot over NVLink and flags in peer memory replace MPI_Wait
// (offt-compute.c:3607-3679, 3789-3861; "fused exchange" below).  Grouped ncclSend/ncclRecv
// between send and receive slots on a second stream, ordered by cudaEvents, remain as the
// fallback (OFFTB_EXCHANGE=nccl, or when the peers' rings cannot be mapped).
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>

#include "engine.h"

namespace offtb {

// ------------------------------------------------------------------------------ layout

namespace {

struct Share { int ceil, floor, extra, mine, start; };

// the last `extra` owners of a split hold one item more (offt-compute.c:141-144, 246-247)
Share share_of(int N, int owners, int r) {
  Share s;
  s.floor = N / owners;
  s.extra = N % owners;
  s.ceil = s.floor + (s.extra ? 1 : 0);
  const int small = owners - s.extra;
  s.mine = r < small ? s.floor : s.floor + 1;
  s.start = r < small ? r * s.floor : small * s.floor + (r - small) * (s.floor + 1);
  return s;
}

bool is_pow2(long long n) { return n > 0 && (n & (n - 1)) == 0; }
int lg2(long long n) { int l = 0; while ((1LL << l) < n) ++l; return l; }

}  // namespace

void comm_fill(struct _offt_comm *c, int Nx, int Ny, int Nz, int p, int p1, int rank, int S, int is_equalxy, int is_r2c) {
  const int p2 = p / p1;
  const int rx = rank / p2, ry = rank % p2;   // rank -> (row, column) of the process grid
  // real-to-complex plans keep Nz/2+1 complex points of every z row: that count replaces Nz in everything but the
  // input box's z extent (offt-compute.c:63, 130-143, 251)
  const int Nzc = is_r2c ? Nz / 2 + 1 : Nz;
  const Share x1 = share_of(Nx, p1, rx), y2 = share_of(Ny, p2, ry), z2 = share_of(Nzc, p2, ry), y1 = share_of(Ny, p1, rx);
  c->p1 = p1; c->p2 = p2;
  c->comm1 = c->comm2 = nullptr; c->group1 = c->group2 = nullptr;
  c->M1 = x1.ceil; c->M2 = y2.ceil; c->M3 = z2.ceil; c->M4 = y1.ceil;
  c->F1 = x1.floor; c->F2 = y2.floor; c->F3 = z2.floor; c->F4 = y1.floor;
  c->m1 = x1.mine; c->m2 = y2.mine; c->m3 = z2.mine; c->m4 = y1.mine;
  c->b1 = x1.extra; c->b2 = y2.extra; c->b3 = z2.extra; c->b4 = y1.extra;
  // input box: x-y-z, z contiguous, rows padded so that both phases fit the same array
  c->istart[0] = x1.start; c->istart[1] = y2.start; c->istart[2] = 0;
  c->isize[0] = x1.mine; c->isize[1] = y2.mine; c->isize[2] = Nz;
  const int yrows = std::max(c->M2 * p2, c->M4 * p1);
  c->istride[0] = yrows * c->M3; c->istride[1] = c->M3 * p2; c->istride[2] = 1;
  // output box: all of x, this rank's y block of the column split and z block of the row split
  c->ostart[0] = 0; c->ostart[1] = y1.start; c->ostart[2] = z2.start;
  c->osize[0] = Nx; c->osize[1] = y1.mine; c->osize[2] = z2.mine;
  const int xrow = c->M1 * p1;
  if (S) { c->ostride[0] = c->M3 * c->M4; c->ostride[1] = c->M3; c->ostride[2] = 1; }                      // x-y-z
  else if (is_equalxy && c->M1 == c->M4) { c->ostride[0] = 1; c->ostride[1] = xrow * c->M3; c->ostride[2] = xrow; }  // y-z-x
  else { c->ostride[0] = 1; c->ostride[1] = xrow; c->ostride[2] = xrow * c->M4; }                          // z-y-x
}

long long alloc_elems(int Nx, int Ny, int Nz, int p, int p1, int is_r2c) {
  const int p2 = p / p1;
  auto cd = [](long long a, long long b) { return (a + b - 1) / b; };
  const long long M1 = cd(Nx, p1), M2 = cd(Ny, p2), M3 = cd(is_r2c ? Nz / 2 + 1 : Nz, p2), M4 = cd(Ny, p1);   // run-fft.c:296-300
  return std::max(M2 * p2, M4 * p1) * M1 * M3;
}

int check_supported(int Nx, int Ny, int Nz, int p, int p1, int is_r2c) {
  if (p < 1 || p1 < 1 || p % p1 != 0) { set_error("process grid %d = %d x ? is not a grid", p, p1); return -2; }
  const int p2 = p / p1;
  if (is_r2c && (size_t)Nz > fft_generic_max_n(PREC_F64)) {
    set_error("real-to-complex plans transform z in shared memory: Nz = %d exceeds %zu", Nz, fft_generic_max_n(PREC_F64));
    return -3;
  }
  FftKernelInfo info;
  for (int n : {Nx, Ny, Nz}) {
    if (n < 1) { set_error("transform length %d", n); return -3; }
    // powers of two up to 8192 run on the register-butterfly kernels, everything else on the generic one
    if (!(is_pow2(n) && fft_kernel_info(n, PREC_F64, &info)) && (size_t)n > fft_generic_max_n(PREC_F64)) {
      set_error("transform length %d: lengths that are not a power of two are limited to %zu (they run in shared memory)", n, fft_generic_max_n(PREC_F64));
      return -3;
    }
  }
  if (p1 > OFFTB_MAX_GROUP || p2 > OFFTB_MAX_GROUP) {
    set_error("process grid %dx%d: exchange groups of more than %d ranks are not supported", p1, p2, OFFTB_MAX_GROUP);
    return -5;
  }
  // the reference's own range of P1 (params_range_setup, offt-compute.c:3005-3012): every rank owns at least one plane of
  // each split - max(p/Nz, p/Ny, 1) <= p1 <= min(Nx, Ny, p)
  if (p1 > Nx || p1 > Ny || p2 > Ny || p2 > (is_r2c ? Nz / 2 + 1 : Nz)) {
    set_error("grid %dx%dx%d cannot be split over %dx%d ranks (P1 outside the reference's range)", Nx, Ny, Nz, p1, p2);
    return -4;
  }
  return 0;
}

// ------------------------------------------------------------------------------ engine

namespace {

int make_twiddles(int N, int prec, void **dev) {
  std::vector<long double> tab(2 * (size_t)N + 2);
  const int count = fft_twiddle_table(N, prec, tab.data());
  if (count < 0) { *dev = nullptr; return 0; }   // not a length of the power-of-two kernels: the generic kernel has its own table
  const size_t n = (size_t)std::max(count, 1);
  std::vector<double> hd(2 * n);
  std::vector<float> hf(2 * n);
  for (size_t j = 0; j < 2 * (size_t)count; ++j) { hd[j] = (double)tab[j]; hf[j] = (float)tab[j]; }
  const size_t bytes = n * (prec == PREC_F64 ? 16 : 8);
  OFFTB_CUDA(cudaMalloc(dev, bytes));
  OFFTB_CUDA(cudaMemcpy(*dev, prec == PREC_F64 ? (void *)hd.data() : (void *)hf.data(), bytes, cudaMemcpyHostToDevice));
  return 0;
}

// full table exp(-2*pi*i*k/N) for the generic kernel
int make_twiddles_full(int N, int prec, void **dev) {
  *dev = nullptr;
  if ((size_t)N > fft_generic_max_n(prec)) return 0;   // a length only the power-of-two kernels take
  std::vector<long double> tab(2 * (size_t)N);
  fft_generic_twiddle_table(N, tab.data());
  std::vector<double> hd(2 * (size_t)N);
  std::vector<float> hf(2 * (size_t)N);
  for (size_t j = 0; j < 2 * (size_t)N; ++j) { hd[j] = (double)tab[j]; hf[j] = (float)tab[j]; }
  const size_t bytes = (size_t)N * (prec == PREC_F64 ? 16 : 8);
  OFFTB_CUDA(cudaMalloc(dev, bytes));
  OFFTB_CUDA(cudaMemcpy(*dev, prec == PREC_F64 ? (void *)hd.data() : (void *)hf.data(), bytes, cudaMemcpyHostToDevice));
  return 0;
}

FftMap mk_map(long long off, long long nlo_count, long long n_hi, long long n_lo, long long B0, long long s0,
              long long B1, long long s1, long long s2) {
  FftMap m;
  m.off = off;
  m.n_lg = nlo_count > 0 ? lg2(nlo_count) : 30;   // 30: no split, n_hi unused
  m.n_hi = n_hi; m.n_lo = n_lo;
  m.B0 = (unsigned)std::max<long long>(B0, 1); m.B1 = (unsigned)std::max<long long>(B1, 1);
  m.s0 = s0; m.s1 = s1; m.s2 = s2;
  m.gF = m.gb = m.gg = 0;
  return m;
}

// The transform index of `m` is divided among `owners` blocks the way the reference divides an axis among ranks
// (offt-compute.c:128-144): floor(N/owners) items each, the last N % owners blocks one more.  Even divisions into
// power-of-two blocks keep the shift/mask split the fast kernels use; everything else takes the general split.
FftMap split_over(FftMap m, long long N, long long owners) {
  const long long F = N / owners, b = N % owners, M = F + (b ? 1 : 0);
  if (b == 0 && is_pow2(M)) { m.n_lg = lg2(M); m.gg = 0; }
  else { m.n_lg = 30; m.gF = (int)F; m.gb = (int)b; m.gg = (int)owners; }
  return m;
}

struct Launch {
  int N;
  int axis;          // 0 x, 1 y, 2 z -> twiddle table
  const void *in;
  void *out;
  FftMap im, om;
  long long nbatch;
  bool load_cfast, store_cfast;
  int ry_level = -1, ry_x0 = 0, ry_lo = 0, ry_hi = 10;
  // fused exchange (filled by fuse_writer / fuse_reader)
  bool split = false;                 // the slot side of the launch is a table of peers' slots
  void *tab[OFFTB_MAX_GROUP] = {nullptr};
  const unsigned *wait_flags = nullptr;
  int wait_count = 0;
  unsigned wait_value = 0;
  unsigned *signal_ptrs[OFFTB_MAX_GROUP] = {nullptr};
  int signal_count = 0;
  unsigned signal_value = 0;
  unsigned *done_counter = nullptr;
  int grid_cap = 0;
  const void *tw = nullptr;   // twiddle table other than the axis' own (the half-length transform of a real z pass)
  int depth = 0; // > 0: ring slots of the launch (0: the launcher decides)
  int r2c = 0;   // z pass of a real-to-complex plan
  int pdl = 0;   // bit 0: let the next launch of the stream start early; bit 1: this launch may itself start early
};

cudaEvent_t pool_event(Engine &E) {
  if (E.event_next == E.event_pool.size()) {
    cudaEvent_t ev;
    cudaEventCreate(&ev);
    E.event_pool.push_back(ev);
  }
  return E.event_pool[E.event_next++];
}

int pick_c_log(const Engine &E, const FftKernelInfo &info, const Launch &L) {
  int lg = fft_pick_c_log(info, E.prec, L.load_cfast || L.store_cfast, L.im.B0, L.nbatch, L.im.n_lo, L.om.n_lo);
  // narrow tiles of a phase whose full-width tiles leave no room for writer and reader side by side (plan_phase_overlap)
  if (E.narrow_now && (L.load_cfast || L.store_cfast)) {
    const int cap = E.prec == PREC_F64 ? 1 : 2;   // 32 bytes per transform index
    lg = std::min(lg, cap);
  }
  return lg;
}

// Can the launch drain its output with TMA bulk stores (fft_kernels.cuh, FftArgs::bulk_store)?  Only launches that
// scatter into peers' slots want it, and block a of the tile's output must be ONE run in memory laid out like the
// kernel's shared-memory slot: [k_lo][column] with the CTA's columns adjacent (strided launches, K3's z-chunked
// slots), or one run of k_lo per column (contiguous rows, K1).  Runs and their addresses must be 16-byte multiples.
bool bulk_store_ok(const Engine &E, const FftArgs &a, int N) {
  // Opt-in (OFFTB_BULK=1).  Measured on 2 GPUs, 1024^3 (profiles/r02_exchange_ab.md): the writer chain alone moves
  // 583 GB/s per direction with bulk stores against 607 with direct stores - the per-SM rate of remote writes is the
  // limit either way - and the three-slot ring leaves one CTA per SM, which forbids the two-stream overlap with the
  // reader (15.2 ms against 11.7).  Emulated-rank worlds always take it so that the parity tests keep covering it.
  static const int env_bulk = getenv("OFFTB_BULK") ? atoi(getenv("OFFTB_BULK")) : -1;
  const bool env_on = env_bulk >= 0 ? env_bulk != 0 : world().local;
  if (!env_on || !a.out_split || a.om.n_lg >= 30) return false;
  const long long C = 1LL << a.c_log, nlo = 1LL << a.om.n_lg;
  if (nlo > N) return false;
  long long run;
  if (a.load_cfast && a.store_cfast) {
    if (a.om.s0 != 1 || (long long)a.om.B0 != C || a.om.n_lo != C) return false;
    run = nlo * C;
  } else if (!a.load_cfast && !a.store_cfast) {
    if (a.om.n_lo != 1) return false;
    run = nlo;
  } else {
    return false;
  }
  const long long per16 = 16 / (long long)E.esz;   // elements per 16 bytes: 1 (complex128) or 2 (complex64)
  if (run % per16) return false;
  if (per16 > 1) {
    const bool rows = !a.store_cfast;
    if (a.om.off % per16 || a.om.s1 % per16 || a.om.s2 % per16 || (rows && a.om.s0 % per16)) return false;
    // batch digits whose extent is 1 never contribute their stride
  }
  for (int j = 0; j < (N >> a.om.n_lg); ++j)
    if (!a.out_tab[j] || ((uintptr_t)a.out_tab[j] & 15)) return false;
  return true;
}

int run_launch(Engine &E, cudaStream_t st, int stage, Launch L, bool inverse) {
  // a launch with nothing to transform still has to take part in the flag protocol of its tile
  if (L.nbatch <= 0 && L.signal_count == 0) return 0;
  if (L.nbatch < 0) L.nbatch = 0;
  FftKernelInfo info;
  const bool generic = g_force_generic > 0 || L.nbatch == 0 || L.r2c || L.im.gg > 0 || L.om.gg > 0 || !fft_kernel_info(L.N, E.prec, &info);
  if (L.nbatch >= (1LL << 32)) { set_error("batch of %lld rows exceeds the 32-bit batch index", L.nbatch); return -1; }
  // Strided launches take their columns from the fastest batch digit, which must hold a whole number of tiles.  An
  // extent that does not (the Nz/2+1 = 257 complex points of a real-to-complex row, an odd local share) would fall
  // back to one column - 16-byte accesses - for the whole launch; instead the launch is split into the part that
  // tiles at full width and the remainder.  Not for launches that take part in a flag protocol (one signal per tile).
  if (!generic && !E.dry_shape && (L.load_cfast || L.store_cfast) && L.signal_count == 0 && L.wait_count == 0 && L.im.B0 == L.om.B0) {
    const unsigned want = (unsigned)(64 / E.esz), B0 = L.im.B0;
    if (B0 > want && B0 % want != 0) {
      const unsigned main = B0 / want * want, rem = B0 - main;
      Launch A = L, B = L;
      A.im.B0 = A.om.B0 = main; A.nbatch = L.nbatch / B0 * main;
      B.im.B0 = B.om.B0 = rem; B.nbatch = L.nbatch / B0 * rem;
      B.im.off += (long long)main * L.im.s0; B.om.off += (long long)main * L.om.s0;
      if (run_launch(E, st, stage, A, inverse)) return -1;
      return run_launch(E, st, stage, B, inverse);
    }
  }
  FftArgs a;
  memset(&a, 0, sizeof(a));
  if (inverse) {
    std::swap(L.im, L.om);
    std::swap(L.load_cfast, L.store_cfast);
    const void *t = L.in; L.in = L.out; L.out = const_cast<void *>(t);
  }
  a.in = L.in; a.out = L.out; a.tw = L.tw ? L.tw : generic ? E.tw_full[L.axis] : E.tw[L.axis];
  a.im = L.im; a.om = L.om;
  a.load_cfast = L.load_cfast; a.store_cfast = L.store_cfast;
  a.conj = inverse ? 1 : 0;
  a.ry_level = L.ry_level; a.ry_x0 = L.ry_x0; a.ry_lo = L.ry_lo; a.ry_hi = L.ry_hi;
  a.out_split = L.split ? 1 : 0;
  for (int j = 0; j < OFFTB_MAX_GROUP; ++j) { a.out_tab[j] = L.tab[j]; a.signal_ptrs[j] = L.signal_ptrs[j]; }
  a.wait_flags = L.wait_flags; a.wait_count = L.wait_count; a.wait_value = L.wait_value;
  a.signal_count = L.signal_count; a.signal_value = L.signal_value; a.done_counter = L.done_counter;
  a.grid_cap = L.grid_cap;
  a.wait_timeout_ns = E.wait_timeout_ns; a.error_word = E.d_error;
  a.real_mode = L.r2c ? (inverse ? 2 : 1) : 0;
  if (generic) {
    // any length, uneven splits (fft_generic.cu)
    if (!a.tw) { set_error("length %d with this split needs the generic kernel, which holds at most %zu points", L.N, fft_generic_max_n(E.prec)); return -1; }
    if (E.dry_shape) {
      cudaError_t se = fft_generic_launch(L.N, E.prec, a, L.nbatch, st, E.dry_shape);
      if (se != cudaSuccess) { set_error("generic kernel shape (N=%d, batch=%lld): %s", L.N, L.nbatch, cudaGetErrorString(se)); return -1; }
      return 0;
    }
    cudaEvent_t g0 = nullptr, g1 = nullptr;
    const bool gtimed = E.stage_timing && !E.async && !E.chain_timing;
    if (gtimed) { g0 = pool_event(E); g1 = pool_event(E); cudaEventRecord(g0, st); }
    cudaError_t ge = fft_generic_launch(L.N, E.prec, a, L.nbatch, st, nullptr);
    if (ge != cudaSuccess) { set_error("generic kernel launch (N=%d, batch=%lld): %s", L.N, L.nbatch, cudaGetErrorString(ge)); return -1; }
    if (gtimed) { cudaEventRecord(g1, st); E.timed.push_back({stage, {g0, g1}}); }
    E.launches++;
    return 0;
  }
  a.pdl = L.pdl;
  a.depth = L.depth;
  a.c_log = pick_c_log(E, info, L);
  a.bulk_store = bulk_store_ok(E, a, L.N) ? 1 : 0;
  if (a.ry_level >= 0 && a.load_cfast != a.store_cfast) { set_error("internal: Ry rule on a transposing launch"); return -1; }
  if (E.dry_shape) {
    cudaError_t se = fft_shape(L.N, E.prec, a, L.nbatch, E.dry_shape);
    if (se != cudaSuccess) { set_error("kernel shape (N=%d, batch=%lld): %s", L.N, L.nbatch, cudaGetErrorString(se)); return -1; }
    return 0;
  }
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = E.stage_timing && !E.async && !E.chain_timing;
  if (timed) { e0 = pool_event(E); e1 = pool_event(E); cudaEventRecord(e0, st); }
  cudaError_t err = fft_launch(L.N, E.prec, a, L.nbatch, st);
  if (err != cudaSuccess) { set_error("kernel launch (N=%d, batch=%lld): %s", L.N, L.nbatch, cudaGetErrorString(err)); return -1; }
  if (timed) { cudaEventRecord(e1, st); E.timed.push_back({stage, {e0, e1}}); }
  E.launches++;
  return 0;
}

struct Dims {
  long long Nzc;   // complex points per z row: Nz, or Nz/2+1 in real-to-complex plans
  int r2c;
  long long Nx, Ny, Nz, p1, p2, M1, M2, M3, M4, m1, m2, m3, m4, isx, isy, dX, dY, os0, os1, os2;
  int T1, T2, W1, W2, Ry, S;
};

Dims dims_of(const struct _offt_plan *po) {
  const struct _offt_comm *c = po->comm;
  Dims d;
  d.Nx = po->Nx; d.Ny = po->Ny; d.Nz = po->Nz; d.p1 = c->p1; d.p2 = c->p2;
  d.r2c = po->is_r2c ? 1 : 0;
  d.Nzc = d.r2c ? d.Nz / 2 + 1 : d.Nz;
  d.M1 = c->M1; d.M2 = c->M2; d.M3 = c->M3; d.M4 = c->M4;
  d.m1 = c->m1; d.m2 = c->m2; d.m3 = c->m3; d.m4 = c->m4;
  d.isx = c->istride[0]; d.isy = c->istride[1];
  d.dY = d.M3; d.dX = d.M3 * d.M4 * d.p1;           // the x-y-z_local array between the phases
  d.os0 = c->ostride[0]; d.os1 = c->ostride[1]; d.os2 = c->ostride[2];
  const int *v = po->params->v;
  d.T1 = v[_T1_]; d.T2 = v[_T2_]; d.W1 = v[_W1_]; d.W2 = v[_W2_]; d.Ry = v[_Ry_]; d.S = v[_S_];
  return d;
}

inline char *at(void *base, long long elems, size_t esz) { return (char *)base + (size_t)elems * esz; }

// ---- the launches -----------------------------------------------------------------------

// FFTz over x planes [x0, x0+nx) of the input box, rows contiguous
Launch L_fftz_local(const Dims &d, const void *in, void *out, long long x0, long long nx) {
  Launch L;
  L.N = (int)d.Nz; L.axis = 2; L.in = in; L.out = out;
  L.im = L.om = mk_map(x0 * d.isx, 0, 0, 1, d.m2, d.isy, nx, d.isx, 0);
  L.nbatch = d.m2 * nx; L.load_cfast = L.store_cfast = false;
  L.r2c = d.r2c;   // real rows of Nz points in, Nz/2+1 complex points out (offt-compute.c:960-961, 3973-3974)
  return L;
}

// K1: FFTz + pack1 (reference compute_fftz_pack1, offt-compute.c:905-1206, buffer layout of :1015-1032)
Launch L_k1(const Dims &d, const void *U, void *send, long long x0, long long myT) {
  Launch L = L_fftz_local(d, U, send, x0, myT);
  // z splits into (destination, z_local): block a at a*myT*M2*M3, inside it [x][y][z_local]
  L.om = split_over(mk_map(0, d.M3, myT * d.M2 * d.M3, 1, d.m2, d.M3, myT, d.M2 * d.M3, 0), d.Nzc, d.p2);
  return L;
}

// K2: unpack1 + FFTy (compute_unpack1_ffty, offt-compute.c:1208-1520; addresses of :1307-1311)
Launch L_k2(const Dims &d, const void *recv, void *A, long long x0, long long myT) {
  Launch L;
  L.N = (int)d.Ny; L.axis = 1; L.in = recv; L.out = A;
  // y splits into (source, y_local): block a holds [x][y_local][z]
  L.im = split_over(mk_map(0, d.M2, myT * d.M2 * d.M3, d.M3, d.m3, 1, myT, d.M2 * d.M3, 0), d.Ny, d.p2);
  L.om = mk_map(x0 * d.dX, 0, 0, d.dY, d.m3, 1, myT, d.dX, 0);
  L.nbatch = d.m3 * myT; L.load_cfast = L.store_cfast = true;
  return L;
}

// z chunk of the phase-2 slots: the tile's z planes are split z = z_hi*Cz + z_lo and a slot block is laid out
// [x][z_hi][y_local][z_lo].  One CTA of K3 (Cz columns of z, all y of one x) then writes, per destination, ONE
// contiguous run of M4*Cz elements with consecutive lanes on consecutive addresses - 512-byte stores that cross
// NVLink efficiently - instead of 64-byte pieces (the reference's [x][y_local][z] order, offt-compute.c:1773-1776,
// measured 390 GB/s per direction in the fused exchange).  K4 reads Cz-element rows.  The layout is internal to
// the ring; what the caller sees (ostride) is unchanged.
long long z_chunk(const Engine &E, long long myT) {
  static const long long env_cz = getenv("OFFTB_CZ") ? atoll(getenv("OFFTB_CZ")) : 0;   // experiments
  long long cz = env_cz > 0 ? env_cz : (E.prec == PREC_F64 ? 4 : 8);
  if (E.narrow_now) cz /= 2;   // the chunk is what one CTA writes as a run: keep it equal to the tile's columns
  while (cz > 1 && myT % cz) cz /= 2;
  return cz;
}

// K3: FFTy + pack2 (compute_ffty_pack2, offt-compute.c:1636-2345)
Launch L_k3(const Engine &E, const Dims &d, const void *A, void *send, long long z0, long long myT) {
  Launch L;
  const long long cz = z_chunk(E, myT);
  L.N = (int)d.Ny; L.axis = 1; L.in = A; L.out = send;
  L.im = mk_map(z0, 0, 0, d.dY, cz, 1, myT / cz, cz, d.dX);
  // y splits into (destination, y_local): block a holds [x][z_hi][y_local][z_lo]
  L.om = split_over(mk_map(0, d.M4, d.M1 * d.M4 * myT, cz, cz, 1, myT / cz, d.M4 * cz, myT * d.M4), d.Ny, d.p1);
  L.nbatch = myT * d.m1; L.load_cfast = L.store_cfast = true;
  return L;
}

// K4: unpack2 + FFTx (compute_unpack2_fftx, offt-compute.c:2347-2993; addresses :2447-2450, 2567-2570, 2684-2687)
Launch L_k4(const Engine &E, const Dims &d, const void *recv, void *U, long long z0, long long myT) {
  Launch L;
  const long long cz = z_chunk(E, myT);
  L.N = (int)d.Nx; L.axis = 0; L.in = recv; L.out = U;
  // x splits into (source, x_local): block a holds [x_local][z_hi][y][z_lo]; batch digits (z_lo, y, z_hi)
  L.im = split_over(mk_map(0, d.M1, d.M1 * d.M4 * myT, myT * d.M4, cz, 1, d.m4, cz, d.M4 * cz), d.Nx, d.p1);
  L.om = mk_map(z0 * d.os2, 0, 0, d.os0, cz, d.os2, d.m4, d.os1, cz * d.os2);
  L.nbatch = myT * d.m4; L.load_cfast = true; L.store_cfast = (d.os2 == 1);
  return L;
}

// whole-array FFTy in the x-y-z_local array (single rank)
Launch L_ffty_local(const Dims &d, void *A) {
  Launch L;
  L.N = (int)d.Ny; L.axis = 1; L.in = A; L.out = A;
  L.im = L.om = mk_map(0, 0, 0, d.dY, d.m3, 1, d.m1, d.dX, 0);
  L.nbatch = d.m3 * d.m1; L.load_cfast = L.store_cfast = true;
  return L;
}

// whole-array FFTx from the x-y-z_local array into the output layout (slab 1 x p tail,
// offt-compute.c:3913-3949, and the single-rank case)
Launch L_fftx_local(const Dims &d, const void *A, void *U) {
  Launch L;
  L.N = (int)d.Nx; L.axis = 0; L.in = A; L.out = U;
  L.im = mk_map(0, 0, 0, d.dX, d.m3, 1, d.Ny, d.dY, 0);
  L.om = mk_map(0, 0, 0, d.os0, d.m3, d.os2, d.Ny, d.os1, 0);
  L.nbatch = d.m3 * d.Ny; L.load_cfast = true; L.store_cfast = (d.os2 == 1);
  return L;
}

// Single rank, x-y-z output (_S_ = 1).  In place, the x pass strides by a whole y-z plane on BOTH sides and every
// 128-byte piece it touches sits in a page of its own (512^3: 4.5 TB/s).  With one side at the row stride the same
// pass runs at 5.2 TB/s (tools/kbench.py modes xs / sx), so the z pass - whose stores are whole rows and do not care
// where a row goes - files row (x, y) of its output as row (y, x) of a scratch array, the x pass reads that at the
// row stride and stores into the caller's x-y-z layout, and the y pass finishes in place at the row stride.
Launch L_fftz_swap(const Dims &d, const void *U, void *A) {
  Launch L = L_fftz_local(d, U, A, 0, d.m1);
  L.om = mk_map(0, 0, 0, 1, d.m2, d.Nx * d.M3, d.m1, d.M3, 0);          // A is [y][x][z]
  return L;
}
Launch L_fftx_swap(const Dims &d, const void *A, void *U) {
  Launch L;
  L.N = (int)d.Nx; L.axis = 0; L.in = A; L.out = U;
  L.im = mk_map(0, 0, 0, d.M3, d.m3, 1, d.Ny, d.Nx * d.M3, 0);          // rows along x at the row stride
  L.om = mk_map(0, 0, 0, d.os0, d.m3, d.os2, d.Ny, d.os1, 0);
  L.nbatch = d.m3 * d.Ny; L.load_cfast = L.store_cfast = true;
  return L;
}
// y pass in the caller's x-y-z array
Launch L_ffty_out(const Dims &d, void *U) {
  Launch L;
  L.N = (int)d.Ny; L.axis = 1; L.in = U; L.out = U;
  L.im = L.om = mk_map(0, 0, 0, d.os1, d.m3, d.os2, d.Nx, d.os0, 0);
  L.nbatch = d.m3 * d.Nx; L.load_cfast = L.store_cfast = true;
  return L;
}

// ---- exchanges ----------------------------------------------------------------------------

// members of this plan's row (phase 1) or column (phase 2) group and its own index in it
// (comm1 / comm2 of the reference, offt-compute.c:78-125)
void group_of(const Engine &E, int phase, std::vector<int> &members, int &me) {
  const struct _offt_comm *c = E.po->comm;
  members.clear();
  if (phase == 1) { for (int j = 0; j < c->p2; ++j) members.push_back(E.rank_x * c->p2 + j); me = E.rank_y; }
  else { for (int i = 0; i < c->p1; ++i) members.push_back(i * c->p2 + E.rank_y); me = E.rank_x; }
}

long long block_elems(const Dims &d, int phase, long long myT) {
  return phase == 1 ? myT * d.M2 * d.M3 : d.M1 * d.M4 * myT;   // offt-compute.c:3523, 3704
}

// One tile's all-to-all.  `from`/`to` are ring slots: forward sends `send` -> peers' `recv`,
// the backward transform runs the same exchange from `recv` to `send`.
int exchange(std::vector<Engine *> &engs, int phase, int slot, const std::vector<long long> &tile_T, bool inverse, cudaStream_t st) {
  World &w = world();
  for (size_t ke = 0; ke < engs.size(); ++ke) {
    Engine &E = *engs[ke];
    const long long myT = tile_T[ke];
    if (myT <= 0) continue;   // this rank's group has no such tile (uneven division: groups differ in their plane counts)
    const Dims d = dims_of(E.po);
    const long long blk = block_elems(d, phase, myT);
    std::vector<int> members;
    int me;
    group_of(E, phase, members, me);
    Ring &R = E.ring[phase - 1];
    char *src = (char *)(inverse ? R.recv[slot] : R.send[slot]);
    char *dst = (char *)(inverse ? R.send[slot] : R.recv[slot]);
    const size_t bytes = (size_t)blk * E.esz;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    const bool xtimed = E.stage_timing && !E.async;
    if (xtimed) { e0 = pool_event(E); e1 = pool_event(E); cudaEventRecord(e0, st); }
    if (w.local) {
      // every emulated rank pulls its blocks out of its peers' slots
      for (size_t j = 0; j < members.size(); ++j) {
        Engine *peer = nullptr;
        for (Engine *Q : engs) if (Q->po->rank == members[j]) peer = Q;
        if (!peer) { set_error("local world: plan of rank %d missing from the group", members[j]); return -1; }
        Ring &PR = peer->ring[phase - 1];
        const char *psrc = (const char *)(inverse ? PR.recv[slot] : PR.send[slot]);
        OFFTB_CUDA(cudaMemcpyAsync(dst + j * bytes, psrc + (size_t)me * bytes, bytes, cudaMemcpyDeviceToDevice, st));
      }
    } else if (members.size() == 1) {
      OFFTB_CUDA(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, st));
    } else {
      const ncclDataType_t ty = E.prec == PREC_F64 ? ncclDouble : ncclFloat;
      const NcclApi *nc = nccl_api();
      if (!nc) return -1;
      OFFTB_NCCL(nc->GroupStart());
      for (size_t j = 0; j < members.size(); ++j) {
        if ((int)j == me) continue;
        OFFTB_NCCL(nc->Send(src + j * bytes, (size_t)blk * 2, ty, members[j], w.nccl, st));
        OFFTB_NCCL(nc->Recv(dst + j * bytes, (size_t)blk * 2, ty, members[j], w.nccl, st));
      }
      OFFTB_NCCL(nc->GroupEnd());
      OFFTB_CUDA(cudaMemcpyAsync(dst + (size_t)me * bytes, src + (size_t)me * bytes, bytes, cudaMemcpyDeviceToDevice, st));
    }
    if (xtimed) { cudaEventRecord(e1, st); E.timed.push_back({phase == 1 ? ST_X1 : ST_X2, {e0, e1}}); }
  }
  return 0;
}

// ---- fused exchange ---------------------------------------------------------------------------
// The writer launch of a tile (K1/K3 forward, the inverses of K2/K4 backward) stores block a of its output
// straight into group member a's landing slot - over NVLink for the other GPUs of the box - at block index
// `me`; the reader launch finds the tile complete in its own landing slot.  Landing = receive slots forward,
// send slots backward (the backward exchange runs recv -> send, see exchange()).  Flags replace MPI_Wait:
// tile number s = tiles_done + i + 1 of this phase; a writer waits until every member has released tile
// s - depth (the previous tenant of the slot) and announces s; a reader waits for s from every member and
// releases s.  Local worlds need no flags: their launches are already ordered on one stream.

// Writer and reader launches of a phase can run on two streams, ordered only by the flags, so that the writer's
// NVLink stores overlap the reader's HBM passes.  A reader that filled the SMs would spin on its flags while the
// writer it waits for could not start, so both grids are capped: the SMs offer `slots` CTA places that fit either
// kernel (sized for the larger of the two), the writer gets a share of them and the reader the rest, and whatever
// the placement the writer always finds room.  Fewer than two places per SM (2048-point strided tiles take
// 131 KB and the whole register file): one stream, launch order - splitting the SMs between the two kernels
// instead was measured and loses, because the rate of remote stores scales with the number of SMs issuing
// them (2 GPUs, 64x2048x2048: 4.33 ms on one stream, 7.56 / 5.40 ms with 25 % / 40 % of the SMs for the writer).
// OFFTB_OVERLAP=0 forces one stream; OFFTB_WRITER_SHARE sets the writer's percentage of the places (default 50:
// 1024^3 on 8 GPUs 4.39 ms at 50, 4.58 at 30, 5.24 on one stream).
bool overlap_wanted() {
  static const int v = getenv("OFFTB_OVERLAP") ? atoi(getenv("OFFTB_OVERLAP")) : 1;
  return v != 0;
}

bool plan_overlap(Engine &E, const FftShape &w, const FftShape &r) {
  E.grid_cap[0] = E.grid_cap[1] = 0;
  if (!overlap_wanted() || w.grid == 0 || r.grid == 0) return false;
  int dev = 0, smem_sm = 0, regs_sm = 0, thr_sm = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev);
  cudaDeviceGetAttribute(&thr_sm, cudaDevAttrMaxThreadsPerMultiProcessor, dev);
  const int threads = std::max(w.threads, r.threads);
  const int regs = (std::max(w.regs, r.regs) + 7) / 8 * 8;
  const size_t smem = std::max(w.smem, r.smem) + 1024 + 256;   // per-CTA reservation and the kernel's static shared memory
  const int warps = (threads + 31) / 32;
  long long per_sm = std::min<long long>({(long long)regs_sm / ((long long)regs * 32 * warps), (long long)(smem_sm / smem),
                                          (long long)thr_sm / threads, 32LL});
  if (per_sm < 2) return false;
  const long long slots = per_sm * w.sm_count;
  static const int share = getenv("OFFTB_WRITER_SHARE") ? atoi(getenv("OFFTB_WRITER_SHARE")) : 50;
  long long gw = std::max<long long>(1, std::min<long long>(slots - 1, slots * std::min(std::max(share, 1), 99) / 100));
  // A writer whose CTAs are smaller than the places above (sized for the larger kernel: contiguous-row K1 beside a
  // strided reader) may fill what the reader's share leaves of an SM with more of them: the exchange-bound kernel wants
  // every store-issuing warp it can get.  Only the reader's cap is what keeps the SMs from filling with waiting CTAs;
  // writers never wait for a later tile, so they drain whatever their number.  OFFTB_WRITER_PACK=0 keeps equal places.
  static const bool pack = !(getenv("OFFTB_WRITER_PACK") && atoi(getenv("OFFTB_WRITER_PACK")) == 0);
  const long long gr = slots - gw;
  if (pack) {
    const long long nr = (gr + w.sm_count - 1) / w.sm_count;   // readers per SM when they spread evenly
    const int wregs = (w.regs + 7) / 8 * 8, rregs = (r.regs + 7) / 8 * 8;
    const int wwarps = (w.threads + 31) / 32, rwarps = (r.threads + 31) / 32;
    const size_t wsmem = w.smem + 1024 + 256, rsmem = r.smem + 1024 + 256;
    const long long by_regs = ((long long)regs_sm - nr * rregs * 32LL * rwarps) / ((long long)wregs * 32 * wwarps);
    const long long by_smem = ((long long)smem_sm - nr * (long long)rsmem) / (long long)wsmem;
    const long long by_thr = ((long long)thr_sm - nr * r.threads) / w.threads;
    const long long nw = std::min({by_regs, by_smem, by_thr, 32LL});
    if (nw * w.sm_count > gw) gw = nw * w.sm_count;
  }
  E.grid_cap[0] = (int)gw;
  E.grid_cap[1] = (int)gr;
  // Side by side each kernel holds about half the CTAs it would hold alone, and the reader - HBM-bound, long strided
  // tiles, a one-slot ring when it sizes itself for an SM of its own - becomes the longer of the two chains (1024^3:
  // 8.5 ms against 7.9 on 2 GPUs, 5.3 against 4.9 on 4).  The shared memory its missing twin CTAs would have taken is
  // free, so its ring can be deepened as far as the SM's shared memory allows next to the writer's CTAs
  // (OFFTB_READER_DEPTH=2 or 3).  Off by default: measured on 2 GPUs it shortens the reader chain by 5 % and the
  // transform by under 1 % (11.11 -> 11.01 ms, profiles/r02_exchange_ab.md) - the phase is paced by the writer.
  E.reader_depth = 0;
  static const int env_rd = getenv("OFFTB_READER_DEPTH") ? atoi(getenv("OFFTB_READER_DEPTH")) : 0;
  if (env_rd > 0 && r.depth > 0 && r.depth < 3) {
    const size_t reserve = 1024 + 256;
    const long long nw = (gw + w.sm_count - 1) / w.sm_count, nr = (gr + w.sm_count - 1) / w.sm_count;
    const size_t rslot = r.smem / (size_t)r.depth;
    for (int d = env_rd > 0 ? env_rd : 3; d > r.depth; --d) {
      if (rslot * d + reserve > (size_t)smem_sm) continue;
      if ((size_t)nw * (w.smem + reserve) + (size_t)nr * (rslot * d + reserve) <= (size_t)smem_sm) { E.reader_depth = d; break; }
    }
  }
  return true;
}

void *landing_slot(Engine &E, void *ring_base, int phase, int slot, bool inverse) {
  Ring &R = E.ring[phase - 1];
  const long long off = inverse ? R.send_off[slot] : R.recv_off[slot];
  return at(ring_base, off, E.esz);
}

// slot of a tile: by its running number over all executes of the plan, so that the tenant before it in the slot
// is always the tile `depth` numbers earlier - the one the flags make the writer wait for - also when the tiles
// of one execute are not a multiple of the depth
int slot_of(const Ring &R, int tile) { return (int)((R.tiles_done + (unsigned long long)tile) % (unsigned long long)R.depth); }

int fuse_writer(std::vector<Engine *> &engs, Engine &E, Launch &L, int phase, int tile, long long myT, bool inverse) {
  const Dims d = dims_of(E.po);
  Ring &R = E.ring[phase - 1];
  const int slot = slot_of(R, tile);
  const long long blk = block_elems(d, phase, myT);
  std::vector<int> members;
  int me;
  group_of(E, phase, members, me);
  if ((int)members.size() > OFFTB_MAX_GROUP) { set_error("fused exchange: groups of more than %d ranks are not supported", OFFTB_MAX_GROUP); return -1; }
  World &w = world();
  L.split = true;
  for (size_t j = 0; j < members.size(); ++j) {
    void *base = nullptr;
    if (w.local) {
      for (Engine *Q : engs) if (Q->po->rank == members[j]) base = Q->d_ring;
    } else {
      base = E.peer_ring[members[j]];
    }
    if (!base) { set_error("fused exchange: ring of rank %d is not mapped", members[j]); return -1; }
    Engine *owner = &E;
    if (w.local) for (Engine *Q : engs) if (Q->po->rank == members[j]) owner = Q;
    L.tab[j] = at(landing_slot(*owner, base, phase, slot, inverse), (long long)me * blk, E.esz);
  }
  if (!w.local) {
    const unsigned seq = (unsigned)(R.tiles_done + (unsigned long long)tile + 1ULL);
    if (R.tiles_done + (unsigned long long)tile >= (unsigned long long)R.depth) {
      L.wait_flags = E.d_flags->released[phase - 1][slot];
      L.wait_count = (int)members.size();
      L.wait_value = seq - (unsigned)R.depth;
    }
    for (size_t j = 0; j < members.size(); ++j)
      L.signal_ptrs[j] = &((XFlags *)E.peer_flags[members[j]])->arrived[phase - 1][slot][me];
    L.signal_count = (int)members.size();
    L.signal_value = seq;
    L.done_counter = &E.d_flags->done_counter[phase - 1][0][seq % OFFTB_DONE_SLOTS];
    L.grid_cap = E.grid_cap[0];
  }
  return 0;
}

int fuse_reader(Engine &E, Launch &L, int phase, int tile) {
  World &w = world();
  if (w.local) return 0;
  Ring &R = E.ring[phase - 1];
  std::vector<int> members;
  int me;
  group_of(E, phase, members, me);
  const unsigned seq = (unsigned)(R.tiles_done + (unsigned long long)tile + 1ULL);
  const int slot = slot_of(R, tile);
  L.wait_flags = E.d_flags->arrived[phase - 1][slot];
  L.wait_count = (int)members.size();
  L.wait_value = seq;
  for (size_t j = 0; j < members.size(); ++j)
    L.signal_ptrs[j] = &((XFlags *)E.peer_flags[members[j]])->released[phase - 1][slot][me];
  L.signal_count = (int)members.size();
  L.signal_value = seq;
  L.done_counter = &E.d_flags->done_counter[phase - 1][1][seq % OFFTB_DONE_SLOTS];
  L.grid_cap = E.grid_cap[1];
  return 0;
}

// ---- phases ---------------------------------------------------------------------------------

struct Bufs { void *U; void *A; };   // caller's array and the array between the phases (U or scratch)

// `writer`: this launch is the first of its tile (the one that feeds the exchange)
// `visit` = position of the tile in time (slot and sequence number of the flag protocol), `tile` = which planes it holds
int produce(std::vector<Engine *> &engs, Engine &E, const Bufs &b, int phase, int visit, int tile, long long myT, bool inverse, cudaStream_t st) {
  const Dims d = dims_of(E.po);
  Ring &R = E.ring[phase - 1];
  const int slot = slot_of(R, visit);
  const bool fused = E.xmode == XCHG_FUSED;
  // fused: forward, the kernel scatters into the peers' receive slots; backward, it reads this rank's send slot
  void *buf = R.send[slot];
  Launch L = phase == 1 ? L_k1(d, b.U, buf, (long long)tile * d.T1, myT) : L_k3(E, d, b.A, buf, (long long)tile * d.T2, myT);
  if (phase == 2 && E.sched == SCHED_PENCIL) { L.ry_level = 2; L.ry_x0 = 0; L.ry_lo = d.Ry; L.ry_hi = 10; }   // :1708, 1988
  if (fused && (inverse ? fuse_reader(E, L, phase, visit) : fuse_writer(engs, E, L, phase, visit, myT, inverse))) return -1;
  L.pdl = E.pdl_next;
  L.depth = E.depth_next;
  return run_launch(E, st, phase == 1 ? ST_K1 : ST_K3, L, inverse);
}

int consume(std::vector<Engine *> &engs, Engine &E, const Bufs &b, int phase, int visit, int tile, long long myT, bool inverse, cudaStream_t st) {
  const Dims d = dims_of(E.po);
  Ring &R = E.ring[phase - 1];
  const int slot = slot_of(R, visit);
  const bool fused = E.xmode == XCHG_FUSED;
  void *buf = R.recv[slot];
  Launch L = phase == 2 ? L_k4(E, d, buf, b.U, (long long)tile * d.T2, myT) : L_k2(d, buf, b.A, (long long)tile * d.T1, myT);
  if (phase == 1 && E.sched == SCHED_PENCIL) { L.ry_level = 1; L.ry_x0 = tile * d.T1; L.ry_lo = 0; L.ry_hi = d.Ry; }   // :1484
  if (fused && (inverse ? fuse_writer(engs, E, L, phase, visit, myT, inverse) : fuse_reader(E, L, phase, visit))) return -1;
  L.pdl = E.pdl_next;
  L.depth = E.depth_next;
  return run_launch(E, st, phase == 2 ? ST_K4 : ST_K2, L, inverse);
}

// Which of a phase's nb tiles is visited i-th (-1: none).  Forward: ascending.  Backward in phase 1: descending - when
// the array between the phases IS the caller's array (_S_ = 1) the reader of a tile rewrites x planes in the input layout,
// whose plane stride (istride[0]) is at least that of the planes the writers still have to read (M3*M4*p1); written regions
// must therefore trail the read ones, which they do going up in the forward direction and going down in the backward one.
// (With equal strides - even divisions - the two coincide plane by plane and the order does not matter; found as a wrong
// backward transform of 27x20x45 on 8 real ranks, where istride[0] = 24*M3 against 20*M3.  tests/test_tile_order.py checks
// the rule against an interval model of what each launch reads and writes.)
int tile_visited(int nb, int visit, int phase, bool inverse) {
  if (visit < 0 || visit >= nb) return -1;
  return inverse && phase == 1 ? nb - 1 - visit : visit;
}

// The tile pipeline of offt_3d_execute_phase1/2 (offt-compute.c:3501-3862):
//   produce(i); [wait(i-W)]; exchange(i); consume(i-W); ... drain the last W tiles.
// Forward: produce = K1/K3, consume = K2/K4.  Backward: the same loop with the roles and
// the ring directions swapped (consume's inverse gathers, produce's inverse scatters).
int run_phase(std::vector<Engine *> &engs, std::vector<Bufs> &bufs, int phase, bool inverse) {
  Engine &E0 = *engs[0];
  const Dims d0 = dims_of(E0.po);
  const int tiling = phase == 1 ? d0.T1 : d0.T2;
  const int W = phase == 1 ? d0.W1 : d0.W2;
  // planes of the tiled axis: the same for all members of an exchange group, but with an uneven division not for all
  // groups (offt-compute.c:3529, 3710 use the rank's own m1 / m3), so emulated-rank worlds count per plan
  std::vector<long long> planes(engs.size());
  int blocks = 0;
  for (size_t k = 0; k < engs.size(); ++k) {
    const Dims dk = dims_of(engs[k]->po);
    planes[k] = phase == 1 ? dk.m1 : dk.m3;
    blocks = std::max(blocks, (int)((planes[k] + tiling - 1) / tiling));
  }
  Ring &R0 = E0.ring[phase - 1];
  cudaStream_t sc = E0.s_user ? E0.s_user : E0.s_comp, sx = E0.s_comm;
  const bool fused = E0.xmode == XCHG_FUSED;
  auto tile_of = [&](size_t k, int visit) { return tile_visited((int)((planes[k] + tiling - 1) / tiling), visit, phase, inverse); };
  auto tile_Tk = [&](size_t k, int visit) {
    const int t = tile_of(k, visit);
    return t < 0 ? 0LL : std::max<long long>(0, std::min<long long>(tiling, planes[k] - (long long)t * tiling));
  };
  auto tile_T = [&](int visit) { return tile_Tk(0, visit); };
  // fused exchange between processes: readers on the second stream, ordered against the writers by flags alone
  bool two = false;
  E0.narrow_now = false;
  if (fused && !world().local && blocks > 1) {
    // shapes of tile 0's two launches, without launching
    auto try_overlap = [&]() -> int {
      FftShape shw, shr;
      Engine &E = E0;
      E.grid_cap[0] = E.grid_cap[1] = 0;
      E.dry_shape = &shw;
      int rc = inverse ? consume(engs, E, bufs[0], phase, 0, tile_of(0, 0), tile_T(0), true, sc) : produce(engs, E, bufs[0], phase, 0, tile_of(0, 0), tile_T(0), false, sc);
      E.dry_shape = &shr;
      if (!rc) rc = inverse ? produce(engs, E, bufs[0], phase, 0, tile_of(0, 0), tile_T(0), true, sc) : consume(engs, E, bufs[0], phase, 0, tile_of(0, 0), tile_T(0), false, sc);
      E.dry_shape = nullptr;
      if (rc) return -1;
      return plan_overlap(E, shw, shr) ? 1 : 0;
    };
    int ov = try_overlap();
    if (ov < 0) return -1;
    // Long strided transforms (2048 points: 131 KB and the whole register file per CTA at 64 bytes per index) leave no
    // room for a second kernel on the SM, and the exchange-bound writer would then run before the HBM-bound reader
    // instead of beside it.  Half-width tiles (32 bytes per index, half the threads, registers and shared memory) fit
    // two per SM: each kernel alone is slower, the phase as a whole is faster because the two overlap.
    // OFFTB_NARROW=0 keeps the full-width tiles on one stream.  The decision is a function of the plan and the device
    // only, so every rank takes the same one (the slot layout - the z chunk - depends on it).
    static const bool narrow_env = !(getenv("OFFTB_NARROW") && atoi(getenv("OFFTB_NARROW")) == 0);
    if (ov == 0 && narrow_env && overlap_wanted()) {
      E0.narrow_now = true;
      ov = try_overlap();
      if (ov < 0) return -1;
      if (ov == 0) { E0.narrow_now = false; E0.grid_cap[0] = E0.grid_cap[1] = 0; }
    }
    two = ov > 0;
  } else {
    E0.grid_cap[0] = E0.grid_cap[1] = 0;
  }
  cudaStream_t s2 = two ? sx : sc;
  if (two) {
    OFFTB_CUDA(cudaEventRecord(R0.packed[0], sc));
    OFFTB_CUDA(cudaStreamWaitEvent(s2, R0.packed[0], 0));
  }
  // Dependent-launch chains: on each of the two streams the launches of consecutive tiles are ordered by flags
  // alone (tile i+1's writer consumes nothing tile i's writer produced), so from the second launch on they carry
  // the programmatic-serialization attribute and every launch releases its successor once its CTAs are past their
  // flag wait: the successor's CTAs fill the SMs as this launch's CTAs leave instead of waiting for the last one
  // (per-tile ramp and tail).  The first launch of a chain is an ordinary one - it does consume what ran before it
  // on the stream.  Blocked CTAs stay bounded: a launch is released only after its predecessor's CTAs have all
  // passed their wait, so at most one grid per stream can be spinning and the grid caps still leave the other
  // stream its share of the SMs.  OFFTB_PDL=0 turns the chains off.
  static const bool pdl_env = !(getenv("OFFTB_PDL") && atoi(getenv("OFFTB_PDL")) == 0);
  // chained launches may finish out of order; where the reader of one tile rewrites memory next to what the writer of a
  // neighbouring tile reads (phase 1 in place with unequal plane strides, see tile_visited) the writers must finish in order
  const bool in_place_skew = phase == 1 && bufs[0].A == bufs[0].U && d0.isx != d0.dX;
  const bool pdl = two && pdl_env && !in_place_skew;
  int n_first = 0, n_second = 0;
  cudaEvent_t ce[4] = {nullptr, nullptr, nullptr, nullptr};
  E0.chain_timing = pdl && E0.stage_timing && !E0.async;
  if (E0.chain_timing) for (cudaEvent_t &e : ce) e = pool_event(E0);
  auto first = [&](size_t k, int i) {
    Engine &E = *engs[k];
    if (tile_Tk(k, i) <= 0) return 0;
    E.pdl_next = pdl ? (1 | (n_first ? 2 : 0)) : 0;
    E.depth_next = 0;
    if (E.chain_timing && !n_first) cudaEventRecord(ce[0], sc);
    ++n_first;
    return inverse ? consume(engs, E, bufs[k], phase, i, tile_of(k, i), tile_Tk(k, i), true, sc) : produce(engs, E, bufs[k], phase, i, tile_of(k, i), tile_Tk(k, i), false, sc);
  };
  auto second = [&](size_t k, int i) {
    Engine &E = *engs[k];
    if (tile_Tk(k, i) <= 0) return 0;
    E.pdl_next = pdl ? (1 | (n_second ? 2 : 0)) : 0;
    E.depth_next = two ? E.reader_depth : 0;
    if (E.chain_timing && !n_second) cudaEventRecord(ce[2], s2);
    ++n_second;
    return inverse ? produce(engs, E, bufs[k], phase, i, tile_of(k, i), tile_Tk(k, i), true, s2) : consume(engs, E, bufs[k], phase, i, tile_of(k, i), tile_Tk(k, i), false, s2);
  };
  std::vector<long long> tts(engs.size());
  for (int i = 0; i < blocks; ++i) {
    const int slot = slot_of(R0, i);
    for (size_t k = 0; k < engs.size(); ++k)
      if (first(k, i)) return -1;
    if (!fused) {
      OFFTB_CUDA(cudaEventRecord(R0.packed[slot], sc));
      OFFTB_CUDA(cudaStreamWaitEvent(sx, R0.packed[slot], 0));
      for (size_t k = 0; k < engs.size(); ++k) tts[k] = tile_Tk(k, i);
      if (exchange(engs, phase, slot, tts, inverse, sx)) return -1;
      OFFTB_CUDA(cudaEventRecord(R0.recvd[slot], sx));
    }
    if (i >= W) {
      const int j = i - W;
      if (!fused) OFFTB_CUDA(cudaStreamWaitEvent(sc, R0.recvd[slot_of(R0, j)], 0));
      for (size_t k = 0; k < engs.size(); ++k)
        if (second(k, j)) return -1;
    }
  }
  for (int j = std::max(blocks - W, 0); j < blocks; ++j) {
    if (!fused) OFFTB_CUDA(cudaStreamWaitEvent(sc, R0.recvd[slot_of(R0, j)], 0));
    for (size_t k = 0; k < engs.size(); ++k)
      if (second(k, j)) return -1;
  }
  if (E0.chain_timing) {
    // one event pair per chain: the span from the first launch's start to the last launch's end, flag waits included
    cudaEventRecord(ce[1], sc);
    cudaEventRecord(ce[3], s2);
    const int st_first = inverse ? (phase == 1 ? ST_K2 : ST_K4) : (phase == 1 ? ST_K1 : ST_K3);
    const int st_second = inverse ? (phase == 1 ? ST_K1 : ST_K3) : (phase == 1 ? ST_K2 : ST_K4);
    E0.timed.push_back({st_first, {ce[0], ce[1]}});
    E0.timed.push_back({st_second, {ce[2], ce[3]}});
    E0.chain_timing = false;
  }
  for (Engine *Ep : engs) { Ep->pdl_next = 0; Ep->depth_next = 0; }
  E0.narrow_now = false;
  if (two) {
    OFFTB_CUDA(cudaEventRecord(R0.recvd[0], s2));
    OFFTB_CUDA(cudaStreamWaitEvent(sc, R0.recvd[0], 0));
  }
  for (Engine *Ep : engs) Ep->ring[phase - 1].tiles_done += (unsigned long long)blocks;
  return 0;
}

// Local z pass.  Complex plans: one launch.  Real-to-complex plans whose half length has a register-butterfly kernel:
// the rows are transformed as Nz/2 complex points in place and r2c_pass.cu turns that spectrum into the Nz/2+1 points of
// the real transform on the way to where the launch would have stored them (backward: the reverse).  Other real plans
// (odd Nz, half lengths without a fast kernel) and the fused z pass of phase 1 run on the any-length kernel.
int run_z_pass(Engine &E, cudaStream_t st, Launch L, bool inverse) {
  if (!(L.r2c && E.r2c_fast)) return run_launch(E, st, ST_K1, L, inverse);
  const int H = L.N / 2;
  Launch Lh = L;
  Lh.N = H; Lh.r2c = 0; Lh.tw = E.tw_half;
  Lh.out = const_cast<void *>(L.in); Lh.om = L.im;     // in place in the caller's array, where the real rows live
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  const bool timed = E.stage_timing && !E.async;
  const bool was = E.stage_timing;
  if (timed) { e0 = pool_event(E); e1 = pool_event(E); cudaEventRecord(e0, st); E.stage_timing = false; }   // one event pair for both kernels
  int rc = 0;
  cudaError_t ce = cudaSuccess;
  if (!inverse) {
    rc = run_launch(E, st, ST_K1, Lh, false);
    if (!rc && !E.dry_shape) ce = r2c_step_launch(E.prec, false, L.in, L.out, E.tw_r2c, L.im, L.om, H, L.nbatch, st);
  } else {
    if (!E.dry_shape) ce = r2c_step_launch(E.prec, true, L.out, const_cast<void *>(L.in), E.tw_r2c, L.om, L.im, H, L.nbatch, st);
    if (ce == cudaSuccess) rc = run_launch(E, st, ST_K1, Lh, true);
  }
  E.stage_timing = was;
  if (ce != cudaSuccess) { set_error("r2c step (N=%d, batch=%lld): %s", L.N, L.nbatch, cudaGetErrorString(ce)); return -1; }
  if (timed) { cudaEventRecord(e1, st); E.timed.push_back({ST_K1, {e0, e1}}); }
  if (!rc && !E.dry_shape) E.launches++;
  return rc;
}

int run_schedule(std::vector<Engine *> &engs, std::vector<Bufs> &bufs, bool inverse) {
  Engine &E0 = *engs[0];
  cudaStream_t sc = E0.s_user ? E0.s_user : E0.s_comp;
  // the forward schedule as a list of steps; the backward transform walks it in reverse
  enum Step { STEP_Z_LOCAL, STEP_Y_LOCAL, STEP_X_LOCAL, STEP_PHASE1, STEP_PHASE2, STEP_Z_SWAP, STEP_X_SWAP, STEP_Y_OUT };
  std::vector<Step> steps;
  const bool swapped = E0.sched == SCHED_SINGLE && E0.po->params->v[_S_] == 1 && E0.d_scratch != nullptr;
  switch (E0.sched) {
    case SCHED_SINGLE:
      if (swapped) steps = {STEP_Z_SWAP, STEP_X_SWAP, STEP_Y_OUT};
      else steps = {STEP_Z_LOCAL, STEP_Y_LOCAL, STEP_X_LOCAL};
      break;
    case SCHED_SLAB_1XP: steps = {STEP_PHASE1, STEP_X_LOCAL}; break;        // offt-compute.c:3896-3950
    case SCHED_SLAB_PX1: steps = {STEP_Z_LOCAL, STEP_PHASE2}; break;        // offt-compute.c:3951-3998
    case SCHED_PENCIL: steps = {STEP_PHASE1, STEP_PHASE2}; break;           // offt-compute.c:3999-4037
  }
  if (inverse) std::reverse(steps.begin(), steps.end());
  for (Step s : steps) {
    if (s == STEP_PHASE1 || s == STEP_PHASE2) {
      const auto h0 = std::chrono::steady_clock::now();
      if (run_phase(engs, bufs, s == STEP_PHASE1 ? 1 : 2, inverse)) return -1;
      const double hs = std::chrono::duration<double>(std::chrono::steady_clock::now() - h0).count();
      for (Engine *Ep : engs) Ep->post_s[s == STEP_PHASE1 ? 0 : 1] = hs;
      continue;
    }
    for (size_t k = 0; k < engs.size(); ++k) {
      Engine &E = *engs[k];
      const Dims d = dims_of(E.po);
      int rc = 0;
      if (s == STEP_Z_SWAP) rc = run_z_pass(E, sc, L_fftz_swap(d, bufs[k].U, bufs[k].A), inverse);
      else if (s == STEP_X_SWAP) rc = run_launch(E, sc, ST_K4, L_fftx_swap(d, bufs[k].A, bufs[k].U), inverse);
      else if (s == STEP_Y_OUT) rc = run_launch(E, sc, ST_K2, L_ffty_out(d, bufs[k].U), inverse);
      else if (s == STEP_Z_LOCAL) rc = run_z_pass(E, sc, L_fftz_local(d, bufs[k].U, bufs[k].A, 0, d.m1), inverse);
      else if (s == STEP_Y_LOCAL) rc = run_launch(E, sc, ST_K2, L_ffty_local(d, bufs[k].A), inverse);
      else rc = run_launch(E, sc, ST_K4, L_fftx_local(d, bufs[k].A, bufs[k].U), inverse);
      if (rc) return -1;
    }
  }
  return 0;
}

void free_ring(Ring &R) {
  for (cudaEvent_t e : R.packed) cudaEventDestroy(e);
  for (cudaEvent_t e : R.recvd) cudaEventDestroy(e);
  R = Ring();
}

}  // namespace

// ---------------------------------------------------------------------------------- create / destroy

int engine_create(struct _offt_plan *po) {
  World &w = world();
  if (!w.up) { set_error("no world: call offtb_world_init / offtb_world_init_local (or the compat MPI_Init) first"); return -1; }
  const int *v = po->params->v;
  if (check_supported(po->Nx, po->Ny, po->Nz, po->p, v[_P1_], po->is_r2c)) return -1;
  if (v[_T1_] < 1 || v[_T2_] < 1 || v[_W1_] < 0 || v[_W2_] < 0 || v[_W1_] > 64 || v[_W2_] > 64) {
    set_error("tile sizes must be >= 1 and windows in 0..64 (T1 %d W1 %d T2 %d W2 %d)", v[_T1_], v[_W1_], v[_T2_], v[_W2_]);
    return -1;
  }
  Engine *E = new Engine();
  po->b200 = E;
  E->po = po;
  extern int g_default_precision;
  E->prec = g_default_precision;
  E->esz = E->prec == PREC_F64 ? 16 : 8;
  const struct _offt_comm *c = po->comm;
  E->rank_x = po->rank / c->p2; E->rank_y = po->rank % c->p2;
  if (po->p == 1) E->sched = SCHED_SINGLE;
  else if (po->is_oned && c->p1 == 1) E->sched = SCHED_SLAB_1XP;
  else if (po->is_oned && c->p1 == po->p) E->sched = SCHED_SLAB_PX1;
  else E->sched = SCHED_PENCIL;
  E->alloc = alloc_elems(po->Nx, po->Ny, po->Nz, po->p, c->p1, po->is_r2c);
  const int Ns[3] = {po->Nx, po->Ny, po->Nz};
  for (int a = 0; a < 3; ++a)
    if (make_twiddles(Ns[a], E->prec, &E->tw[a]) || make_twiddles_full(Ns[a], E->prec, &E->tw_full[a])) return -1;
  // real-to-complex plans: the local z pass runs as a half-length complex transform plus the O(N) step of r2c_pass.cu
  // when that half length has a register-butterfly kernel (OFFTB_R2C_FAST=0: always the any-length kernel)
  {
    FftKernelInfo hinfo;
    const bool env_on = !(getenv("OFFTB_R2C_FAST") && atoi(getenv("OFFTB_R2C_FAST")) == 0);
    if (po->is_r2c && env_on && po->Nz % 2 == 0 && po->Nz >= 4 && fft_kernel_info(po->Nz / 2, E->prec, &hinfo)) {
      const int H = po->Nz / 2;
      if (make_twiddles(H, E->prec, &E->tw_half)) return -1;
      const long double two_pi = 6.283185307179586476925286766559005768L;
      const int cnt = H / 2 + 1;
      std::vector<double> hd(2 * (size_t)cnt);
      std::vector<float> hf(2 * (size_t)cnt);
      for (int k = 0; k < cnt; ++k) {
        const long double ang = two_pi * (long double)k / (long double)po->Nz;
        hd[2 * k] = (double)cosl(ang); hd[2 * k + 1] = (double)-sinl(ang);
        hf[2 * k] = (float)cosl(ang); hf[2 * k + 1] = (float)-sinl(ang);
      }
      const size_t bytes = (size_t)cnt * E->esz;
      OFFTB_CUDA(cudaMalloc(&E->tw_r2c, bytes));
      OFFTB_CUDA(cudaMemcpy(E->tw_r2c, E->prec == PREC_F64 ? (void *)hd.data() : (void *)hf.data(), bytes, cudaMemcpyHostToDevice));
      E->r2c_fast = true;
    }
  }
  OFFTB_CUDA(cudaStreamCreateWithFlags(&E->s_comp, cudaStreamNonBlocking));
  OFFTB_CUDA(cudaStreamCreateWithFlags(&E->s_comm, cudaStreamNonBlocking));
  OFFTB_CUDA(cudaEventCreate(&E->ev_begin));
  OFFTB_CUDA(cudaEventCreate(&E->ev_end));
  // a second array: for the transposed output layouts, and for the single-rank x-y-z schedule (see L_fftz_swap);
  // OFFTB_SINGLE_INPLACE=1 keeps the latter in place (three passes in the caller's array, no scratch)
  const bool single_swap = E->sched == SCHED_SINGLE && v[_S_] == 1 && !(getenv("OFFTB_SINGLE_INPLACE") && atoi(getenv("OFFTB_SINGLE_INPLACE")));
  if (!v[_S_] || single_swap) OFFTB_CUDA(cudaMalloc(&E->d_scratch, (size_t)E->alloc * E->esz));
  // rings: (W+1) {send, recv} pairs per phase carved from one chunk (set_buffer_chunk / set_buffer,
  // offt-compute.c:684-746: slot sizes T1*M2*M3*p2 and M1*M4*p1*T2)
  const Dims d = dims_of(po);
  const bool use1 = E->sched == SCHED_PENCIL || E->sched == SCHED_SLAB_1XP;
  const bool use2 = E->sched == SCHED_PENCIL || E->sched == SCHED_SLAB_PX1;
  long long slot[2] = {use1 ? (long long)d.T1 * d.M2 * d.M3 * d.p2 : 0, use2 ? d.M1 * d.M4 * d.p1 * (long long)d.T2 : 0};
  int depth[2] = {d.W1 + 1, d.W2 + 1};
  // The reference lets the two phases share one chunk (set_buffer_chunk, offt-compute.c:684-710).  Here the phases
  // get disjoint parts: in the fused exchange a rank that has entered phase 2 stores into its peers' phase-2 slots
  // while a slower peer may still be reading its phase-1 slots, and only same-phase slots are guarded by flags.
  const long long part[2] = {2 * slot[0] * depth[0], 2 * slot[1] * depth[1]};
  const long long base[2] = {0, part[0]};
  const long long need = part[0] + part[1];
  bool ring_ok = true, ring_mine = true;
  if (need > 0 && cudaMalloc(&E->d_ring, (size_t)need * E->esz) != cudaSuccess) {
    cudaGetLastError();
    set_error("ring of %.1f MiB does not fit in device memory (T1 %d W1 %d T2 %d W2 %d)", (double)need * E->esz / 1048576.0, d.T1, d.W1, d.T2, d.W2);
    E->d_ring = nullptr;
    ring_ok = ring_mine = false;
  }
  // every rank must take the same path from here on: agree whether all rings exist
  if (w.nccl && po->p > 1 && world_agree_ok(ring_ok) != 0) ring_ok = false;
  if (!ring_ok) { if (ring_mine) set_error("ring allocation failed on another rank"); return -1; }
  for (int ph = 0; ph < 2; ++ph) {
    Ring &R = E->ring[ph];
    R.depth = depth[ph]; R.slot_elems = slot[ph];
    for (int s = 0; s < R.depth; ++s) {
      R.send.push_back(slot[ph] ? at(E->d_ring, base[ph] + (2LL * s) * slot[ph], E->esz) : nullptr);
      R.recv.push_back(slot[ph] ? at(E->d_ring, base[ph] + (2LL * s + 1) * slot[ph], E->esz) : nullptr);
      R.send_off.push_back(base[ph] + (2LL * s) * slot[ph]);
      R.recv_off.push_back(base[ph] + (2LL * s + 1) * slot[ph]);
      cudaEvent_t a, b;
      OFFTB_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
      OFFTB_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
      R.packed.push_back(a); R.recvd.push_back(b);
    }
  }
  // exchange mode: kernels that store into the peers' slots (default), or grouped ncclSend/ncclRecv between
  // send and receive slots (OFFTB_EXCHANGE=nccl, and whenever peer mapping is not possible)
  const char *xe = getenv("OFFTB_EXCHANGE");
  const bool want_fused = !(xe && strcmp(xe, "nccl") == 0) && std::max(c->p1, c->p2) <= OFFTB_MAX_GROUP;
  if (need > 0 && po->p > 1 && want_fused) {
    if (w.local) {
      E->xmode = XCHG_FUSED;
    } else {
      // a rank whose allocations failed still takes part in the collective share (with nothing to offer), so that
      // all ranks leave it together and agree on the outcome
      if (cudaMalloc((void **)&E->d_flags, sizeof(XFlags)) != cudaSuccess) { cudaGetLastError(); E->d_flags = nullptr; }
      if (E->d_flags) cudaMemset(E->d_flags, 0, sizeof(XFlags));
      if (cudaHostAlloc((void **)&E->h_error, sizeof(unsigned), cudaHostAllocMapped) == cudaSuccess) {
        *E->h_error = 0;
        if (cudaHostGetDevicePointer((void **)&E->d_error, E->h_error, 0) != cudaSuccess) { cudaGetLastError(); E->d_error = nullptr; }
      } else { cudaGetLastError(); E->h_error = nullptr; }
      const char *ts = getenv("OFFTB_FLAG_TIMEOUT_S");
      const double tsec = ts ? atof(ts) : 300.0;
      E->wait_timeout_ns = tsec > 0 ? (unsigned long long)(tsec * 1e9) : 0ULL;
      const int r1 = world_ipc_share(ring_ok ? E->d_ring : nullptr, E->peer_ring);
      const int r2 = r1 ? -1 : world_ipc_share(E->d_flags, E->peer_flags);
      if (r1 == 0 && r2 == 0) {
        E->xmode = XCHG_FUSED;
      } else {
        if (!po->rank) fprintf(stderr, "offt_b200: peer mapping unavailable (%s); exchanging with NCCL send/recv\n", last_error());
        world_ipc_release(E->peer_ring);
        world_ipc_release(E->peer_flags);
      }
    }
  }
  po->buffer_chunk = E->d_ring;
  po->buffers1 = &E->ring[0];
  po->buffers2 = &E->ring[1];
  return 0;
}

void engine_destroy(struct _offt_plan *po) {
  Engine *E = (Engine *)po->b200;
  if (!E) return;
  cudaDeviceSynchronize();
  if (!E->peer_ring.empty() || !E->peer_flags.empty()) {
    // peers may still be storing flags into this rank's memory: leave together (offt_3d_fin is collective) ...
    offtb_world_barrier();
    world_ipc_release(E->peer_ring);
    world_ipc_release(E->peer_flags);
    // ... and free the exported allocations only after every importer has closed its mapping of them
    offtb_world_barrier();
  }
  cudaFree(E->d_flags);
  if (E->h_error) cudaFreeHost(E->h_error);
  if (E->registered_host) cudaHostUnregister(E->registered_host);
  for (int a = 0; a < 3; ++a) { cudaFree(E->tw[a]); cudaFree(E->tw_full[a]); }
  cudaFree(E->tw_half); cudaFree(E->tw_r2c);
  cudaFree(E->d_user); cudaFree(E->d_scratch); cudaFree(E->d_ring);
  free_ring(E->ring[0]); free_ring(E->ring[1]);
  for (cudaEvent_t e : E->event_pool) cudaEventDestroy(e);
  if (E->ev_begin) cudaEventDestroy(E->ev_begin);
  if (E->ev_end) cudaEventDestroy(E->ev_end);
  if (E->s_comp) cudaStreamDestroy(E->s_comp);
  if (E->s_comm) cudaStreamDestroy(E->s_comm);
  delete E;
  po->b200 = nullptr;
  po->buffer_chunk = po->buffers1 = po->buffers2 = nullptr;
}

// ---------------------------------------------------------------------------------- execute

int engine_execute(std::vector<struct _offt_plan *> &group, std::vector<double *> &arrays, bool inverse) {
  std::vector<Engine *> engs;
  for (auto *po : group) {
    if (!po || !po->b200) { set_error("plan has no engine"); return -1; }
    engs.push_back((Engine *)po->b200);
  }
  Engine &E0 = *engs[0];
  if (E0.failed) { set_error("this plan's exchange timed out earlier; its flags are no longer consistent - destroy it"); return -1; }
  cudaStream_t sc = E0.s_user ? E0.s_user : E0.s_comp;
  std::vector<Bufs> bufs(engs.size());
  std::vector<bool> on_host(engs.size(), false);
  for (size_t k = 0; k < engs.size(); ++k) {
    Engine &E = *engs[k];
    E.launches = 0; E.timed.clear(); E.event_next = 0; E.post_s[0] = E.post_s[1] = 0.0;
    for (double &m : E.stage_ms) m = 0.0;
    cudaPointerAttributes attr;
    cudaError_t pe = cudaPointerGetAttributes(&attr, arrays[k]);
    if (pe != cudaSuccess) { cudaGetLastError(); attr.type = cudaMemoryTypeUnregistered; }
    on_host[k] = !(attr.type == cudaMemoryTypeDevice || attr.type == cudaMemoryTypeManaged);
    if (on_host[k] && E.async) { set_error("asynchronous execution needs device arrays"); return -1; }
  }
  OFFTB_CUDA(cudaEventRecord(E0.ev_begin, sc));
  for (size_t k = 0; k < engs.size(); ++k) {
    Engine &E = *engs[k];
    const size_t bytes = (size_t)E.alloc * E.esz;
    if (on_host[k]) {
      if (!E.d_user) OFFTB_CUDA(cudaMalloc(&E.d_user, bytes));
      if (E.registered_host != arrays[k]) {
        // pin the caller's array in place once: the reference driver reuses it every repetition
        if (E.registered_host) { cudaHostUnregister(E.registered_host); E.registered_host = nullptr; }
        cudaError_t re = cudaHostRegister(arrays[k], bytes, cudaHostRegisterDefault);
        if (re == cudaSuccess) { E.registered_host = arrays[k]; E.registered_bytes = bytes; }
        else cudaGetLastError();   // already pinned by the caller, or not pinnable: plain copies still work
      }
      cudaEvent_t e0 = nullptr, e1 = nullptr;
      if (E.stage_timing && !E.async) { e0 = pool_event(E); e1 = pool_event(E); cudaEventRecord(e0, sc); }
      OFFTB_CUDA(cudaMemcpyAsync(E.d_user, arrays[k], bytes, cudaMemcpyHostToDevice, sc));
      if (e1) { cudaEventRecord(e1, sc); E.timed.push_back({ST_H2D, {e0, e1}}); }
      bufs[k].U = E.d_user;
    } else {
      bufs[k].U = arrays[k];
    }
    bufs[k].A = E.d_scratch ? E.d_scratch : bufs[k].U;
  }
  if (run_schedule(engs, bufs, inverse)) return -1;
  for (size_t k = 0; k < engs.size(); ++k) {
    Engine &E = *engs[k];
    if (!on_host[k]) continue;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (E.stage_timing && !E.async) { e0 = pool_event(E); e1 = pool_event(E); cudaEventRecord(e0, sc); }
    OFFTB_CUDA(cudaMemcpyAsync(arrays[k], E.d_user, (size_t)E.alloc * E.esz, cudaMemcpyDeviceToHost, sc));
    if (e1) { cudaEventRecord(e1, sc); E.timed.push_back({ST_D2H, {e0, e1}}); }
  }
  OFFTB_CUDA(cudaEventRecord(E0.ev_end, sc));
  if (E0.async) return 0;
  OFFTB_CUDA(cudaStreamSynchronize(sc));
  OFFTB_CUDA(cudaStreamSynchronize(E0.s_comm));
  if (E0.h_error && *(volatile unsigned *)E0.h_error) {
    const unsigned code = *(volatile unsigned *)E0.h_error;
    E0.failed = true;
    set_error("exchange timed out after %.0f s waiting for group member %u's flag (a peer rank is missing, far behind, or runs a "
              "different plan; OFFTB_FLAG_TIMEOUT_S sets the limit, 0 waits for ever)", (double)E0.wait_timeout_ns * 1e-9, code - 1);
    return -1;
  }
  float ms = 0.f;
  OFFTB_CUDA(cudaEventElapsedTime(&ms, E0.ev_begin, E0.ev_end));
  for (Engine *Ep : engs) {
    Ep->last_ms = ms;
    for (auto &t : Ep->timed) {
      float m = 0.f;
      if (cudaEventElapsedTime(&m, t.second.first, t.second.second) == cudaSuccess) Ep->stage_ms[t.first] += m;
    }
  }
  return 0;
}

}  // namespace offtb

// the visit order of run_phase for the host-side model test (tests/test_tile_order.py)
extern "C" int offtb_tile_visited(int nb, int visit, int phase, int inverse) { return offtb::tile_visited(nb, visit, phase, inverse != 0); }
