// extern "C" entry points: the reference's plan / execute / destroy API (include/offt.h) and
// the additions of include/offt_b200.h.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "engine.h"

namespace offtb {
int g_default_precision = PREC_F64;
int g_force_generic = getenv("OFFTB_GENERIC") ? atoi(getenv("OFFTB_GENERIC")) : 0;

static double now_s() {
  using namespace std::chrono;
  return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}

static Engine *eng(const struct _offt_plan *po) { return po ? (Engine *)po->b200 : nullptr; }

// po->t[]: the reference's 16 wall-clock buckets (offt.h:171-188), filled on every execute from CUDA events.
//   ALL                      first kernel -> last kernel (host <-> device copies included when the caller passed host memory)
//   FFTz, FFTy1, FFTy2, FFTx the four fused launches K1..K4 (a dependent-launch chain is booked as its span)
//   WAIT1, WAIT2             device time of the exchanges where they are separate operations (NCCL send/recv mode);
//                            in the fused exchange the NVLink stores are part of FFTz / FFTy2
//   INIT1, INIT2             host seconds spent enqueueing the phase - what posting the collectives costs the caller
//   PACK*, UNPACK*, TRANSPOSE always 0: those steps are the address maps of the FFT launches, they have no time of their own
//   TEST1, TEST2             always 0: there is no MPI_Test to poll
static void fill_timers(struct _offt_plan *po) {
  Engine *E = eng(po);
  double *t = po->t;
  memset(t, 0, sizeof(double) * GES);
  t[ALL] = E->last_ms * 1e-3;
  if (!E->stage_timing) return;
  const double *s = E->stage_ms;
  t[FFTz] = s[ST_K1] * 1e-3;
  t[FFTy1] = s[ST_K2] * 1e-3;
  t[FFTy2] = s[ST_K3] * 1e-3;
  t[FFTx] = s[ST_K4] * 1e-3;
  t[WAIT1] = s[ST_X1] * 1e-3;
  t[WAIT2] = s[ST_X2] * 1e-3;
  t[INIT1] = E->post_s[0];
  t[INIT2] = E->post_s[1];
}
}  // namespace offtb

using namespace offtb;

extern "C" {

struct _offt_comm *offt_comm_malloc(struct _offt_plan *po) {
  struct _offt_comm *c = (struct _offt_comm *)calloc(1, sizeof(struct _offt_comm));
  comm_fill(c, po->Nx, po->Ny, po->Nz, po->p, po->params->v[_P1_], po->rank, po->params->v[_S_], po->is_equalxy, po->is_r2c);
  return c;
}

void offt_comm_free(struct _offt_comm *comm) { free(comm); }

int offtb_comm_fill(struct _offt_comm *c, int Nx, int Ny, int Nz, int p, int p1, int rank, int S, int is_equalxy) {
  if (p < 1 || p1 < 1 || p % p1 || rank < 0 || rank >= p) { set_error("bad process grid"); return -1; }
  comm_fill(c, Nx, Ny, Nz, p, p1, rank, S, is_equalxy);
  return 0;
}

int offtb_comm_fill_r2c(struct _offt_comm *c, int Nx, int Ny, int Nz, int p, int p1, int rank, int S, int is_equalxy, int is_r2c) {
  if (p < 1 || p1 < 1 || p % p1 || rank < 0 || rank >= p) { set_error("bad process grid"); return -1; }
  comm_fill(c, Nx, Ny, Nz, p, p1, rank, S, is_equalxy, is_r2c);
  return 0;
}

long long offtb_alloc_elems(int Nx, int Ny, int Nz, int p, int p1) { return alloc_elems(Nx, Ny, Nz, p, p1); }
long long offtb_alloc_elems_r2c(int Nx, int Ny, int Nz, int p, int p1, int is_r2c) { return alloc_elems(Nx, Ny, Nz, p, p1, is_r2c); }
int offtb_check_supported(int Nx, int Ny, int Nz, int p, int p1) { return check_supported(Nx, Ny, Nz, p, p1); }

struct _offt_plan *offt_3d_init(int Nx, int Ny, int Nz, double *in, double *out, int is_r2c, int fftw_flag,
                                int is_oned, int is_a2a, int is_equalxy, int is_notest, int ah_strategy,
                                int max_loop, int tuning_mode, int is_W0, int extrapolation_window,
                                struct _offt_params *custom_params) {
  const double t0 = now_s();
  World &w = world();
  if (!w.up) {
    set_error("no world: call offtb_world_init / offtb_world_init_local (or the compat MPI_Init) first");
    fatal_or_return("offt_3d_init");
    return nullptr;
  }
  struct _offt_plan *po = (struct _offt_plan *)calloc(1, sizeof(struct _offt_plan));
  po->Nx = Nx; po->Ny = Ny; po->Nz = Nz;
  po->p = w.size; po->rank = w.rank;
  po->is_r2c = is_r2c; po->fftw_flag = fftw_flag; po->is_oned = is_oned; po->is_a2a = is_a2a;
  po->is_equalxy = is_equalxy; po->is_notest = is_notest; po->ah_strategy = ah_strategy;
  po->max_loop = max_loop; po->tuning_mode = tuning_mode; po->is_W0 = is_W0;
  po->extrapolation_window = extrapolation_window;
  po->params = (struct _offt_params *)calloc(1, sizeof(struct _offt_params));
  // defaults on the value grid, then the caller's non-negative overrides (which need not be on it)
  // real-to-complex plans size every z-related tunable by the Nz/2+1 complex points that remain (offt-compute.c:3008, 3045, 3141)
  params_default(Nx, Ny, is_r2c ? Nz / 2 + 1 : Nz, po->p, is_W0, is_notest, po->params->v);
  po->params->is_converged = 1;
  if (!po->rank) print_params(po->params->v);
  // the caller's overrides count only without tuning: the reference's tuner starts from the defaults and searches all
  // 24 tunables itself ("custom values must be adjusted for tuning", offt-compute.c:3417-3423)
  if (custom_params && max_loop == 0)
    for (int i = 0; i < PARAM_COUNT; ++i)
      if (custom_params->v[i] >= 0) po->params->v[i] = custom_params->v[i];
  int rc = 0;
  if (po->params->v[_P1_] < 1 || po->p % po->params->v[_P1_]) {
    set_error("P1 = %d does not divide p = %d", po->params->v[_P1_], po->p);
    rc = -1;
  }
  if (!rc) {
    po->comm = offt_comm_malloc(po);
    const double tb = now_s();
    rc = engine_create(po);
    po->t_init[INIT_BUFFER] = now_s() - tb;
  }
  if (!rc && max_loop > 0) {
    const double ta = now_s();
    rc = ah_tuning(po, in, out) < 0 ? -1 : 0;
    po->t_init[INIT_AH] = now_s() - ta;
  }
  if (rc) {
    fatal_or_return("offt_3d_init");
    if (po->b200) engine_destroy(po);
    if (po->comm) offt_comm_free(po->comm);
    free(po->params);
    free(po);
    return nullptr;
  }
  po->t_init[INIT_ALL] = now_s() - t0;
  if (!po->rank) {
    const struct _offt_comm *c = po->comm;
    printf("M1 %d M2 %d M3 %d M4 %d m1 %d m2 %d m3 %d m4 %d\n", c->M1, c->M2, c->M3, c->M4, c->m1, c->m2, c->m3, c->m4);
  }
  return po;
}

// The reference driver reads po->comm->ostride for its -v print AFTER offt_3d_fin (run-fft.c:421 vs :477-478),
// which only works there because freed heap blocks keep their tail.  The plan, its parameters and its layout
// descriptor (a few hundred bytes) therefore stay readable until the world is torn down (MPI_Finalize).
static std::vector<struct _offt_plan *> &graveyard() {
  static std::vector<struct _offt_plan *> g;
  return g;
}

void offtb_release_finished_plans(void) {
  for (struct _offt_plan *po : graveyard()) {
    offt_comm_free(po->comm);
    free(po->params);
    free(po);
  }
  graveyard().clear();
}

void offt_3d_fin(struct _offt_plan *po) {
  if (!po) return;
  engine_destroy(po);
  graveyard().push_back(po);
  if (graveyard().size() > 64) {   // long-lived processes that create many plans: keep only the recent ones
    struct _offt_plan *old = graveyard().front();
    graveyard().erase(graveyard().begin());
    offt_comm_free(old->comm);
    free(old->params);
    free(old);
  }
}

static int execute_one(struct _offt_plan *po, double *in, double *out, bool inverse) {
  if (!po || !po->b200) { set_error("null plan"); return -1; }
  if (in != out) { set_error("in-place only: in must equal out (as in the reference, offt-compute.c:3866)"); return -1; }
  if (world().local && world().size > 1) {
    set_error("local worlds run all ranks together: use offtb_execute_group");
    return -1;
  }
  std::vector<struct _offt_plan *> g{po};
  std::vector<double *> a{out};
  int rc = engine_execute(g, a, inverse);
  if (!rc && !eng(po)->async) fill_timers(po);
  return rc;
}

void offt_3d_execute(struct _offt_plan *po, double *in, double *out, int is_tuning) {
  (void)is_tuning;
  static const bool dbg = getenv("OFFTB_DEBUG") != nullptr;
  if (dbg) fprintf(stderr, "offt_3d_execute(%p, %p, %p): before\n", (void *)po, (void *)in, (void *)out);
  if (execute_one(po, in, out, false)) fatal_or_return("offt_3d_execute");
  if (dbg) {
    cudaPointerAttributes attr;
    const bool host = cudaPointerGetAttributes(&attr, out) != cudaSuccess || attr.type == cudaMemoryTypeUnregistered || attr.type == cudaMemoryTypeHost;
    cudaGetLastError();
    if (host) fprintf(stderr, "offt_3d_execute: after, out[0..3] = %g %g %g %g, %d launches, %.3f ms\n", out[0], out[1], out[2], out[3],
                      eng(po)->launches, eng(po)->last_ms);
  }
}

int offt_3d_execute_inverse(struct _offt_plan *po, double *in, double *out) {
  int rc = execute_one(po, in, out, true);
  if (rc) fatal_or_return("offt_3d_execute_inverse");
  return rc;
}

int offtb_execute_group(struct _offt_plan **plans, double **arrays, int n, int inverse) {
  World &w = world();
  if (!w.up || !w.local) { set_error("offtb_execute_group needs a local world"); return -1; }
  if (n != w.size) { set_error("group of %d plans in a world of %d ranks", n, w.size); return -1; }
  std::vector<struct _offt_plan *> g(plans, plans + n);
  std::vector<double *> a(arrays, arrays + n);
  for (int r = 0; r < n; ++r)
    if (!g[r] || g[r]->rank != r) { set_error("plans[%d] must be the plan of rank %d", r, r); return -1; }
  int rc = engine_execute(g, a, inverse != 0);
  if (!rc && !eng(g[0])->async)
    for (auto *po : g) fill_timers(po);
  return rc;
}

// ---- options ---------------------------------------------------------------------------------
int offtb_set_default_precision(int bits) {
  if (bits != 64 && bits != 32) { set_error("precision must be 64 or 32"); return -1; }
  g_default_precision = bits;
  return 0;
}
int offtb_set_force_generic(int on) { g_force_generic = on ? 1 : 0; return 0; }
int offtb_plan_precision(const struct _offt_plan *po) { return eng(po) ? eng(po)->prec : -1; }
int offtb_plan_set_stream(struct _offt_plan *po, void *stream) {
  if (!eng(po)) { set_error("null plan"); return -1; }
  eng(po)->s_user = (cudaStream_t)stream;
  return 0;
}
int offtb_plan_set_async(struct _offt_plan *po, int is_async) {
  if (!eng(po)) { set_error("null plan"); return -1; }
  eng(po)->async = is_async != 0;
  return 0;
}
int offtb_plan_set_stage_timing(struct _offt_plan *po, int on) {
  if (!eng(po)) { set_error("null plan"); return -1; }
  eng(po)->stage_timing = on != 0;
  return 0;
}
long long offtb_plan_alloc_elems(const struct _offt_plan *po) { return eng(po) ? eng(po)->alloc : -1; }
int offtb_plan_last_launches(const struct _offt_plan *po) { return eng(po) ? eng(po)->launches : -1; }
double offtb_plan_last_ms(const struct _offt_plan *po) { return eng(po) ? eng(po)->last_ms : -1.0; }
int offtb_plan_stage_ms(const struct _offt_plan *po, double *ms, int n) {
  if (!eng(po)) { set_error("null plan"); return -1; }
  for (int i = 0; i < n && i < ST_COUNT; ++i) ms[i] = eng(po)->stage_ms[i];
  return n < ST_COUNT ? n : ST_COUNT;
}
long long offtb_exchange_block_elems(const struct _offt_plan *po, int phase, int myT) {
  if (!po || !po->comm) return -1;
  const struct _offt_comm *c = po->comm;
  return phase == 1 ? (long long)myT * c->M2 * c->M3 : (long long)c->M1 * c->M4 * myT;
}

// ---- the 1-D kernel on its own -------------------------------------------------------------------
// maps are 9 integers each: {off, nlo_count (0 = no split), n_hi, n_lo, B0, s0, B1, s1, s2}
double offtb_fft_launch_raw(const void *in, void *out, int n, int bits, int sign, long long nbatch,
                            const long long *im9, const long long *om9, int c_log, int load_cfast, int store_cfast,
                            int ry_level, int ry_x0, int ry_lo, int ry_hi, int repeat, void *stream) {
  FftKernelInfo info;
  const bool generic = g_force_generic > 0 || !fft_kernel_info(n, bits, &info);
  if (generic && (n < 1 || (size_t)n > fft_generic_max_n(bits))) { set_error("unsupported length %d", n); return -1.0; }
  auto mk = [](const long long *m) {
    FftMap f;
    f.off = m[0];
    int lg = 30;
    if (m[1] > 0) { lg = 0; while ((1LL << lg) < m[1]) ++lg; }
    f.n_lg = lg; f.n_hi = m[2]; f.n_lo = m[3];
    f.B0 = (unsigned)(m[4] > 0 ? m[4] : 1); f.s0 = m[5];
    f.B1 = (unsigned)(m[6] > 0 ? m[6] : 1); f.s1 = m[7]; f.s2 = m[8];
    f.gF = f.gb = f.gg = 0;
    return f;
  };
  const size_t esz = bits == 64 ? 16 : 8;
  static std::map<long long, void *> tw_cache;   // (length, precision, table kind) -> device table
  void *&tw = tw_cache[((long long)n << 2) | (bits == 64 ? 0 : 1) | (generic ? 2 : 0)];
  if (!tw) {
    std::vector<long double> tab(2 * (size_t)n + 2);
    const int count = generic ? fft_generic_twiddle_table(n, tab.data()) : fft_twiddle_table(n, bits, tab.data());
    const size_t cnt = (size_t)std::max(count, 1);
    std::vector<double> hd(2 * cnt);
    std::vector<float> hf(2 * cnt);
    for (size_t j = 0; j < 2 * (size_t)count; ++j) { hd[j] = (double)tab[j]; hf[j] = (float)tab[j]; }
    if (cudaMalloc(&tw, cnt * esz) != cudaSuccess) { set_error("cudaMalloc twiddles"); return -1.0; }
    cudaMemcpy(tw, bits == 64 ? (void *)hd.data() : (void *)hf.data(), cnt * esz, cudaMemcpyHostToDevice);
  }
  FftArgs a;
  memset(&a, 0, sizeof(a));
  a.in = in; a.out = out; a.tw = tw; a.im = mk(im9); a.om = mk(om9);
  a.load_cfast = load_cfast; a.store_cfast = store_cfast; a.conj = sign > 0;
  a.ry_level = ry_level; a.ry_x0 = ry_x0; a.ry_lo = ry_lo; a.ry_hi = ry_hi;
  if (!generic) {
    if (c_log < 0) c_log = fft_pick_c_log(info, bits, load_cfast || store_cfast, a.im.B0, nbatch, a.im.n_lo, a.om.n_lo);
    a.c_log = c_log;
    if ((info.T << c_log) > info.maxt) { set_error("c_log %d: %d threads exceed the kernel's bound %d", c_log, info.T << c_log, info.maxt); return -1.0; }
  }
  cudaStream_t st = (cudaStream_t)stream;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0, st);
  for (int r = 0; r < (repeat > 0 ? repeat : 1); ++r) {
    cudaError_t err = generic ? fft_generic_launch(n, bits, a, nbatch, st, nullptr) : fft_launch(n, bits, a, nbatch, st);
    if (err != cudaSuccess) { set_error("launch: %s", cudaGetErrorString(err)); return -1.0; }
  }
  cudaEventRecord(e1, st);
  cudaError_t err = cudaEventSynchronize(e1);
  if (err != cudaSuccess) { set_error("kernel: %s", cudaGetErrorString(err)); return -1.0; }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  return (double)ms / (repeat > 0 ? repeat : 1);
}

double offtb_fft_rows(void *data, int n, long long stride, long long dist, long long howmany, int sign, int bits,
                      int repeat, void *stream) {
  if (stride != 1 && dist != 1) { set_error("offtb_fft_rows: one of stride, dist must be 1"); return -1.0; }
  long long m[9] = {0, 0, 0, stride, howmany, dist, 1, 0, 0};
  const int cfast = (stride != 1);
  return offtb_fft_launch_raw(data, data, n, bits, sign, howmany, m, m, -1, cfast, cfast, -1, 0, 0, 10, repeat, stream);
}

// ---- tuning --------------------------------------------------------------------------------------
// ah_tuning (offt-tuning.c:744-1022) as offt_3d_init calls it (offt-compute.c:3440): the loop of tune.cu with the
// strategy the caller chose (-s: 0 nm, 1 pro, 2 random, 3 brute) - through the reference's own Active Harmony server
// where its back end was built (offt_b200/ah), through the built-in sources otherwise - over all decompositions (the
// caller lays out its array only after init returns, run-fft.c:269-304, 314)
int offtb_tune_harmony(struct _offt_plan *po, int max_loop, int verbose, int strategy, int search_p1);

int ah_tuning(struct _offt_plan *po, double *in, double *out) {
  (void)in; (void)out;
  return offtb_tune_harmony(po, po->max_loop, !po->rank, po->ah_strategy, 1);
}

}  // extern "C"
