// Search over the reference's tunables on the device.
//
// Keeps the shape of ah_tuning (offt-tuning.c:744-1022): a candidate in grid-index space -> params_convert with its
// ADJUST_POINT repairs (:80-136) -> feasibility (:144-226) -> point database lookup (:231-263) -> build the layout and
// the engine for the point (:929-948) -> one measured execute on zeroed data (TUNING_REPS = 1, :958-966) -> report ->
// install the best point (:995-1006).  Infeasible points and database hits cost no measurement and do not count
// against max_loop; ten times max_loop fetches end the search regardless (:893).
//
// What proposes the candidates:
//   * Active Harmony itself, as in the reference, where its back end exists: offt_b200/ah/Makefile compiles the
//     reference's UNMODIFIED hserver, session-core, strategy plug-ins (the patched nm.so included) and client library
//     out of tree, and rank 0 drives a session exactly as ah_tuning does - 24 integer variables V00..V23 over grid
//     indices, the strategy of -s, the initial simplex through SHSONG_USER_VERTEX_FILE, a server launched on localhost
//     if none answers, fetch / report until max_loop or convergence, then the best point (search_active_harmony below,
//     offt_b200/ah/ah_glue.c; offt-tuning.c:773-854, 893-913, 988-1019).  OFFTB_TUNER=builtin skips it, OFFTB_TUNER=ah
//     insists on it.
//   * otherwise (no reference tree at build time, no TCP, ...) the built-in sources of the same shape:
//   strategy 0 / 1  Nelder-Mead over all 24 index coordinates, started from the reference's own initial simplex
//                   (write_initial_simplex, offt-tuning.c:426-737; its patched nm.so does the same, strategies/nm.c:369-396)
//   strategy 2      uniformly random grid points (random.so)
//   strategy 3      coordinate descent along the tunables that change the GPU schedule, from the default point
// Every rank runs the same deterministic search on times agreed over the world (the slowest rank's), so no broadcast
// of the point is needed.  P1 is searched like any other tunable when the caller allows it: every trial rebuilds the
// layout descriptor and runs on an internal zeroed device array, like the reference's memset(in, 0) (:958).
//
// Tunables that do nothing on a GPU - the CPU cache sub-tile sizes and the MPI_Test frequencies - stay in the point
// (the reference's space has 24 coordinates) but are pulled into range before the feasibility test and left out of the
// database key, so that points differing only in them are measured once.
#include <dlfcn.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "engine.h"

namespace offtb {

extern int g_default_precision;
static int g_default_precision_of(const struct _offt_plan *po) {
  return po && po->b200 ? ((Engine *)po->b200)->prec : g_default_precision;
}

namespace {

const int kLive[] = {_P1_, _T1_, _W1_, _Ry_, _T2_, _W2_, _V_, _S_};   // what changes the schedule, the layout or the exchange
const double kInfeasible = 99999999.0;                                // what the reference reports for such points (:906)

// every rank must take the same decisions: agree on the slowest rank's time
double agree_max(double x) {
  World &w = world();
  if (!w.nccl) return x;
  static double *d = nullptr;
  if (!d) cudaMalloc(&d, sizeof(double));
  cudaMemcpy(d, &x, sizeof(double), cudaMemcpyHostToDevice);
  nccl_api()->AllReduce(d, d, 1, ncclDouble, ncclMax, w.nccl, 0);
  cudaMemcpy(&x, d, sizeof(double), cudaMemcpyDeviceToHost);
  return x;
}

// The CPU cache sub-tile sizes and MPI_Test frequencies are accepted and ignored by this implementation, but the
// reference's feasibility rules tie them to T (e.g. Pz2 <= T2, offt-tuning.c:178-205): a search that moves T would
// be blocked by knobs that do nothing here.  Pull them into range instead (the same spirit as ADJUST_POINT,
// offt-tuning.c:90-118), so that feasibility is decided by P1, T and W alone.
void repair_ignored(int Nx, int Ny, int Nz, int p, int *v) {
  const int p1 = v[_P1_] > 0 ? v[_P1_] : 1, p2 = p / p1 > 0 ? p / p1 : 1;
  auto cd = [](int a, int b) { return (a + b - 1) / b; };
  const int M1 = cd(Nx, p1), M2 = cd(Ny, p2), M3 = cd(Nz, p2), M4 = cd(Ny, p1);
  auto clamp = [&](int i, int hi) { v[i] = std::max(1, std::min(v[i], std::max(hi, 1))); };
  clamp(_Px1_, v[_T1_]); clamp(_Py1_, M2); clamp(_Ux1_, v[_T1_]); clamp(_Uz1_, M3);
  clamp(_Px2_, M1); clamp(_Pz2_, v[_T2_]); clamp(_Uy2_, M4); clamp(_Uz2_, v[_T2_]);
  for (int i : {_Fz_, _FP1_, _FU1_, _Fy1_, _Fy2_, _FP2_, _FU2_, _Fx_}) v[i] = 0;
}

struct Search {
  struct _offt_plan *po;
  int Nx, Ny, Nz, p;         // Nz: the z extent the tunables are sized by (Nz/2+1 for real-to-complex plans)
  bool search_p1, verbose;
  int max_loop;
  std::vector<std::vector<int>> grid;
  std::map<std::vector<int>, double> database;   // the reference's tmp-db file, kept in memory, keyed by the live tunables
  int measured = 0, fetched = 0;
  std::vector<int> best_v;
  double best_t = 1e300;
  int fixed_p1, fixed_S = 0;
  bool search_layout = false;   // the decomposition P1 and the output layout S may move (nobody has read po->comm yet)
  bool dead = false;         // a rebuild failed in a way the search cannot continue from

  bool budget_left() const { return !dead && measured < max_loop && fetched < 10 * max_loop; }

  std::vector<int> values_of(const std::vector<int> &idx) const {
    std::vector<int> v(PARAM_COUNT);
    for (int i = 0; i < PARAM_COUNT; ++i) {
      const int n = (int)grid[i].size();
      v[i] = grid[i][std::min(std::max(idx[i], 0), n - 1)];
    }
    return v;
  }
  std::vector<int> index_of(const std::vector<int> &v) const {
    std::vector<int> idx(PARAM_COUNT, 0);
    for (int i = 0; i < PARAM_COUNT; ++i) {
      int bi = 0;
      for (int c = 0; c < (int)grid[i].size(); ++c)
        if (std::abs(grid[i][c] - v[i]) < std::abs(grid[i][bi] - v[i])) bi = c;
      idx[i] = bi;
    }
    return idx;
  }

  bool rebuild(const int *v) {
    engine_destroy(po);
    memcpy(po->params->v, v, sizeof(int) * PARAM_COUNT);
    offt_comm_free(po->comm);
    po->comm = offt_comm_malloc(po);
    return engine_create(po) == 0;
  }

  // what a candidate becomes before it is tested: pinned layout knobs, ADJUST_POINT, the dead knobs pulled into range
  std::vector<int> repaired(std::vector<int> v) const {
    if (!search_p1) v[_P1_] = fixed_p1;
    if (!search_layout) v[_S_] = fixed_S;   // the caller has read the output strides already
    params_adjust(Nx, Ny, Nz, p, po->is_oned, v.data());
    repair_ignored(Nx, Ny, Nz, p, v.data());
    if (po->is_W0) v[_W1_] = v[_W2_] = 0;
    return v;
  }
  static std::vector<int> key_of(const std::vector<int> &v) {
    std::vector<int> key;
    for (int k : kLive) key.push_back(v[k]);
    return key;
  }

  // one point: repaired, tested, looked up, measured.  Returns its time in seconds (kInfeasible if it cannot run).
  double evaluate(std::vector<int> v) {
    ++fetched;
    v = repaired(v);
    int bad;
    // the ring-size rule of the reference is about MPI_Alloc_mem on its clusters; HBM has room for far larger rings,
    // so only the structural rules decide here - and the memory actually free on the device (below)
    if (params_infeasible(Nx, Ny, Nz, p, v.data(), &bad) && bad != _W1_ && bad != _W2_) {
      if (verbose) { printf("INFEASIBLE POINT err_i:%d ", bad); print_params(v.data()); }
      return kInfeasible;
    }
    auto cd = [](int a, int b) { return (a + b - 1) / b; };
    const int p1 = v[_P1_], p2 = p / p1;
    const long long M1 = cd(Nx, p1), M2 = cd(Ny, p2), M3 = cd(Nz, p2), M4 = cd(Ny, p1);
    if (v[_W1_] > cd((int)M1, v[_T1_]) || v[_W2_] > cd((int)M3, v[_T2_])) return kInfeasible;
    const std::vector<int> key = key_of(v);
    auto it = database.find(key);
    if (it != database.end()) {
      if (verbose) { printf("%.5f FOUND IN DATABASE ", it->second); print_params(v.data()); }
      return it->second;
    }
    // a cheap upper bound on what the point allocates (rings, scratch, the trial array) against the memory that is
    // free right now: a point that cannot fit is infeasible for everybody, before any collective is entered
    {
      const size_t esz = g_default_precision_of(po) == PREC_F64 ? 16 : 8;
      const bool ph1 = !(po->is_oned && p1 == p) && p > 1, ph2 = !(po->is_oned && p1 == 1) && p > 1;
      const long long ring = (ph1 ? 2LL * v[_T1_] * M2 * M3 * p2 * (v[_W1_] + 1) : 0) + (ph2 ? 2LL * M1 * M4 * p1 * v[_T2_] * (v[_W2_] + 1) : 0);
      const long long arrays = 2 * std::max(M2 * p2, M4 * p1) * M1 * M3;
      size_t free_b = 0, total_b = 0;
      cudaMemGetInfo(&free_b, &total_b);
      Engine *cur = (Engine *)po->b200;
      size_t mine = 0;   // what the current engine holds is released before the point is built
      if (cur) mine = ((size_t)cur->alloc * (cur->d_scratch ? 1 : 0) + (size_t)(cur->ring[0].slot_elems * 2 * cur->ring[0].depth + cur->ring[1].slot_elems * 2 * cur->ring[1].depth)) * esz;
      const double fits = agree_max((double)((size_t)(ring + arrays) * esz > free_b + mine ? 1 : 0));
      if (fits > 0) {
        if (verbose) { printf("INFEASIBLE POINT (device memory) "); print_params(v.data()); }
        database[key] = kInfeasible;
        return kInfeasible;
      }
    }
    if (!rebuild(v.data())) {
      // all ranks fail together (engine_create agrees on allocation failures): the point is infeasible, the search goes on
      if (verbose) { printf("INFEASIBLE POINT (%s) ", last_error()); print_params(v.data()); }
      database[key] = kInfeasible;
      if (!po->b200 && !rebuild(best_v.empty() ? v.data() : best_v.data()) && best_v.empty()) dead = true;
      return kInfeasible;
    }
    Engine *E = (Engine *)po->b200;
    void *trial = nullptr;
    const size_t bytes = (size_t)E->alloc * E->esz;
    const double ok = agree_max(cudaMalloc(&trial, bytes) == cudaSuccess ? 0.0 : 1.0);
    if (ok > 0) {
      cudaGetLastError();
      if (trial) cudaFree(trial);
      database[key] = kInfeasible;
      return kInfeasible;
    }
    double t = 1e30;
    const bool was_timing = E->stage_timing;
    E->stage_timing = false;
    for (int rep = 0; rep < 1 + TUNING_REPS; ++rep) {   // one warm-up, then the measured run(s), best of them (:966)
      cudaMemset(trial, 0, bytes);                       // the reference tunes on zeros too (:958)
      offtb_world_barrier();
      offt_3d_execute(po, (double *)trial, (double *)trial, 1);
      if (rep > 0) t = std::min(t, po->t[ALL]);
    }
    E->stage_timing = was_timing;
    cudaFree(trial);
    t = agree_max(t);
    database[key] = t;
    ++measured;
    if (verbose) { printf("@ TUNE %.6f ", t); print_params(v.data()); }
    if (t < best_t) { best_t = t; best_v = v; }
    return t;
  }
};

// ---- strategy 0 / 1: Nelder-Mead in index space from the reference's initial simplex -----------------------------
void search_nelder_mead(Search &S) {
  const int D = PARAM_COUNT, NV = PARAM_COUNT + 1;
  std::vector<int *> v_list(D);
  std::vector<int> v_size(D);
  for (int i = 0; i < D; ++i) { v_list[i] = S.grid[i].data(); v_size[i] = (int)S.grid[i].size(); }
  int x0[PARAM_COUNT + 1][PARAM_COUNT];
  srand(20161);   // the same vertices on every rank (the reference draws them on rank 0 and broadcasts the points)
  initial_simplex(S.Nx, S.Ny, S.Nz, S.p, S.po->is_oned, S.po->is_W0, S.po->is_notest, S.search_p1 ? S.po->tuning_mode : 0, v_list.data(), v_size.data(), x0);
  std::vector<std::vector<double>> X(NV, std::vector<double>(D));
  std::vector<double> F(NV);
  auto eval = [&](const std::vector<double> &x) {
    std::vector<int> idx(D);
    for (int i = 0; i < D; ++i) idx[i] = (int)std::lround(std::min(std::max(x[i], 0.0), (double)(v_size[i] - 1)));
    return S.evaluate(S.values_of(idx));
  };
  for (int k = 0; k < NV && S.budget_left(); ++k) {
    for (int i = 0; i < D; ++i) X[k][i] = x0[k][i];
    F[k] = eval(X[k]);
  }
  std::vector<int> order(NV);
  while (S.budget_left()) {
    for (int k = 0; k < NV; ++k) order[k] = k;
    std::sort(order.begin(), order.end(), [&](int a, int b) { return F[a] < F[b]; });
    const int lo = order[0], hi = order[NV - 1], nhi = order[NV - 2];
    // converged: every vertex rounds to the same point
    bool same = true;
    for (int k = 1; k < NV && same; ++k)
      for (int i = 0; i < D && same; ++i) same = std::lround(X[order[k]][i]) == std::lround(X[lo][i]);
    if (same) break;
    std::vector<double> c(D, 0.0);
    for (int k = 0; k < NV; ++k) if (k != hi) for (int i = 0; i < D; ++i) c[i] += X[k][i] / (NV - 1);
    auto along = [&](double a) { std::vector<double> y(D); for (int i = 0; i < D; ++i) y[i] = c[i] + a * (X[hi][i] - c[i]); return y; };
    std::vector<double> xr = along(-1.0);
    const double fr = eval(xr);
    if (fr < F[lo]) {
      std::vector<double> xe = along(-2.0);
      const double fe = S.budget_left() ? eval(xe) : kInfeasible;
      if (fe < fr) { X[hi] = xe; F[hi] = fe; } else { X[hi] = xr; F[hi] = fr; }
    } else if (fr < F[nhi]) {
      X[hi] = xr; F[hi] = fr;
    } else {
      std::vector<double> xc = along(fr < F[hi] ? -0.5 : 0.5);
      const double fc = S.budget_left() ? eval(xc) : kInfeasible;
      if (fc < std::min(fr, F[hi])) { X[hi] = xc; F[hi] = fc; }
      else {
        for (int k = 0; k < NV && S.budget_left(); ++k) {   // shrink towards the best vertex
          if (k == lo) continue;
          for (int i = 0; i < D; ++i) X[k][i] = X[lo][i] + 0.5 * (X[k][i] - X[lo][i]);
          F[k] = eval(X[k]);
        }
      }
    }
  }
}

// ---- strategy 2: random grid points -------------------------------------------------------------------------------
void search_random(Search &S) {
  unsigned long long state = 0x9E3779B97F4A7C15ULL;
  auto next = [&]() { state = state * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(state >> 33); };
  std::vector<int> v(PARAM_COUNT);
  S.evaluate(std::vector<int>(S.po->params->v, S.po->params->v + PARAM_COUNT));
  while (S.budget_left()) {
    for (int i = 0; i < PARAM_COUNT; ++i) v[i] = S.grid[i][next() % S.grid[i].size()];
    S.evaluate(v);
  }
}

// ---- strategy 3: coordinate descent along the live tunables ---------------------------------------------------------
void search_coordinate(Search &S) {
  std::vector<int> best(S.po->params->v, S.po->params->v + PARAM_COUNT);
  double best_t = S.evaluate(best);
  if (best_t >= kInfeasible) return;
  bool improved = true;
  while (improved && S.budget_left()) {
    improved = false;
    for (int k : kLive) {
      if (k == _P1_ && !S.search_p1) continue;
      if (k == _S_ && !S.search_layout) continue;
      if (k == _V_) continue;   // exact counts change nothing a fused exchange can measure
      const std::vector<int> &g = S.grid[k];
      auto pos = std::find(g.begin(), g.end(), best[k]);
      if (pos == g.end()) pos = g.begin() + S.index_of(best)[k];
      for (int dir : {+1, -1}) {
        std::vector<int> cur = best;
        auto q = pos;
        while (S.budget_left()) {
          if (dir > 0) { if (q + 1 == g.end()) break; ++q; } else { if (q == g.begin()) break; --q; }
          cur[k] = *q;
          if (k == _P1_) {   // another decomposition: start from that decomposition's default tiles (M/16, window 2; offt-compute.c:3147-3175)
            const int p1 = cur[_P1_], p2 = S.p / p1;
            auto snap = [&](int knob, int raw) { int b = S.grid[knob][0]; for (int c : S.grid[knob]) if (c <= raw) b = c; return b; };
            cur[_T1_] = snap(_T1_, std::max(1, (S.Nx + p1 - 1) / p1 / 16));
            cur[_T2_] = snap(_T2_, std::max(1, (S.Nz + p2 - 1) / p2 / 16));
            cur[_W1_] = cur[_W2_] = S.po->is_W0 ? 0 : 2;
            cur[_Ry_] = 5;
          }
          const double t = S.evaluate(cur);
          if (t >= kInfeasible) { if (k == _P1_) continue; break; }
          if (t < best_t) { best_t = t; best = S.best_v.empty() ? cur : S.best_v; improved = true; pos = q; } else break;
        }
      }
    }
  }
}

// ---- Active Harmony: the reference's own search server ---------------------------------------------------------------
struct AhApi {
  void *handle = nullptr;
  std::string root;
  int (*open)(const char *, int, const int *, int, const char *, int) = nullptr;
  int (*fetch)(long *) = nullptr;
  int (*report)(double) = nullptr;
  int (*converged)() = nullptr;
  int (*best)(long *) = nullptr;
  void (*close)() = nullptr;
  const char *(*error)() = nullptr;
};

// <dir of libofft_b200.so>/../ah/_root, or OFFTB_AH_ROOT
bool ah_load(AhApi &A) {
  std::string root;
  if (const char *e = getenv("OFFTB_AH_ROOT")) root = e;
  else {
    Dl_info info;
    if (!dladdr((void *)&ah_load, &info) || !info.dli_fname) return false;
    std::string lib = info.dli_fname;
    const size_t slash = lib.rfind('/');
    root = (slash == std::string::npos ? std::string(".") : lib.substr(0, slash)) + "/../ah/_root";
  }
  const std::string so = root + "/lib/libofft_ah.so";
  if (access(so.c_str(), R_OK) != 0) return false;
  A.handle = dlopen(so.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!A.handle) return false;
  A.root = root;
  A.open = (int (*)(const char *, int, const int *, int, const char *, int))dlsym(A.handle, "offtb_ah_open");
  A.fetch = (int (*)(long *))dlsym(A.handle, "offtb_ah_fetch");
  A.report = (int (*)(double))dlsym(A.handle, "offtb_ah_report");
  A.converged = (int (*)())dlsym(A.handle, "offtb_ah_converged");
  A.best = (int (*)(long *))dlsym(A.handle, "offtb_ah_best");
  A.close = (void (*)())dlsym(A.handle, "offtb_ah_close");
  A.error = (const char *(*)())dlsym(A.handle, "offtb_ah_error");
  return A.open && A.fetch && A.report && A.converged && A.best && A.close && A.error;
}

// rank 0's integers to every rank (the reference's MPI_Bcast of the point, offt-tuning.c:920): a sum in which the
// other ranks contribute zeros
void bcast_ints(int *v, int n) {
  World &w = world();
  if (!w.nccl) return;
  static int *d = nullptr;
  static int cap = 0;
  if (cap < n) { if (d) cudaFree(d); cudaMalloc(&d, sizeof(int) * n); cap = n; }
  std::vector<int> h(v, v + n);
  if (w.rank != 0) std::fill(h.begin(), h.end(), 0);
  cudaMemcpy(d, h.data(), sizeof(int) * n, cudaMemcpyHostToDevice);
  nccl_api()->AllReduce(d, d, n, ncclInt, ncclSum, w.nccl, 0);
  cudaMemcpy(v, d, sizeof(int) * n, cudaMemcpyDeviceToHost);
}

// Returns false if the server could not be reached on rank 0 (every rank learns it), so that the caller falls back.
bool search_active_harmony(Search &S, int strategy) {
  World &w = world();
  const bool lead = w.rank == 0;
  AhApi A;
  int ok = 0;
  char vertex_file[64] = "";
  if (lead && ah_load(A)) {
    std::vector<int> sizes(PARAM_COUNT);
    std::vector<int *> v_list(PARAM_COUNT);
    for (int i = 0; i < PARAM_COUNT; ++i) { sizes[i] = (int)S.grid[i].size(); v_list[i] = S.grid[i].data(); }
    const char *vf = nullptr;
    if (strategy == 0 || strategy == 1) {   // only nm and pro take the user simplex (offt-tuning.c:795-796)
      int x0[PARAM_COUNT + 1][PARAM_COUNT];
      srand(20161);
      initial_simplex(S.Nx, S.Ny, S.Nz, S.p, S.po->is_oned, S.po->is_W0, S.po->is_notest, S.search_p1 ? S.po->tuning_mode : 0, v_list.data(), sizes.data(), x0);
      snprintf(vertex_file, sizeof(vertex_file), "/tmp/offtb-uv-%d", (int)getpid());
      if (FILE *f = fopen(vertex_file, "w")) {
        for (int i = 0; i < PARAM_COUNT + 1; ++i) {
          // a decomposition that is not being searched stays where it is in every vertex
          if (!S.search_p1) x0[i][_P1_] = S.index_of(std::vector<int>(S.po->params->v, S.po->params->v + PARAM_COUNT))[_P1_];
          for (int j = 0; j < PARAM_COUNT; ++j) fprintf(f, "%d ", x0[i][j]);
          fprintf(f, "\n");
        }
        fclose(f);
        vf = vertex_file;
      }
    }
    const int port = getenv("HARMONY_S_PORT") ? 0 : 20000 + (int)(getpid() % 20000);
    if (A.open(A.root.c_str(), PARAM_COUNT, sizes.data(), strategy, vf, port) == 0) ok = 1;
    else if (S.verbose) printf("Active Harmony not available (%s): using the built-in search\n", A.error());
  }
  bcast_ints(&ok, 1);
  if (!ok) { if (lead && vertex_file[0]) remove(vertex_file); return false; }
  if (lead && S.verbose) printf("Starting Harmony...\n");
  while (true) {
    int msg[1 + PARAM_COUNT] = {0};   // [0]: 1 = converged / out of budget
    if (lead) {
      if (!S.budget_left() || A.converged() == 1) msg[0] = 1;
      else {
        long idx[PARAM_COUNT];
        if (A.fetch(idx) < 0) msg[0] = 1;
        for (int i = 0; i < PARAM_COUNT; ++i) msg[1 + i] = (int)idx[i];
      }
    }
    bcast_ints(msg, 1 + PARAM_COUNT);
    if (msg[0]) break;
    const double perf = S.evaluate(S.values_of(std::vector<int>(msg + 1, msg + 1 + PARAM_COUNT)));
    if (lead) A.report(perf);
  }
  // the server's best point (harmony_best, offt-tuning.c:995-1000) is installed if it was measured and is no slower
  // than the best measurement this loop saw (the two differ only when equal keys or repaired points are involved)
  int bmsg[1 + PARAM_COUNT] = {0};
  if (lead) {
    long idx[PARAM_COUNT];
    if (A.best(idx) >= 0) { bmsg[0] = 1; for (int i = 0; i < PARAM_COUNT; ++i) bmsg[1 + i] = (int)idx[i]; }
    A.close();
    if (vertex_file[0]) remove(vertex_file);
  }
  bcast_ints(bmsg, 1 + PARAM_COUNT);
  if (bmsg[0]) {
    const std::vector<int> b = S.repaired(S.values_of(std::vector<int>(bmsg + 1, bmsg + 1 + PARAM_COUNT)));
    auto it = S.database.find(Search::key_of(b));
    if (lead && S.verbose) { printf("@ HARMONY BEST "); print_params(const_cast<int *>(b.data())); }
    if (it != S.database.end() && it->second < kInfeasible && it->second <= S.best_t) { S.best_t = it->second; S.best_v = b; }
  }
  return true;
}

}  // namespace

}  // namespace offtb

using namespace offtb;

// strategy: 0 / 1 Nelder-Mead from the reference's initial simplex, 2 random, 3 coordinate descent (see the header
// comment); search_p1: also search what changes the caller's layout - the decomposition P1 and the output order S
// (the caller must not have read po->comm yet, as in the reference where tuning happens inside offt_3d_init).
// Returns the number of points measured, or a negative code.
static int tune_impl(struct _offt_plan *po, int max_loop, int verbose, int strategy, int search_p1, int use_ah) {
  if (!po || !po->b200) { set_error("null plan"); return -1; }
  if (world().local && world().size > 1) { set_error("tuning runs one rank per process"); return -1; }
  Search S;
  S.po = po;
  S.Nx = po->Nx; S.Ny = po->Ny; S.Nz = po->is_r2c ? po->Nz / 2 + 1 : po->Nz; S.p = po->p;   // offt-tuning.c:110, 161
  S.search_p1 = search_p1 != 0 && po->p > 1;
  S.search_layout = search_p1 != 0;
  S.fixed_S = po->params->v[_S_];
  S.verbose = verbose != 0;
  S.max_loop = max_loop;
  S.grid = params_grid(S.Nx, S.Ny, S.Nz, S.p);
  S.fixed_p1 = po->params->v[_P1_];
  if (S.search_p1 && po->tuning_mode == 1) { S.search_p1 = false; S.fixed_p1 = 1; }
  if (S.search_p1 && po->tuning_mode == 2) { S.search_p1 = false; S.fixed_p1 = po->p; }
  const std::vector<int> start(po->params->v, po->params->v + PARAM_COUNT);
  if (((Engine *)po->b200)->sched == SCHED_SINGLE && !S.search_p1) max_loop = S.max_loop = std::min(max_loop, 2);   // one rank: nothing to search but S
  // the reference's own search server where its back end was built (strategies 0..3 = nm, pro, random, brute); the
  // built-in sources otherwise.  Strategy 3 maps to the built-in coordinate descent unless Active Harmony is asked for.
  const char *tuner = getenv("OFFTB_TUNER");
  const bool want_ah = tuner ? strcmp(tuner, "ah") == 0 : (use_ah != 0 && strategy != 3);
  bool done = false;
  if (want_ah && world().up && !(world().local && world().size > 1)) done = search_active_harmony(S, strategy);
  if (!done && tuner && strcmp(tuner, "ah") == 0) { set_error("OFFTB_TUNER=ah: the Active Harmony back end is not available (offt_b200/ah/_root)"); return -1; }
  if (!done) {
    switch (strategy) {
      case 0: case 1: search_nelder_mead(S); break;
      case 2: search_random(S); break;
      default: search_coordinate(S); break;
    }
  }
  if (S.best_v.empty()) {
    // nothing feasible was measured: put the plan back where it started
    std::vector<int> v = start;
    if (!po->b200 && !S.rebuild(v.data())) { set_error("tuning found no feasible point and could not restore the plan: %s", last_error()); return -1; }
    set_error("the search found no feasible point");
    return -1;
  }
  if (!S.rebuild(S.best_v.data())) return -1;
  po->params->is_converged = 1;
  if (verbose) { printf("@ BEST "); print_params(po->params->v); printf("@ BEST %.5f\n", S.best_t); }
  return S.measured;
}

// trials run on an internal zeroed device array (offt-tuning.c:958 zeroes the caller's), so in / out are not touched
extern "C" int offtb_tune_ex(struct _offt_plan *po, double *in, double *out, int max_loop, int verbose, int strategy, int search_p1) {
  (void)in; (void)out;
  return tune_impl(po, max_loop, verbose, strategy, search_p1, 0);
}

// what ah_tuning calls: like offtb_tune_ex, with the Active Harmony server as the candidate source where available
extern "C" int offtb_tune_harmony(struct _offt_plan *po, int max_loop, int verbose, int strategy, int search_p1) {
  return tune_impl(po, max_loop, verbose, strategy, search_p1, 1);
}

extern "C" int offtb_tune(struct _offt_plan *po, double *in, double *out, int max_loop, int verbose) {
  return offtb_tune_ex(po, in, out, max_loop, verbose, 3, 0);
}
