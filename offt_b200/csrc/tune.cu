// Built-in search over the reference's tunables on the device.
//
// Keeps the shape of ah_tuning (offt-tuning.c:879-1006): candidate -> ADJUST_POINT repair
// (:90-118) -> feasibility (:144-226) -> point database lookup (:231-263) -> one measured
// execute per point (TUNING_REPS = 1, :966) -> report -> install the best point (:995-1006).
// The candidates come from a coordinate descent on the reference's value grid over the knobs
// that change the GPU schedule - tile thickness T (message size / launch count) and window W
// (ring depth) of each phase - instead of from the Active Harmony server; P1 stays what the
// plan was created with because the caller's array is already laid out for it.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <vector>

#include "engine.h"

namespace offtb {

namespace {

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

bool rebuild(struct _offt_plan *po, const int *v) {
  engine_destroy(po);
  memcpy(po->params->v, v, sizeof(int) * PARAM_COUNT);
  offt_comm_free(po->comm);
  po->comm = offt_comm_malloc(po);
  return engine_create(po) == 0;
}

}  // namespace

}  // namespace offtb

using namespace offtb;

extern "C" int offtb_tune(struct _offt_plan *po, double *in, double *out, int max_loop, int verbose) {
  if (!po || !po->b200) { set_error("null plan"); return -1; }
  if (world().local && world().size > 1) { set_error("tuning runs one rank per process"); return -1; }
  (void)in;
  const int Nx = po->Nx, Ny = po->Ny, Nz = po->is_r2c ? po->Nz / 2 + 1 : po->Nz, p = po->p;   // offt-tuning.c:110, 161
  const auto grid = params_grid(Nx, Ny, Nz, p);
  std::vector<int> best(po->params->v, po->params->v + PARAM_COUNT);
  params_adjust(Nx, Ny, Nz, p, po->is_oned, best.data());
  repair_ignored(Nx, Ny, Nz, p, best.data());
  std::map<std::vector<int>, double> database;   // the reference's tmp-db file, kept in memory
  int evaluated = 0;

  auto measure = [&](std::vector<int> v) -> double {
    params_adjust(Nx, Ny, Nz, p, po->is_oned, v.data());
    repair_ignored(Nx, Ny, Nz, p, v.data());
    int bad;
    // the ring-size rule of the reference is about MPI_Alloc_mem on its clusters; HBM has room
    // for far larger rings, so only the structural rules decide here
    if (params_infeasible(Nx, Ny, Nz, p, v.data(), &bad) && bad != _W1_ && bad != _W2_) return -1.0;
    if (v[_W1_] > (po->comm->M1 + v[_T1_] - 1) / v[_T1_] || v[_W2_] > (po->comm->M3 + v[_T2_] - 1) / v[_T2_]) return -1.0;
    auto it = database.find(v);
    if (it != database.end()) return it->second;
    if (!rebuild(po, v.data())) return -1.0;
    double t = 1e30;
    for (int rep = 0; rep < 1 + TUNING_REPS; ++rep) {   // one warm-up, then the measured run
      offt_3d_execute(po, out, out, 1);
      t = po->t[ALL];
    }
    t = agree_max(t);
    database[v] = t;
    ++evaluated;
    if (verbose) { printf("@ TUNE %.6f ", t); print_params(v.data()); }
    return t;
  };

  double best_t = measure(best);
  if (best_t < 0) { set_error("the starting point is infeasible"); return -1; }
  const int knobs[4] = {_T1_, _W1_, _T2_, _W2_};
  bool improved = true;
  while (improved && evaluated < max_loop) {
    improved = false;
    for (int k : knobs) {
      // a slab schedule has only one live phase
      Engine *E = (Engine *)po->b200;
      if ((k == _T1_ || k == _W1_) && E->sched == SCHED_SLAB_PX1) continue;
      if ((k == _T2_ || k == _W2_) && E->sched == SCHED_SLAB_1XP) continue;
      if (E->sched == SCHED_SINGLE) continue;
      const std::vector<int> &g = grid[k];
      auto pos = std::find(g.begin(), g.end(), best[k]);
      for (int dir : {+1, -1}) {
        std::vector<int> cur = best;
        auto q = pos;
        while (evaluated < max_loop) {
          if (q == g.end()) break;
          if (dir > 0) { if (q + 1 == g.end()) break; ++q; } else { if (q == g.begin()) break; --q; }
          cur[k] = *q;
          const double t = measure(cur);
          if (t < 0) break;
          if (t < best_t) { best_t = t; best = cur; improved = true; pos = q; } else break;
        }
      }
    }
  }
  params_adjust(Nx, Ny, Nz, p, po->is_oned, best.data());
  repair_ignored(Nx, Ny, Nz, p, best.data());
  if (!rebuild(po, best.data())) return -1;
  po->params->is_converged = 1;
  if (verbose) { printf("@ BEST %.6f ", best_t); print_params(po->params->v); }
  return evaluated;
}
