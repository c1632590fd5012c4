// The reference's tunable-parameter space, host side only (no GPU needed).
//
// Same 24 tunables with the same meaning, value grids, default heuristics and
// feasibility rules as the hooks offt exposes to Active Harmony:
//   params_range_setup   offt-compute.c:2998-3093
//   grid_value_floor/ceil offt-compute.c:3096-3125
//   params_set_default   offt-compute.c:3127-3225
//   set_params_custom    offt-compute.c:3227-3234
//   print_params / offt_print_time  offt-compute.c:3239-3294
//   params_convert (ADJUST_POINT)   offt-tuning.c:90-118
//   is_infeasible_point  offt-tuning.c:144-226
// Written table-driven from those rules; pinned against the unmodified reference functions themselves
// (oracle/ref_hooks.c -> tests/test_hooks_vs_reference.py) and against the oracle (tests/test_oracle.py).
#define OFFT_NO_MINMAX
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "engine.h"

namespace offtb {

namespace {

enum Axis { AX_NONE, AX_X, AX_Y, AX_Z, AX_XY, AX_XZ, AX_YZ };
enum Kind { K_DIVISORS, K_SMALL_RANGE, K_POW2, K_POW2_ZERO };

struct ParamSpec {
  const char *name;
  Kind kind;
  int arg;  // K_SMALL_RANGE: number of values; K_POW2*: Axis of the upper bound
};

// index order = offt.h:74-98
const ParamSpec kSpec[PARAM_COUNT] = {
    {"P1", K_DIVISORS, 0},        {"T1", K_POW2, AX_X},        {"W1", K_SMALL_RANGE, 11},   {"Px1", K_POW2, AX_X},
    {"Py1", K_POW2, AX_Y},        {"Fz", K_POW2_ZERO, AX_XY},  {"FP1", K_POW2_ZERO, AX_XY}, {"Ux1", K_POW2, AX_X},
    {"Uz1", K_POW2, AX_Z},        {"FU1", K_POW2_ZERO, AX_XZ}, {"Fy1", K_POW2_ZERO, AX_XZ}, {"Ry", K_SMALL_RANGE, 11},
    {"T2", K_POW2, AX_Z},         {"W2", K_SMALL_RANGE, 11},   {"Pz2", K_POW2, AX_Z},       {"Px2", K_POW2, AX_X},
    {"Fy2", K_POW2_ZERO, AX_XZ},  {"FP2", K_POW2_ZERO, AX_XZ}, {"Uz2", K_POW2, AX_Z},       {"Uy2", K_POW2, AX_Y},
    {"FU2", K_POW2_ZERO, AX_YZ},  {"Fx", K_POW2_ZERO, AX_YZ},  {"V", K_SMALL_RANGE, 4},     {"S", K_SMALL_RANGE, 2},
};

int axis_extent(int ax, int Nx, int Ny, int Nz) {
  switch (ax) {
    case AX_X: return Nx;
    case AX_Y: return Ny;
    case AX_Z: return Nz;
    case AX_XY: return Nx * Ny;
    case AX_XZ: return Nx * Nz;
    case AX_YZ: return Ny * Nz;
    default: return 1;
  }
}

int snap_down(const std::vector<int> &grid, int raw) {
  for (size_t j = grid.size(); j-- > 0;)
    if (grid[j] <= raw) return grid[j];
  return raw;
}

const int kFreq[] = {_Fz_, _FP1_, _FU1_, _Fy1_, _Fy2_, _FP2_, _FU2_, _Fx_};

}  // namespace

std::vector<std::vector<int>> params_grid(int Nx, int Ny, int Nz, int p) {
  std::vector<std::vector<int>> g(PARAM_COUNT);
  for (int i = 0; i < PARAM_COUNT; ++i) {
    std::vector<int> &L = g[i];
    const ParamSpec &s = kSpec[i];
    if (s.kind == K_DIVISORS) {
      // divisors of p that leave every rank at least one plane on each split axis
      const int hi = std::min(std::min(Nx, Ny), p);
      const int lo = std::max(std::max(p / Nz, p / Ny), 1);
      for (int d = lo; d <= hi; ++d)
        if (p % d == 0) L.push_back(d);
    } else if (s.kind == K_SMALL_RANGE) {
      for (int c = 0; c < s.arg; ++c) L.push_back(c);
    } else {
      const int vmax = axis_extent(s.arg, Nx, Ny, Nz);
      if (s.kind == K_POW2_ZERO) L.push_back(0);
      int pw = 1;
      while (true) {
        L.push_back(pw);
        if (pw > vmax / 2) break;
        pw *= 2;
      }
      if (L.back() < vmax) L.push_back(vmax);
    }
  }
  return g;
}

void params_default(int Nx, int Ny, int Nz, int p, int is_W0, int is_notest, int *v) {
  const auto grid = params_grid(Nx, Ny, Nz, p);
  auto put = [&](int i, int raw) { v[i] = snap_down(grid[i], raw); };
  auto clamp = [](int x, int lo, int hi) { return std::min(std::max(x, lo), hi); };
  auto isqrt = [](int n) { return (int)std::sqrt((double)n); };
  auto ceil_div = [](int a, int b) { return (a + b - 1) / b; };

  put(_P1_, isqrt(p));
  const int p1 = v[_P1_], p2 = p / p1;
  const int M1 = ceil_div(Nx, p1), M2 = ceil_div(Ny, p2), M3 = ceil_div(Nz, p2), M4 = ceil_div(Ny, p1);
  const int cache = SUBTILE_SIZE;  // complex numbers per read/write sub-tile on the reference's CPUs

  // phase 1: x tiles, a window of two tiles in flight
  put(_T1_, std::max(M1 / 16, 1));
  put(_W1_, std::min(2, ceil_div(M1, v[_T1_])));
  const int budget_xy = cache / Nz;
  put(_Px1_, clamp(isqrt(budget_xy), 1, v[_T1_]));
  put(_Py1_, clamp(budget_xy / v[_Px1_], 1, M2));
  put(_Fz_, clamp(p2 / 2, 0, v[_T1_] * M2));
  put(_FP1_, clamp(v[_Fz_], 0, v[_T1_] / v[_Px1_] * M2 / v[_Py1_]));
  const int budget_xz = cache / Ny;
  put(_Ux1_, clamp(isqrt(budget_xz), 1, v[_T1_]));
  put(_Uz1_, clamp(budget_xz / v[_Ux1_], 1, M3));
  put(_FU1_, clamp(v[_Fz_], 0, v[_T1_] / v[_Ux1_] * M3 / v[_Uz1_]));
  put(_Fy1_, clamp(v[_Fz_], 0, v[_T1_] * M3));
  v[_Ry_] = 5;

  // phase 2: z tiles
  put(_T2_, std::max(M3 / 16, 1));
  put(_W2_, std::min(2, ceil_div(M3, v[_T2_])));
  put(_Pz2_, clamp(isqrt(budget_xz), 1, v[_T2_]));
  put(_Px2_, clamp(budget_xz / v[_Pz2_], 1, M1));
  put(_Fy2_, clamp(p1 / 2, 0, v[_T2_] * M1));
  put(_FP2_, clamp(v[_Fy2_], 0, M1 / v[_Px2_] * v[_T2_] / v[_Pz2_]));
  const int budget_yz = cache / Nx;
  put(_Uz2_, clamp(isqrt(budget_yz), 1, v[_T2_]));
  put(_Uy2_, clamp(budget_yz / v[_Uz2_], 1, M4));
  put(_FU2_, clamp(v[_FP2_], 0, M4 / v[_Uy2_] * v[_T2_] / v[_Uz2_]));
  put(_Fx_, clamp(v[_FP2_], 0, v[_T2_] * M4));

  v[_V_] = 0;
  v[_S_] = 0;
  if (is_W0) v[_W1_] = v[_W2_] = 0;
  if (is_W0 || is_notest)
    for (int f : kFreq) v[f] = 0;
}

int params_infeasible(int Nx, int Ny, int Nz, int p, const int *v, int *bad) {
  auto fail = [&](int i) { *bad = i; return 1; };
  auto ceil_div = [](int a, int b) { return (a + b - 1) / b; };
  *bad = -1;
  const int p1 = v[_P1_];
  if (p1 < 1 || p % p1 != 0) return fail(_P1_);
  const int p2 = p / p1;
  const int M1 = ceil_div(Nx, p1), M2 = ceil_div(Ny, p2), M3 = ceil_div(Nz, p2), M4 = ceil_div(Ny, p1);
  if (p1 > Nx || p1 > Ny || p2 > Ny || p2 > Nz) return fail(_P1_);
  // tile, window and ring-buffer budget of each phase (2: complex, 2: send + receive)
  struct Phase { int T, W, planes, slab; };
  const Phase ph1 = {_T1_, _W1_, M1, M2 * (M3 * p2)}, ph2 = {_T2_, _W2_, M3, M1 * (M4 * p1)};
  auto window_bad = [&](const Phase &ph) {
    const int T = v[ph.T], W = v[ph.W];
    return ceil_div(ph.planes, T) < W || (ph.planes == T && W > 0) || T * ph.slab > BUFFER_SIZE_LIMIT / (W + 1) / 2 / 2;
  };
  if (v[_T1_] < 1 || M1 < v[_T1_]) return fail(_T1_);
  if (window_bad(ph1)) return fail(_W1_);
  if (v[_Px1_] < 1 || v[_T1_] < v[_Px1_]) return fail(_Px1_);
  if (v[_Py1_] < 1 || M2 < v[_Py1_]) return fail(_Py1_);
  if (v[_Fz_] < 0 || v[_T1_] * M2 < v[_Fz_]) return fail(_Fz_);
  if (v[_FP1_] < 0 || v[_T1_] / v[_Px1_] * M2 / v[_Py1_] < v[_FP1_]) return fail(_FP1_);
  if (v[_Ux1_] < 1 || v[_T1_] < v[_Ux1_]) return fail(_Ux1_);
  if (v[_Uz1_] < 1 || M3 < v[_Uz1_]) return fail(_Uz1_);
  if (v[_Fy1_] < 0 || v[_T1_] * M3 < v[_Fy1_]) return fail(_Fy1_);
  if (v[_FU1_] < 0 || v[_T1_] / v[_Ux1_] * M3 / v[_Uz1_] < v[_FU1_]) return fail(_FU1_);
  if (v[_T2_] < 1 || M3 < v[_T2_]) return fail(_T2_);
  if (window_bad(ph2)) return fail(_W2_);
  if (v[_Fy2_] < 0 || v[_T2_] * M1 < v[_Fy2_]) return fail(_Fy2_);
  if (v[_Px2_] < 1 || M1 < v[_Px2_]) return fail(_Px2_);
  if (v[_Pz2_] < 1 || v[_T2_] < v[_Pz2_]) return fail(_Pz2_);
  if (v[_FP2_] < 0) return fail(_FP2_);
  if (v[_Uy2_] < 1 || M4 < v[_Uy2_]) return fail(_Uy2_);
  if (v[_Uz2_] < 1 || v[_T2_] < v[_Uz2_]) return fail(_Uz2_);
  if (v[_Fx_] < 0 || v[_T2_] * M4 < v[_Fx_]) return fail(_Fx_);
  if (v[_V_] < 0 || v[_V_] > 3) return fail(_V_);
  if (v[_S_] < 0 || v[_S_] > 1) return fail(_S_);
  return 0;
}

void params_adjust(int Nx, int Ny, int Nz, int p, int is_oned, int *v) {
  // a slab run only has one phase: pin the other one's knobs
  if (is_oned && v[_P1_] == 1) {
    v[_Ry_] = 10; v[_T2_] = 1; v[_W2_] = 0;
    for (int i : {_Fy2_, _FP2_, _FU2_, _Fx_}) v[i] = 0;
    for (int i : {_Pz2_, _Px2_, _Uz2_, _Uy2_}) v[i] = 1;
  }
  if (is_oned && v[_P1_] == p) {
    v[_Ry_] = 0; v[_T1_] = 1; v[_W1_] = 0;
    for (int i : {_Fz_, _FP1_, _FU1_, _Fy1_}) v[i] = 0;
    for (int i : {_Px1_, _Py1_, _Ux1_, _Uz1_}) v[i] = 1;
  }
  if (v[_W1_] == 0) for (int i : {_Fz_, _FP1_, _Fy1_, _FU1_}) v[i] = 0;
  if (v[_W2_] == 0) for (int i : {_Fy2_, _FP2_, _Fx_, _FU2_}) v[i] = 0;
  const int p1 = v[_P1_], p2 = p / p1;
  if (Ny % p2 == 0 && Nz % p2 == 0) v[_V_] &= 1;   // exact counts change nothing when the split is even
  if (Nx % p1 == 0 && Ny % p1 == 0) v[_V_] &= 2;
}

// The starting simplex of the reference's tuner (write_initial_simplex, offt-tuning.c:426-737), re-derived as a table of
// windows: every tunable of every vertex is drawn uniformly - one rand() per tunable, in index order, exactly as the
// reference consumes them, so the same srand() seed gives the same vertices - from the grid points inside a window
// [lo, hi] that depends on the tunables drawn before it:
//   P1        its whole grid (tuning_mode 1 / 2 pin it to 1 / p); vertices 0-1, 2-3, 4-5, 6-7 take the smallest, largest,
//             lower-middle and upper-middle grid point instead of the random one (:662-678)
//   T1, T2    between "at least 64 elements per message" and "at most BUFFER_SIZE_LIMIT/8 elements per buffer"
//   W1, W2    0..5, capped by the tile count and by the buffer budget; 0 for a single tile or is_W0
//   sub-tiles pairs (first, second) sized around the cache budget SUBTILE_SIZE: first in [sqrt(S/8/L), sqrt(16 S/L)]
//             stretched by the other extent, second so that first*second*L stays within [S/8, 16 S]
//   MPI_Test frequencies  between g/8 and 32 g (8 g for Fy2) calls, g = ranks of the phase's group; 0 when is_notest
//   Ry 0..min(10, M1) (pinned for slabs), V 0..3, S 0..1
void initial_simplex(int Nx, int Ny, int Nz, int p, int is_oned, int is_W0, int is_notest, int tuning_mode,
                     int **v_list, int *v_list_size, int (*x)[PARAM_COUNT]) {
  auto isqrt = [](int n) { return (int)std::sqrt((double)n); };
  const int S8 = SUBTILE_SIZE / 8, S16 = SUBTILE_SIZE * 16;
  for (int i = 0; i < PARAM_COUNT + 1; ++i) {
    int vv[PARAM_COUNT];
    int p1 = 1, p2 = 0, M1 = 0, M2 = 0, M3 = 0, M4 = 0;
    // first member of a sub-tile pair: 1..cap around sqrt of the budget over row length L, stretched by the other extent Mo
    auto pair_first = [&](int cap, int L, int Mo, int &lo, int &hi) {
      lo = std::max(1, std::min(std::min(cap, isqrt(S8 / L)), S8 / L / Mo));
      hi = std::min(cap, std::max(std::max(lo, isqrt(S16 / L)), S16 / L / Mo));
    };
    auto pair_second = [&](int cap, int L, int first, int &lo, int &hi) {
      lo = std::max(1, std::min(cap, S8 / L / first));
      hi = std::min(cap, std::max(lo, S16 / L / first));
    };
    auto freq = [&](int cap, int g, int mult, int &lo, int &hi) {
      lo = std::max(0, std::min(cap, g / 8));
      hi = std::min(cap, std::max(lo, g * mult));
      if (is_notest) lo = hi = 0;
    };
    auto window = [&](int planes, int T, int slab, int &lo, int &hi) {   // W of a phase with tile T, slab = elements per plane of the ring
      lo = hi = 0;
      if (is_W0 || planes == T) return;
      hi = std::min(5, std::max(0, std::min((planes + T - 1) / T, BUFFER_SIZE_LIMIT / 2 / 2 / (T * slab) - 1)));
    };
    for (int j = 0; j < PARAM_COUNT; ++j) {
      int lo = 0, hi = 0;
      switch (j) {
        case _P1_: lo = tuning_mode == 2 ? p : 1; hi = tuning_mode == 1 ? 1 : p; break;
        case _T1_:
          lo = std::max(1, std::min(M1, 64 * p2 / M2 / (M3 * p2)));
          hi = std::min(M1, std::max(lo, (BUFFER_SIZE_LIMIT / 8) / M2 / (M3 * p2)));
          break;
        case _W1_: window(M1, vv[_T1_], M2 * (M3 * p2), lo, hi); break;
        case _Px1_: pair_first(vv[_T1_], Nz, M2, lo, hi); break;
        case _Py1_: pair_second(M2, Nz, vv[_Px1_], lo, hi); break;
        case _Fz_: freq(vv[_T1_] * M2, p2, 32, lo, hi); break;
        case _FP1_: freq(vv[_T1_] / vv[_Px1_] * M2 / vv[_Py1_], p2, 32, lo, hi); break;
        case _Ux1_: pair_first(vv[_T1_], Ny, M3, lo, hi); break;
        case _Uz1_: pair_second(M3, Ny, vv[_Ux1_], lo, hi); break;
        case _FU1_: freq(vv[_T1_] / vv[_Ux1_] * M3 / vv[_Uz1_], p2, 32, lo, hi); break;
        case _Fy1_: freq(vv[_T1_] * M3, p2, 32, lo, hi); break;
        case _Ry_:
          lo = 0; hi = std::min(10, std::max(0, M1));
          if (is_oned && p1 == 1) lo = hi;
          else if (is_oned && p1 == p) lo = hi = 0;
          break;
        case _T2_:
          lo = std::max(1, std::min(M3, 64 * p1 / M1 / (M4 * p1)));
          hi = std::min(M3, std::max(lo, (BUFFER_SIZE_LIMIT / 8) / M1 / (M4 * p1)));
          break;
        case _W2_: window(M3, vv[_T2_], M1 * (M4 * p1), lo, hi); break;
        case _Fy2_: freq(vv[_T2_] * M1, p1, 8, lo, hi); break;
        case _Pz2_: pair_first(vv[_T2_], Ny, M1, lo, hi); break;
        case _Px2_: pair_second(M1, Ny, vv[_Pz2_], lo, hi); break;
        case _FP2_: freq(vv[_T2_] / vv[_Pz2_] * M1 / vv[_Px2_], p1, 32, lo, hi); break;
        case _Uz2_: pair_first(vv[_T2_], Nx, M4, lo, hi); break;
        case _Uy2_: pair_second(M4, Nx, vv[_Uz2_], lo, hi); break;
        case _FU2_: freq(vv[_T2_] / vv[_Uz2_] * M4 / vv[_Uy2_], p1, 32, lo, hi); break;
        case _Fx_: freq(vv[_T2_] * M4, p1, 32, lo, hi); break;
        case _V_: lo = 0; hi = 3; break;
        case _S_: lo = 0; hi = 1; break;
      }
      const int g_hi = grid_value_floor(1, v_list, v_list_size, j, hi);
      int g_lo = grid_value_ceil(1, v_list, v_list_size, j, lo);
      if (g_lo > g_hi) g_lo = g_hi;
      x[i][j] = (rand() % (g_hi - g_lo + 1)) + g_lo;
      if (j == _P1_) {
        if (i < 2) x[i][j] = g_lo;
        else if (i < 4) x[i][j] = g_hi;
        else if (i < 6) x[i][j] = (g_lo + g_hi) / 2;
        else if (i < 8) x[i][j] = (g_lo + g_hi + 1) / 2;
      }
      vv[j] = v_list[j][x[i][j]];
      if (j == _P1_) {
        p1 = vv[_P1_]; p2 = p / p1;
        M1 = (Nx + p1 - 1) / p1; M2 = (Ny + p2 - 1) / p2; M3 = (Nz + p2 - 1) / p2; M4 = (Ny + p1 - 1) / p1;
      }
    }
    // slabs: the phase that does not run has no y rows to balance (the other repairs are params_convert's, ADJUST_POINT)
    if (is_oned && p1 == 1) x[i][_Ry_] = 10;
    if (is_oned && p1 == p) x[i][_Ry_] = 0;
  }
}

const char *param_name(int i) { return kSpec[i].name; }

}  // namespace offtb

using namespace offtb;

extern "C" {

// z extent every tunable is sized by: Nz, or the Nz/2+1 complex points of a real-to-complex plan (offt-compute.c:3008, 3045, 3141)
static int nz_eff(const struct _offt_plan *po) { return po->is_r2c ? po->Nz / 2 + 1 : po->Nz; }

void params_range_setup(struct _offt_plan *po, int **v_list, int *v_list_size) {
  const auto g = params_grid(po->Nx, po->Ny, nz_eff(po), po->p);
  for (int i = 0; i < PARAM_COUNT; ++i) {
    v_list_size[i] = (int)g[i].size();
    v_list[i] = (int *)malloc(sizeof(int) * std::max<size_t>(g[i].size(), 1));   // caller frees, as in the reference
    std::copy(g[i].begin(), g[i].end(), v_list[i]);
  }
}

int grid_value_floor(int is_index, int **v_list, int *v_list_size, int i, int raw_v) {
  for (int j = v_list_size[i] - 1; j >= 0; --j)
    if (v_list[i][j] <= raw_v) return is_index ? j : v_list[i][j];
  return raw_v;
}

int grid_value_ceil(int is_index, int **v_list, int *v_list_size, int i, int raw_v) {
  for (int j = 0; j < v_list_size[i]; ++j)
    if (v_list[i][j] >= raw_v) return is_index ? j : v_list[i][j];
  return raw_v;
}

// offt-tuning.c:80-136: Active Harmony works in index space (one integer per tunable, an index into its value grid);
// backward = indices -> values followed by the ADJUST_POINT repairs, forward = values -> indices.  A value or index
// off the grid is fatal in the reference (printf + exit(-1)); so it is here.
void params_convert(int is_backward, int *v, long *ahv, struct _offt_plan *po, int **v_list, int *v_list_size) {
  if (is_backward) {
    for (int i = 0; i < PARAM_COUNT; ++i) {
      if (ahv[i] < 0 || ahv[i] >= v_list_size[i]) { printf("params_convert: bwd OUT OF RANGE ERROR\n"); exit(-1); }
      v[i] = v_list[i][ahv[i]];
    }
    params_adjust(po->Nx, po->Ny, nz_eff(po), po->p, po->is_oned, v);
  } else {
    for (int i = 0; i < PARAM_COUNT; ++i) {
      long found = -1;
      for (int c = 0; c < v_list_size[i] && found < 0; ++c)
        if (v_list[i][c] == v[i]) found = c;
      if (found < 0) { printf("params_convert: fwd OUT OF RANGE ERROR\n"); exit(-1); }
      ahv[i] = found;
    }
  }
}

// offt-tuning.c:144-226: 1 and the offending tunable in *p_i if the relations among the parameters do not hold
int is_infeasible_point(struct _offt_plan *po, int *v, int *p_i) { return params_infeasible(po->Nx, po->Ny, nz_eff(po), po->p, v, p_i); }

// offt-compute.c:3127-3225: heuristic defaults into po->params->v (computed for P1 = floor(sqrt(p)) on the grid,
// before any -d override, as in the reference)
void params_set_default(struct _offt_plan *po) {
  params_default(po->Nx, po->Ny, nz_eff(po), po->p, po->is_W0, po->is_notest, po->params->v);
}

// offt-tuning.c:426-737: the 25 starting vertices of the Nelder-Mead search, in grid-index space, written one per
// line to po->user_vertex_file (the patched nm.so reads them, strategies/nm.c:369-396)
void write_initial_simplex(struct _offt_plan *po, int **v_list, int *v_list_size) {
  int x[PARAM_COUNT + 1][PARAM_COUNT];
  initial_simplex(po->Nx, po->Ny, nz_eff(po), po->p, po->is_oned, po->is_W0, po->is_notest, po->tuning_mode, v_list, v_list_size, x);
  FILE *f = fopen(po->user_vertex_file, "w");
  if (!f) { printf("write_initial_simplex: cannot write %s\n", po->user_vertex_file); exit(-1); }
  for (int i = 0; i < PARAM_COUNT + 1; ++i) {
    for (int j = 0; j < PARAM_COUNT; ++j) fprintf(f, "%d ", x[i][j]);
    fprintf(f, "\n");
  }
  fclose(f);
}

void print_params(int *v) {
  // the reference prints in index order, skipping unset (negative) entries
  for (int i = 0; i < PARAM_COUNT; ++i)
    if (v[i] >= 0) printf("%s %d ", param_name(i), v[i]);
  printf("\n");
}

void offt_print_time(double *t) {
  // column order of the reference's timing line (offt-compute.c:3285-3293)
  static const int order[GES] = {ALL, INIT1, WAIT1, TEST1, INIT2, WAIT2, TEST2, TRANSPOSE,
                                 PACK1, UNPACK1, PACK2, UNPACK2, FFTz, FFTy1, FFTy2, FFTx};
  static const char *sep[GES] = {"", "  ", " ", " ", " ", " ", " ", "  ", "  ", " ", " ", " ", "  ", " ", " ", " "};
  for (int i = 0; i < GES; ++i) printf("%s%.5f", sep[i], t[order[i]]);
  printf("\n");
}

void offtb_params_default(int Nx, int Ny, int Nz, int p, int is_W0, int is_notest, int *v24) {
  params_default(Nx, Ny, Nz, p, is_W0, is_notest, v24);
}

int offtb_is_infeasible_point(int Nx, int Ny, int Nz, int p, const int *v24, int *bad_index) {
  int bad;
  int r = params_infeasible(Nx, Ny, Nz, p, v24, &bad);
  if (bad_index) *bad_index = bad;
  return r;
}

void offtb_params_adjust(int Nx, int Ny, int Nz, int p, int is_oned, int *v24) { params_adjust(Nx, Ny, Nz, p, is_oned, v24); }

void offtb_params_range(int Nx, int Ny, int Nz, int p, int *lists, int stride, int *sizes) {
  const auto g = params_grid(Nx, Ny, Nz, p);
  for (int i = 0; i < PARAM_COUNT; ++i) {
    sizes[i] = (int)std::min<size_t>(g[i].size(), (size_t)stride);
    std::copy(g[i].begin(), g[i].begin() + sizes[i], lists + (size_t)i * stride);
  }
}

}  // extern "C"
