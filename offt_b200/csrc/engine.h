// Internal declarations shared by the host-side translation units.
#pragma once
#ifndef OFFT_NO_MINMAX
#define OFFT_NO_MINMAX
#endif
#include <cuda_runtime.h>
#include "nccl_dyn.h"

#include <string>
#include <vector>

#include "fft_launch.h"
#include "offt.h"
#include "offt_b200.h"

namespace offtb {

// ---- errors ---------------------------------------------------------------------------
void set_error(const char *fmt, ...);
const char *last_error();
// what the reference does on a fatal condition (printf + exit(-1), offt-compute.c:3440-3443);
// bindings that want a return code instead call offtb_set_exit_on_error(0)
extern int g_exit_on_error;
extern int g_force_generic;   // > 0: every launch runs on the generic kernel (tests; OFFTB_GENERIC=1)
void fatal_or_return(const char *where);

#define OFFTB_CUDA(call)                                                                   \
  do {                                                                                     \
    cudaError_t e__ = (call);                                                              \
    if (e__ != cudaSuccess) {                                                              \
      offtb::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)
#define OFFTB_NCCL(call)                                                                   \
  do {                                                                                     \
    ncclResult_t r__ = (call);                                                             \
    if (r__ != ncclSuccess) {                                                              \
      offtb::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, offtb::nccl_api()->GetErrorString(r__)); \
      return -1;                                                                           \
    }                                                                                      \
  } while (0)

// ---- world ----------------------------------------------------------------------------
struct World {
  bool up = false;
  bool local = false;   // all ranks emulated in this process on one device
  int size = 1, rank = 0, device = 0;
  ncclComm_t nccl = nullptr;
};
World &world();
// Collective over the NCCL world: every rank passes one cudaMalloc'd allocation and gets back the addresses at
// which all ranks' allocations are mapped in this process (its own pointer for itself), over cudaIpc handles
// gathered with ncclAllGather.  Returns non-zero (error set) if peer mapping is not possible.
int world_ipc_share(void *mine, std::vector<void *> &mapped);
void world_ipc_release(std::vector<void *> &mapped);
// collective: 0 if `ok` is true on every rank of the world
int world_agree_ok(bool ok);

// ---- tunables (params.cu) ---------------------------------------------------------------
std::vector<std::vector<int>> params_grid(int Nx, int Ny, int Nz, int p);
void params_default(int Nx, int Ny, int Nz, int p, int is_W0, int is_notest, int *v);
int params_infeasible(int Nx, int Ny, int Nz, int p, const int *v, int *bad);
void params_adjust(int Nx, int Ny, int Nz, int p, int is_oned, int *v);
const char *param_name(int i);
void initial_simplex(int Nx, int Ny, int Nz, int p, int is_oned, int is_W0, int is_notest, int tuning_mode,
                     int **v_list, int *v_list_size, int (*x)[PARAM_COUNT]);

// ---- layout (plan.cu) ---------------------------------------------------------------------
void comm_fill(struct _offt_comm *c, int Nx, int Ny, int Nz, int p, int p1, int rank, int S, int is_equalxy, int is_r2c = 0);
long long alloc_elems(int Nx, int Ny, int Nz, int p, int p1, int is_r2c = 0);
int check_supported(int Nx, int Ny, int Nz, int p, int p1, int is_r2c = 0);

enum Schedule { SCHED_SINGLE, SCHED_SLAB_1XP, SCHED_SLAB_PX1, SCHED_PENCIL };
enum StageId { ST_K1, ST_K2, ST_K3, ST_K4, ST_X1, ST_X2, ST_H2D, ST_D2H, ST_COUNT };

struct Ring {
  int depth = 0;                 // W + 1
  long long slot_elems = 0;      // complex elements per send (or recv) slot
  std::vector<void *> send, recv;
  std::vector<long long> send_off, recv_off;   // the same slots as element offsets into the ring chunk
  std::vector<cudaEvent_t> packed, recvd;
  unsigned long long tiles_done = 0;           // tiles of this phase executed so far (sequence base of the flags)
};

// Flag words of the fused exchange, one block per rank in an IPC-shared allocation.  Peers write them over NVLink.
//   arrived[phase][slot][j]  : group member j has stored its block of tile number `arrived` into my landing slot `slot`
//   released[phase][slot][j] : group member j has finished reading tile number `released` out of its slot `slot`
#define OFFTB_DONE_SLOTS 32
#define OFFTB_MAX_SLOTS 65   // ring depth W + 1 with W <= 64
// One word per (phase, ring slot, group member): the launches of a dependent-launch chain run concurrently and may
// finish out of order, so tile s and tile s+1 must not announce themselves through the same word (a later tile's
// number overwritten by an earlier tile's would make the word go backwards, and ">= s" would no longer mean "s is
// complete").  A slot's tenants s, s+depth, ... are ordered by the protocol itself, so each word only ever grows.
struct XFlags {
  unsigned arrived[2][OFFTB_MAX_SLOTS][OFFTB_MAX_GROUP];
  unsigned released[2][OFFTB_MAX_SLOTS][OFFTB_MAX_GROUP];
  // [phase][writer, reader][tile number mod OFFTB_DONE_SLOTS]: launches of a dependent-launch chain overlap, so
  // consecutive tiles count their finished CTAs in different words
  unsigned done_counter[2][2][OFFTB_DONE_SLOTS];
};

enum ExchangeMode { XCHG_NCCL, XCHG_FUSED };

struct Engine {
  struct _offt_plan *po = nullptr;
  int prec = PREC_F64;
  size_t esz = 16;               // bytes per complex element
  Schedule sched = SCHED_SINGLE;
  int rank_x = 0, rank_y = 0;
  long long alloc = 0;           // complex elements of the caller's array
  void *d_user = nullptr;        // device copy when the caller passes host memory
  void *d_scratch = nullptr;     // second array for the transposed output layouts
  void *d_ring = nullptr;        // one chunk carved into both phases' rings (disjoint parts, see engine_create)
  ExchangeMode xmode = XCHG_NCCL;
  int grid_cap[2] = {0, 0};                // CTA budgets of writer and reader launches while they overlap
  FftShape *dry_shape = nullptr;           // run_launch only reports the launch shape
  XFlags *d_flags = nullptr;               // this rank's flag block
  unsigned *h_error = nullptr;             // pinned host word the kernels write when a flag wait times out ...
  unsigned *d_error = nullptr;             // ... and its device alias
  unsigned long long wait_timeout_ns = 0;  // OFFTB_FLAG_TIMEOUT_S (default 300 s; 0: wait for ever)
  bool failed = false;                     // an exchange timed out: the plan's flags are no longer consistent
  bool narrow_now = false;                 // the phase being enqueued uses half-width strided tiles (run_phase)
  int reader_depth = 0, depth_next = 0;    // ring depth of the reader launches while they share the SMs with the writers (plan_overlap)
  int pdl_next = 0;                        // dependent-launch bits of the next launch (run_phase sets, produce/consume read)
  bool chain_timing = false;               // stage timing of a dependent-launch chain: one event pair per chain
  std::vector<void *> peer_ring;           // every world rank's ring chunk as mapped here (fused mode)
  std::vector<void *> peer_flags;          // every world rank's flag block as mapped here
  void *tw[3] = {nullptr, nullptr, nullptr};  // twiddles for Nx, Ny, Nz (compact per-stage tables of the power-of-two kernels)
  void *tw_half = nullptr, *tw_r2c = nullptr;      // real-to-complex plans: tables of the Nz/2-point transform and exp(-2*pi*i*k/Nz), k <= Nz/4
  bool r2c_fast = false;                           // the local z pass runs as Nz/2 complex points + r2c_pass.cu
  void *tw_full[3] = {nullptr, nullptr, nullptr};  // exp(-2*pi*i*k/N), k < N, for the generic kernel
  Ring ring[2];
  cudaStream_t s_comp = nullptr, s_comm = nullptr, s_user = nullptr;
  bool async = false;
  bool stage_timing = true;      // CUDA events around every launch (or chain of launches) feed the reference's po->t[] buckets
  double post_s[2] = {0.0, 0.0}; // host seconds spent enqueueing phase 1 / phase 2 (the INIT1 / INIT2 "posting cost" buckets)
  void *registered_host = nullptr;
  size_t registered_bytes = 0;
  int launches = 0;
  double last_ms = 0.0;
  double stage_ms[ST_COUNT] = {0};
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> timed;  // (stage, begin/end)
  std::vector<cudaEvent_t> event_pool;
  size_t event_next = 0;
};

int engine_create(struct _offt_plan *po);
void engine_destroy(struct _offt_plan *po);
// runs the schedule for every plan of `group` (one plan per process in NCCL worlds, all ranks
// in local worlds); arrays are the callers' in-place arrays (host or device)
int engine_execute(std::vector<struct _offt_plan *> &group, std::vector<double *> &arrays, bool inverse);

}  // namespace offtb
