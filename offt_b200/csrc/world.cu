// Process-global "world": which rank this process is, on which GPU, and the NCCL
// communicator that carries the exchanges.  Replaces the reference's implicit
// MPI_COMM_WORLD (offt-compute.c:3315-3316) and its comm1/comm2 sub-communicators
// (offt-compute.c:78-125): the row and column groups are addressed as explicit peer lists
// on the one world communicator.
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "engine.h"

namespace offtb {

static char g_error[1024] = "";
int g_exit_on_error = 1;

void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
const char *last_error() { return g_error; }

void fatal_or_return(const char *where) {
  fprintf(stderr, "offt_b200: %s: %s\n", where, g_error);
  if (g_exit_on_error) exit(-1);
}

World &world() {
  static World w;
  return w;
}

// 0 if `ok` holds on every rank.  Collective; a rank-local failure inside it still completes the all-reduce where
// it can, so the other ranks are not left blocked.
int world_agree_ok(bool ok) {
  World &w = world();
  if (!w.nccl) return ok ? 0 : -1;
  const NcclApi *nc = nccl_api();
  static int *d_flag = nullptr;
  bool local_ok = ok && nc != nullptr;
  if (!d_flag && cudaMalloc(&d_flag, sizeof(int)) != cudaSuccess) { cudaGetLastError(); d_flag = nullptr; }
  if (!d_flag || !nc) return -1;   // cannot even take part: nothing sane is left to do collectively
  const int bad = local_ok ? 0 : 1;
  int total = 1;
  if (cudaMemcpy(d_flag, &bad, sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess) cudaGetLastError();
  const ncclResult_t r = nc->AllReduce(d_flag, d_flag, 1, ncclInt, ncclSum, w.nccl, 0);
  if (r == ncclSuccess && cudaStreamSynchronize(0) == cudaSuccess && cudaMemcpy(&total, d_flag, sizeof(int), cudaMemcpyDeviceToHost) == cudaSuccess)
    return total == 0 ? 0 : -1;
  cudaGetLastError();
  return -1;
}

// Collective: no early return between its first and last collective call - a rank whose own steps fail (no
// allocation to offer, cudaIpcGetMemHandle, a failed import) offers a zero handle / records the failure, still runs
// the all-gather and the agreement, and every rank returns the same status.
int world_ipc_share(void *mine, std::vector<void *> &mapped) {
  World &w = world();
  mapped.assign(w.size, nullptr);
  if (!w.nccl) { set_error("world_ipc_share needs an NCCL world"); return -1; }
  const NcclApi *nc = nccl_api();
  int rc = nc ? 0 : -1;
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  cudaIpcMemHandle_t h;
  memset(&h, 0, sizeof(h));
  unsigned char have = 0;
  if (!rc && mine) {
    cudaError_t e = cudaIpcGetMemHandle(&h, mine);
    if (e == cudaSuccess) have = 1;
    else { cudaGetLastError(); set_error("cudaIpcGetMemHandle: %s", cudaGetErrorString(e)); rc = -1; }
  } else if (!mine) {
    set_error("nothing to share (allocation failed on this rank)");
    rc = -1;
  }
  // record = handle + one byte "valid"
  const size_t rb = hb + 8;
  std::vector<unsigned char> rec(rb, 0), all(rb * (size_t)w.size, 0);
  memcpy(rec.data(), &h, hb);
  rec[hb] = have;
  unsigned char *d_all = nullptr;
  bool gathered = false;
  if (nc && cudaMalloc(&d_all, rb * (size_t)w.size) == cudaSuccess) {
    cudaMemcpy(d_all + rb * (size_t)w.rank, rec.data(), rb, cudaMemcpyHostToDevice);
    if (nc->AllGather(d_all + rb * (size_t)w.rank, d_all, rb, ncclUint8, w.nccl, 0) == ncclSuccess &&
        cudaStreamSynchronize(0) == cudaSuccess &&
        cudaMemcpy(all.data(), d_all, rb * (size_t)w.size, cudaMemcpyDeviceToHost) == cudaSuccess)
      gathered = true;
  }
  if (!gathered) { cudaGetLastError(); if (!rc) set_error("all-gather of the IPC handles failed"); rc = -1; }
  if (d_all) cudaFree(d_all);
  if (gathered) {
    for (int r = 0; r < w.size; ++r) {
      if (r == w.rank) { mapped[r] = mine; continue; }
      if (!all[rb * (size_t)r + hb]) { if (!rc) set_error("rank %d has nothing to share", r); rc = -1; continue; }
      cudaIpcMemHandle_t hr;
      memcpy(&hr, &all[rb * (size_t)r], hb);
      cudaError_t e = cudaIpcOpenMemHandle(&mapped[r], hr, cudaIpcMemLazyEnablePeerAccess);
      if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
        mapped[r] = nullptr;
        rc = -1;
      }
    }
  }
  // all ranks must agree whether the mapping worked
  if (world_agree_ok(rc == 0) != 0) {
    if (!rc) set_error("peer mapping failed on another rank");
    world_ipc_release(mapped);
    return -1;
  }
  return 0;
}

void world_ipc_release(std::vector<void *> &mapped) {
  World &w = world();
  for (int r = 0; r < (int)mapped.size(); ++r)
    if (mapped[r] && r != w.rank) cudaIpcCloseMemHandle(mapped[r]);
  mapped.clear();
}

}  // namespace offtb

using namespace offtb;

extern "C" {

const char *offtb_last_error(void) { return last_error(); }

int offtb_set_exit_on_error(int on) { g_exit_on_error = on; return 0; }

void offtb_clear_error(void) { g_error[0] = 0; }

int offtb_get_unique_id(void *id128) {
  static_assert(sizeof(ncclUniqueId) <= OFFTB_UNIQUE_ID_BYTES, "unique id does not fit");
  ncclUniqueId id;
  const NcclApi *nc = nccl_api();
  if (!nc) return -1;
  OFFTB_NCCL(nc->GetUniqueId(&id));
  memset(id128, 0, OFFTB_UNIQUE_ID_BYTES);
  memcpy(id128, &id, sizeof(id));
  return 0;
}

int offtb_world_init(int rank, int size, int device, const void *id128) {
  World &w = world();
  if (w.up) { set_error("world already initialised"); return -1; }
  if (size < 1 || rank < 0 || rank >= size) { set_error("bad rank/size %d/%d", rank, size); return -1; }
  int ndev = 0;
  OFFTB_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < 1) { set_error("no CUDA device: this library has no CPU path"); return -1; }
  if (device < 0) device = rank % ndev;
  OFFTB_CUDA(cudaSetDevice(device));
  OFFTB_CUDA(cudaFree(0));
  if (size > 1) {
    if (!id128) { set_error("a unique id is required for size > 1"); return -1; }
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    const NcclApi *nc = nccl_api();
    if (!nc) return -1;
    OFFTB_NCCL(nc->CommInitRank(&w.nccl, size, id, rank));
  }
  w.up = true; w.local = false; w.size = size; w.rank = rank; w.device = device;
  return 0;
}

int offtb_world_init_local(int size, int device) {
  World &w = world();
  if (w.up) { set_error("world already initialised"); return -1; }
  if (size < 1) { set_error("bad size %d", size); return -1; }
  int ndev = 0;
  OFFTB_CUDA(cudaGetDeviceCount(&ndev));
  if (ndev < 1) { set_error("no CUDA device: this library has no CPU path"); return -1; }
  if (device < 0) device = 0;
  OFFTB_CUDA(cudaSetDevice(device));
  OFFTB_CUDA(cudaFree(0));
  w.up = true; w.local = true; w.size = size; w.rank = 0; w.device = device;
  return 0;
}

int offtb_world_set_rank(int rank) {
  World &w = world();
  if (!w.up || !w.local) { set_error("offtb_world_set_rank needs a local world"); return -1; }
  if (rank < 0 || rank >= w.size) { set_error("bad rank %d", rank); return -1; }
  w.rank = rank;
  return 0;
}

int offtb_world_size(void) { return world().up ? world().size : 0; }
int offtb_world_rank(void) { return world().up ? world().rank : -1; }

int offtb_world_barrier(void) {
  World &w = world();
  if (!w.up) { set_error("world not initialised"); return -1; }
  OFFTB_CUDA(cudaDeviceSynchronize());
  if (w.nccl) {
    static int *flag = nullptr;
    if (!flag) OFFTB_CUDA(cudaMalloc(&flag, sizeof(int)));
    OFFTB_NCCL(nccl_api()->AllReduce(flag, flag, 1, ncclInt, ncclSum, w.nccl, 0));
    OFFTB_CUDA(cudaStreamSynchronize(0));
  }
  return 0;
}

void offtb_release_finished_plans(void);

void offtb_world_fin(void) {
  World &w = world();
  offtb_release_finished_plans();
  if (!w.up) return;
  if (w.nccl) { nccl_api()->CommDestroy(w.nccl); w.nccl = nullptr; }
  w = World();
}

}  // extern "C"
