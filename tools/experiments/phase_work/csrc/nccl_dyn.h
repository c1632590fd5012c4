// NCCL resolved at run time instead of at link time.
//
// The library is loaded into processes that may already carry a different NCCL (PyTorch bundles
// its own libnccl.so.2); linking a second copy in clashes on symbols.  The first multi-rank
// world_init binds to the libnccl.so.2 already in the process if there is one, else dlopens the
// system's.  Single-rank and emulated-rank worlds never touch NCCL.
#pragma once
#include <nccl.h>

namespace offtb {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char *(*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int *);
};

// nullptr (with the error set) if no NCCL can be found
const NcclApi *nccl_api();

}  // namespace offtb
