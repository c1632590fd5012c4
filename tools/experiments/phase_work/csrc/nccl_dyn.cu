#include <dlfcn.h>

#include "engine.h"
#include "nccl_dyn.h"

namespace offtb {

const NcclApi *nccl_api() {
  static NcclApi api;
  static int state = 0;   // 0 untried, 1 ok, -1 failed
  if (state == 1) return &api;
  if (state == -1) { set_error("NCCL is not available in this process (libnccl.so.2 not found)"); return nullptr; }
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);   // the copy torch (or the host program) already loaded
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  if (!h) { state = -1; set_error("dlopen libnccl.so.2: %s", dlerror()); return nullptr; }
  bool ok = true;
  auto sym = [&](const char *name) { void *p = dlsym(h, name); if (!p) ok = false; return p; };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.Send = (decltype(api.Send))sym("ncclSend");
  api.Recv = (decltype(api.Recv))sym("ncclRecv");
  api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  if (!ok) { state = -1; set_error("libnccl.so.2 lacks a required symbol"); return nullptr; }
  state = 1;
  return &api;
}

}  // namespace offtb
