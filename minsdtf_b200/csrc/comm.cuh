// comm.cuh — NCCL plumbing for the 2-way CFG split: the cond / uncond UNet evaluations of one
// step (stable_diffusion.py:454-457) run on two GPUs and exchange their epsilon with one ncclAllGather per step,
// enqueued on the compute stream so it is captured inside the step's CUDA graph.
// libnccl is dlopen'ed (the torch-bundled libnccl.so.2 or the system one): libsdtf.so has no link-time dependency.
#pragma once
#include <dlfcn.h>

#include "common.cuh"

namespace sdtf {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* NcclComm;

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, NcclComm, cudaStream_t) = nullptr;
  int (*CommDestroy)(NcclComm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};

inline NcclApi& nccl_api(const char* path) {
  static NcclApi api;
  if (api.lib) return api;
  const char* cands[] = {path, getenv("SDTF_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* c : cands) {
    if (!c || !c[0]) continue;
    api.lib = dlopen(c, RTLD_NOW | RTLD_GLOBAL);
    if (api.lib) break;
  }
  if (!api.lib) throw Error(std::string("cannot load libnccl (tried the given path, $SDTF_NCCL_LIB, libnccl.so.2): ") + dlerror());
  auto sym = [&](const char* n) {
    void* s = dlsym(api.lib, n);
    if (!s) throw Error(std::string("libnccl is missing symbol ") + n);
    return s;
  };
  api.GetUniqueId = reinterpret_cast<int (*)(NcclUniqueId*)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<int (*)(NcclComm*, int, NcclUniqueId, int)>(sym("ncclCommInitRank"));
  api.AllGather = reinterpret_cast<int (*)(const void*, void*, size_t, int, NcclComm, cudaStream_t)>(sym("ncclAllGather"));
  api.CommDestroy = reinterpret_cast<int (*)(NcclComm)>(sym("ncclCommDestroy"));
  api.GetErrorString = reinterpret_cast<const char* (*)(int)>(sym("ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<int (*)(int*)>(sym("ncclGetVersion"));
  return api;
}

#define SDTF_NCCL(api, expr)                                                                                   \
  do {                                                                                                         \
    int _r = (expr);                                                                                           \
    if (_r != 0) throw ::sdtf::Error(std::string(#expr) + " failed: NCCL " + (api).GetErrorString(_r));       \
  } while (0)

static constexpr int kNcclFloat32 = 7;  // ncclFloat32 in nccl.h

struct Comm {
  NcclComm comm = nullptr;
  int rank = 0, world = 1;
  bool active() const { return comm != nullptr && world > 1; }
};

}  // namespace sdtf
