// NCCL is bound at first use, not at load time: the host process may already carry an NCCL
// (PyTorch bundles its own libnccl.so.2, newer than the system one), and two different
// libnccl.so.2 cannot coexist in one process.  dlopen() by SONAME returns the copy that is
// already loaded, or loads the system library when there is none.
#pragma once
#include <dlfcn.h>
#include <nccl.h>   // types and prototypes only; nothing links against it
#include <string>
#include "common.cuh"

namespace scl {

struct NcclApi {
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclAllGather) AllGather = nullptr;
  decltype(&ncclReduce) Reduce = nullptr;
  decltype(&ncclBroadcast) Broadcast = nullptr;
  decltype(&ncclGroupStart) GroupStart = nullptr;
  decltype(&ncclGroupEnd) GroupEnd = nullptr;
};

inline const NcclApi& nccl_api() {
  static const NcclApi api = [] {
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) throw Error(-5 /*SCL_ERR_NCCL*/, std::string("libnccl.so.2 cannot be loaded: ") + dlerror());
    NcclApi a;
#define SCL_NCCL_SYM(field, name)                                                            \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(lib, name));                           \
  if (!a.field) throw Error(-5, std::string("libnccl.so.2 lacks ") + name)
    SCL_NCCL_SYM(GetUniqueId, "ncclGetUniqueId");
    SCL_NCCL_SYM(CommInitRank, "ncclCommInitRank");
    SCL_NCCL_SYM(CommDestroy, "ncclCommDestroy");
    SCL_NCCL_SYM(GetErrorString, "ncclGetErrorString");
    SCL_NCCL_SYM(AllReduce, "ncclAllReduce");
    SCL_NCCL_SYM(AllGather, "ncclAllGather");
    SCL_NCCL_SYM(Reduce, "ncclReduce");
    SCL_NCCL_SYM(Broadcast, "ncclBroadcast");
    SCL_NCCL_SYM(GroupStart, "ncclGroupStart");
    SCL_NCCL_SYM(GroupEnd, "ncclGroupEnd");
#undef SCL_NCCL_SYM
    return a;
  }();
  return api;
}

}  // namespace scl
