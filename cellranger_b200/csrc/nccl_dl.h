// nccl_dl.h — the handful of NCCL entry points the sharded run uses, resolved at run time (dlopen), so that
// libcrgpu.so carries no link-time dependency on NCCL: a single-GPU user never loads it, and inside a process
// that already holds an NCCL (torch bundles one) the same library instance is reused. Internal.
//
// Only types whose layout NCCL has kept stable since 2.0 are restated here (nccl.h: ncclUniqueId 128 bytes by
// value, ncclDataType_t / ncclRedOp_t enumerators, ncclComm_t opaque).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <stdlib.h>

#include <string>

namespace nccl_dl {

typedef struct ncclComm* comm_t;
struct unique_id {
  char internal[128];
};
enum { kSuccess = 0 };
enum { kUint8 = 1, kUint32 = 3, kUint64 = 5 };
enum { kSum = 0 };

struct Api {
  void* handle = nullptr;
  int (*GetUniqueId)(unique_id*) = nullptr;
  int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
  int (*CommDestroy)(comm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, comm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, comm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
  std::string error;
};

inline Api make_api() {
  Api api;
  const char* names[] = {getenv("CRGPU_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  for (const char* n : names) {
    if (!n || !*n) continue;
    api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (api.handle) break;
    const char* e = dlerror();
    api.error = e ? e : "dlopen failed";
  }
  if (!api.handle) return api;
  bool ok = true;
  auto sym = [&](const char* name) -> void* {
    void* p = dlsym(api.handle, name);
    if (!p) {
      ok = false;
      api.error = std::string("missing symbol ") + name;
    }
    return p;
  };
  api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(sym("ncclGetUniqueId"));
  api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(sym("ncclCommInitRank"));
  api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(sym("ncclCommDestroy"));
  api.AllReduce = reinterpret_cast<decltype(api.AllReduce)>(sym("ncclAllReduce"));
  api.AllGather = reinterpret_cast<decltype(api.AllGather)>(sym("ncclAllGather"));
  api.GroupStart = reinterpret_cast<decltype(api.GroupStart)>(sym("ncclGroupStart"));
  api.GroupEnd = reinterpret_cast<decltype(api.GroupEnd)>(sym("ncclGroupEnd"));
  api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(sym("ncclGetErrorString"));
  api.GetVersion = reinterpret_cast<decltype(api.GetVersion)>(sym("ncclGetVersion"));
  if (!ok) api.handle = nullptr;
  return api;
}

// the process-wide table (initialised once, thread safe); nullptr (with *why filled) when no NCCL can be loaded
inline const Api* load(std::string* why) {
  static const Api api = make_api();
  if (!api.handle) {
    if (why) *why = "NCCL is not available (" + api.error + "); set CRGPU_NCCL_LIB to the path of libnccl.so.2";
    return nullptr;
  }
  return &api;
}

}  // namespace nccl_dl
