// NCCL bound at run time (dlopen): libb200exec has no link-time dependency on it, an executor that runs a
// single GPU never loads it, and a process that already carries a libnccl (e.g. a Python harness that imported
// torch) shares that copy instead of loading a second one.
//
// Reference counterpart: the Arrow Flight client/server pair that moves shuffle partitions between executors
// (ballista/core/src/client.rs:143-220, ballista/executor/src/flight_service.rs:88-306).  On one box with one
// executor per GPU the same bytes cross NVLink/NVSwitch with grouped ncclSend/ncclRecv (an all-to-all-v).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

#include <mutex>
#include <string>

namespace b200 {

// the few NCCL types needed (ABI-stable since NCCL 2.0; see nccl.h)
typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
typedef int ncclResult_t;   // 0 == ncclSuccess
typedef int ncclDataType_t; // ncclInt8/ncclChar = 0, ncclUint8 = 1
static const int kNcclUint8 = 1;

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
  std::string error;

  static NcclApi& get() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] { api.load(); });
    return api;
  }
  bool ok() const { return handle != nullptr; }

 private:
  template <class F>
  bool sym(F& f, const char* name) {
    f = (F)dlsym(handle, name);
    if (!f) error = std::string("libnccl lacks ") + name;
    return f != nullptr;
  }
  void load() {
    const char* override_path = getenv("B200_NCCL_LIB");
    const char* names[] = {override_path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
      error = dlerror();
    }
    if (!handle) return;
    if (!(sym(GetUniqueId, "ncclGetUniqueId") && sym(CommInitRank, "ncclCommInitRank") && sym(CommDestroy, "ncclCommDestroy") && sym(Send, "ncclSend") &&
          sym(Recv, "ncclRecv") && sym(GroupStart, "ncclGroupStart") && sym(GroupEnd, "ncclGroupEnd") && sym(GetErrorString, "ncclGetErrorString") &&
          sym(GetVersion, "ncclGetVersion"))) {
      handle = nullptr;
    }
  }
};

}  // namespace b200
