// NCCL bound at run time (dlopen), so that the library loads -- and the single-GPU path runs -- where no NCCL is
// installed, and so that inside a PyTorch process the SAME libnccl.so.2 instance torch already loaded is used.
// Only the few entry points of the cell-sharded exchange are bound; the declarations restate the stable NCCL 2.x ABI
// (nccl.h: ncclUniqueId is 128 opaque bytes passed by value; ncclFloat32 = 7, ncclBfloat16 = 9, ncclSum = 0).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>

namespace tgb {

struct NcclUniqueId { char internal[128]; };
constexpr int kNcclFloat32 = 7, kNcclBfloat16 = 9, kNcclSum = 0;

struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*ReduceScatter)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  // optional (NCCL >= 2.19): memory NCCL can map into the NVSwitch multicast space + user-buffer registration, so that an
  // in-place all-reduce of that buffer runs in the switch (NVLS) without staging copies
  int (*MemAlloc)(void**, size_t) = nullptr;
  int (*MemFree)(void*) = nullptr;
  int (*CommRegister)(void*, void*, size_t, void**) = nullptr;
  int (*CommDeregister)(void*, void*) = nullptr;
  bool ok() const { return lib != nullptr; }
};

// Process-wide, resolved on first use.  Returns nullptr (and a message) when no libnccl.so.2 can be found.
static inline NcclApi* nccl_api(char* err, size_t n) {
  static NcclApi api;
  static bool tried = false;
  if (api.ok()) return &api;
  if (tried) { snprintf(err, n, "libnccl.so.2 is not available in this process"); return nullptr; }
  tried = true;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);      // the instance torch (or the host) already loaded
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) { snprintf(err, n, "dlopen(libnccl.so.2): %s", dlerror()); return nullptr; }
  NcclApi a;
  a.lib = lib;
  bool good = true;
  auto sym = [&](const char* name) { void* p = dlsym(lib, name); if (!p) good = false; return p; };
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(sym("ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(sym("ncclCommInitRank"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(sym("ncclCommDestroy"));
  a.AllReduce = reinterpret_cast<decltype(a.AllReduce)>(sym("ncclAllReduce"));
  a.ReduceScatter = reinterpret_cast<decltype(a.ReduceScatter)>(sym("ncclReduceScatter"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(sym("ncclAllGather"));
  a.GroupStart = reinterpret_cast<decltype(a.GroupStart)>(sym("ncclGroupStart"));
  a.GroupEnd = reinterpret_cast<decltype(a.GroupEnd)>(sym("ncclGroupEnd"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(sym("ncclGetErrorString"));
  if (!good) { snprintf(err, n, "libnccl.so.2 lacks an expected entry point"); return nullptr; }
  a.MemAlloc = reinterpret_cast<decltype(a.MemAlloc)>(dlsym(lib, "ncclMemAlloc"));
  a.MemFree = reinterpret_cast<decltype(a.MemFree)>(dlsym(lib, "ncclMemFree"));
  a.CommRegister = reinterpret_cast<decltype(a.CommRegister)>(dlsym(lib, "ncclCommRegister"));
  a.CommDeregister = reinterpret_cast<decltype(a.CommDeregister)>(dlsym(lib, "ncclCommDeregister"));
  api = a;
  return &api;
}

}  // namespace tgb
