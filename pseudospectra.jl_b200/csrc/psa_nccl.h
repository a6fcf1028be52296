// NCCL transport for multi-device contexts (one host process driving several GPUs): broadcast of the matrix T
// to every device and gather of the sigma_min tiles on device 0, both over NVLink/NVSwitch.
// libnccl.so.2 is resolved with dlopen on first use, so single-device users (and one-process-per-GPU hosts that
// shard rows themselves, e.g. bench.py under torchrun) carry no NCCL dependency.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <string>
#include <vector>

namespace psa {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;

  bool load(std::string& err) {
    if (handle) return true;
    handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) handle = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!handle) {
      err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror();
      return false;
    }
#define PSA_NCCL_SYM(field, name)                                  \
  field = reinterpret_cast<decltype(field)>(dlsym(handle, name)); \
  if (!field) {                                                    \
    err = std::string("missing NCCL symbol ") + name;              \
    return false;                                                  \
  }
    PSA_NCCL_SYM(CommInitAll, "ncclCommInitAll")
    PSA_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PSA_NCCL_SYM(GroupStart, "ncclGroupStart")
    PSA_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PSA_NCCL_SYM(Broadcast, "ncclBroadcast")
    PSA_NCCL_SYM(Send, "ncclSend")
    PSA_NCCL_SYM(Recv, "ncclRecv")
    PSA_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef PSA_NCCL_SYM
    return true;
  }
};

}  // namespace psa
