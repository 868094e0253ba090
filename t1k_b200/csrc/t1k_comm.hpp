// NCCL plumbing of the read-sharded path: one communicator per process/GPU, created from a
// unique id that the launcher distributes (torch.distributed broadcast, a file, MPI ...).  libnccl is bound at run
// time with dlopen so that the library loads on boxes without NCCL and single-GPU use never touches it; inside a
// torch process the already loaded (torch-bundled) libnccl.so.2 is the one that resolves.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

#include <string>

namespace t1k {

// the part of nccl.h this path uses (values are ABI-stable across NCCL 2.x)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT8 = 0, NCCL_UINT8 = 1, NCCL_INT32 = 2, NCCL_UINT32 = 3, NCCL_INT64 = 4, NCCL_UINT64 = 5, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi {
  void *handle = nullptr;
  int (*GetUniqueId)(ncclUniqueId *) = nullptr;
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(int) = nullptr;
  std::string error;

  bool load() {
    if (handle) return true;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (int i = 0; i < 2 && !handle; ++i) handle = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    if (!handle) { error = std::string("cannot load libnccl: ") + dlerror(); return false; }
#define T1K_SYM(field, name)                                                              \
  do {                                                                                    \
    *(void **)(&field) = dlsym(handle, name);                                             \
    if (!field) { error = std::string("libnccl lacks ") + name; handle = nullptr; return false; } \
  } while (0)
    T1K_SYM(GetUniqueId, "ncclGetUniqueId");
    T1K_SYM(CommInitRank, "ncclCommInitRank");
    T1K_SYM(CommDestroy, "ncclCommDestroy");
    T1K_SYM(AllReduce, "ncclAllReduce");
    T1K_SYM(AllGather, "ncclAllGather");
    T1K_SYM(Broadcast, "ncclBroadcast");
    T1K_SYM(Send, "ncclSend");
    T1K_SYM(Recv, "ncclRecv");
    T1K_SYM(GroupStart, "ncclGroupStart");
    T1K_SYM(GroupEnd, "ncclGroupEnd");
    T1K_SYM(GetErrorString, "ncclGetErrorString");
#undef T1K_SYM
    return true;
  }
};

inline NcclApi &nccl() { static NcclApi api; return api; }

}  // namespace t1k

struct T1KComm {
  t1k::ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
};
