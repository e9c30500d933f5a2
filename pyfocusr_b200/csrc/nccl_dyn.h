// NCCL, resolved at run time.  Only the row-partitioned multi-GPU solve (focusr_eigs_smallest_dist)
// talks to other GPUs -- one halo exchange per SpMM (grouped ncclSend/ncclRecv of boundary rows over
// NVLink) and one small all-reduce per Gram / residual / norm -- so the library itself has no link-time
// dependency on NCCL: the symbols are looked up in the libnccl.so.2 that torch has already loaded
// (NCCL 2.28.9 here), or the system one.
#pragma once
#include <dlfcn.h>
#include <stddef.h>

#include <cuda_runtime.h>

namespace fb {

typedef struct ncclComm* ncclComm_t;
typedef struct {
  char internal[128];
} ncclUniqueId;
// values of nccl.h (stable across NCCL 2.x)
enum { NCCL_SUCCESS = 0 };
enum { NCCL_INT64 = 4, NCCL_FLOAT64 = 8 };
enum { NCCL_SUM = 0, NCCL_MAX = 2, NCCL_MIN = 3 };

struct NcclApi {
  int (*GetUniqueId)(ncclUniqueId*);
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t);
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)();
  int (*GroupEnd)();
  const char* (*GetErrorString)(int);
  bool ok = false;
};

inline NcclApi& nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (tried) return api;
  tried = true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) return api;
#define FB_NCCL_SYM(field, name)                                  \
  *(void**)(&api.field) = dlsym(h, name);                         \
  if (!api.field) return api;
  FB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
  FB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
  FB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
  FB_NCCL_SYM(AllReduce, "ncclAllReduce")
  FB_NCCL_SYM(AllGather, "ncclAllGather")
  FB_NCCL_SYM(Send, "ncclSend")
  FB_NCCL_SYM(Recv, "ncclRecv")
  FB_NCCL_SYM(GroupStart, "ncclGroupStart")
  FB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
  FB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef FB_NCCL_SYM
  api.ok = true;
  return api;
}

}  // namespace fb
