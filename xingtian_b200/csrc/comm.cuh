// comm.cuh -- data-parallel gradient exchange: one ncclAllReduce(sum) over the flat fp32 gradient bucket
// per SGD step, issued by the library itself on the library's stream so that it is captured in the training CUDA graph.
// Reference precedent: tower-gradient averaging zeus/trainer/trainer_tf.py:187-203, xt/framework/trainer.py:82-92.
//
// NCCL is the copy PyTorch ships (nvidia/nccl/lib/libnccl.so.2), resolved at run time with dlopen: the library has
// no link-time dependency on it and single-GPU hosts never load it.  The communicator is the library's own (the host
// broadcasts the unique id over whatever channel it has -- torch.distributed in engine.py).
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stddef.h>
#include <string.h>

namespace xtb {

struct NcclUniqueId { char internal[128]; };
typedef struct ncclComm* ncclComm_t;

struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, NcclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  const char* load(const char* path) {
    if (handle) return nullptr;
    const char* cands[3] = {path, "libnccl.so.2", "libnccl.so"};
    for (int i = 0; i < 3 && !handle; i++)
      if (cands[i] && cands[i][0]) handle = dlopen(cands[i], RTLD_NOW | RTLD_GLOBAL);
    if (!handle) return dlerror();
#define XTB_NCCL_SYM(field, name) \
  *(void**)(&field) = dlsym(handle, name); \
  if (!field) return "missing NCCL symbol " name;
    XTB_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    XTB_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    XTB_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    XTB_NCCL_SYM(AllReduce, "ncclAllReduce")
    XTB_NCCL_SYM(GroupStart, "ncclGroupStart")
    XTB_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    XTB_NCCL_SYM(GetErrorString, "ncclGetErrorString")
#undef XTB_NCCL_SYM
    return nullptr;
  }
};

constexpr int kNcclFloat = 7;   // ncclFloat32
constexpr int kNcclSum = 0;     // ncclSum

}  // namespace xtb

struct xtb_comm {
  xtb::ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  cudaStream_t side = nullptr;            // the early bucket is reduced here while the backward pass continues
  cudaEvent_t fork = nullptr, join = nullptr;
};
