// NCCL bound at run time (dlopen) so that libdjets_b200.so loads on hosts without NCCL, and so
// that a process which already holds an NCCL (e.g. the one bundled with PyTorch) shares it.
// Only the point-to-point and all-reduce entry points the block-operator exchange needs.
#pragma once
#include <nccl.h>

namespace djets {

struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*);
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
  const char* (*GetErrorString)(ncclResult_t);
  ncclResult_t (*GetVersion)(int*);
};

// nullptr (and *why filled) when no NCCL library could be bound.
const NcclApi* nccl_api(const char** why);

}  // namespace djets
