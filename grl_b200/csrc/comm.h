// grl_b200 — NCCL entry points resolved at run time (dlopen), so that libgrl_b200.so itself has no link-time dependency on
// NCCL: it loads on a machine without it, and inside a PyTorch process it binds to the very libnccl.so.2 torch already loaded.
#pragma once
#include <nccl.h>

#include "api.h"

namespace grl {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*ReduceScatter)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
};

// The process-wide table (loaded on first use); NULL with the handle's error set when NCCL cannot be loaded.
const NcclApi* nccl_api(grl_handle* h);

#define GRL_NCCL(h, api, expr)                                                                         \
    do {                                                                                               \
        ncclResult_t _r = (expr);                                                                      \
        if (_r != ncclSuccess)                                                                         \
            return grl::set_error((h), GRL_ENCCL, "%s failed: %s (%s:%d)", #expr, (api)->GetErrorString(_r), \
                                  __FILE__, __LINE__);                                                 \
    } while (0)

}  // namespace grl
