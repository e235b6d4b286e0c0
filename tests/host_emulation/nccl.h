// TEST INFRASTRUCTURE ONLY -- see cuda_runtime.h in this directory.  The host emulation is single
// rank: communicators cannot be created.
#pragma once
#include "cuda_runtime.h"
typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclInvalidUsage = 5 };
typedef struct rj_emu_comm *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum ncclDataType_t { ncclInt = 2, ncclDouble = 8 };
enum ncclRedOp_t { ncclMax = 2, ncclMin = 3 };
inline const char *ncclGetErrorString(ncclResult_t) { return "NCCL is not available in the host emulation"; }
inline ncclResult_t ncclGetUniqueId(ncclUniqueId *) { return ncclInvalidUsage; }
inline ncclResult_t ncclCommInitRank(ncclComm_t *, int, ncclUniqueId, int) { return ncclInvalidUsage; }
inline ncclResult_t ncclCommDestroy(ncclComm_t) { return ncclSuccess; }
inline ncclResult_t ncclGroupStart() { return ncclInvalidUsage; }
inline ncclResult_t ncclGroupEnd() { return ncclInvalidUsage; }
inline ncclResult_t ncclSend(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
inline ncclResult_t ncclRecv(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
inline ncclResult_t ncclAllReduce(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) { return ncclInvalidUsage; }
