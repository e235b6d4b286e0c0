// TEST INFRASTRUCTURE ONLY -- see cuda_runtime.h in this directory.
// In-process emulation of the few NCCL calls the module makes: the ranks of a communicator are
// host threads of one process (each with its own emulated device); point-to-point messages are
// eager copies into per-(source, destination) FIFO mailboxes, receives block until the matching
// message has been posted, all-reduces are a rendezvous of all ranks.  Per-peer ordering is FIFO
// like NCCL's.  (host_emulation.cc)
#pragma once
#include "cuda_runtime.h"

typedef int ncclResult_t;
enum { ncclSuccess = 0, ncclInvalidUsage = 5 };
struct rj_emu_comm;
typedef rj_emu_comm *ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum ncclDataType_t { ncclInt = 2, ncclDouble = 8 };
enum ncclRedOp_t { ncclMax = 2, ncclMin = 3 };

const char *ncclGetErrorString(ncclResult_t);
ncclResult_t ncclGetUniqueId(ncclUniqueId *id);
ncclResult_t ncclCommInitRank(ncclComm_t *comm, int n_ranks, ncclUniqueId id, int rank);
ncclResult_t ncclCommDestroy(ncclComm_t comm);
ncclResult_t ncclGroupStart();
ncclResult_t ncclGroupEnd();
ncclResult_t ncclSend(const void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t);
ncclResult_t ncclRecv(void *buffer, size_t count, ncclDataType_t type, int peer, ncclComm_t comm, cudaStream_t);
ncclResult_t ncclAllReduce(const void *send, void *recv, size_t count, ncclDataType_t type, ncclRedOp_t op, ncclComm_t comm, cudaStream_t);
