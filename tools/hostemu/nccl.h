// tools/hostemu/nccl.h -- TEST INFRASTRUCTURE: single-rank stand-in for NCCL (see cuda_runtime.h in this directory).
#pragma once
#include <cstring>
typedef int ncclResult_t; typedef void* ncclComm_t;
enum { ncclSuccess = 0, ncclInvalidArgument = 4 };
enum ncclDataType_t { ncclUint8 = 1, ncclInt32 = 2, ncclDouble = 8 };
enum ncclRedOp_t { ncclSum = 0, ncclMin = 3 };
struct ncclUniqueId { char internal[128]; };
inline ncclResult_t ncclGetUniqueId(ncclUniqueId* id) { std::memset(id, 0x5a, sizeof *id); return ncclSuccess; }
inline ncclResult_t ncclCommInitRank(ncclComm_t* c, int n, ncclUniqueId, int) { if (n != 1) return ncclInvalidArgument; *c = (void*)0x1; return ncclSuccess; }
inline ncclResult_t ncclCommDestroy(ncclComm_t) { return ncclSuccess; }
inline const char* ncclGetErrorString(ncclResult_t) { return "hostemu nccl: one rank only"; }
inline size_t hostemu_nccl_size(ncclDataType_t t) { return t == ncclDouble ? 8 : (t == ncclInt32 ? 4 : 1); }
inline ncclResult_t ncclAllReduce(const void* s, void* r, size_t n, ncclDataType_t t, ncclRedOp_t, ncclComm_t, void*) { if (s != r) std::memmove(r, s, n * hostemu_nccl_size(t)); return ncclSuccess; }
inline ncclResult_t ncclAllGather(const void* s, void* r, size_t n, ncclDataType_t t, ncclComm_t, void*) { if (s != r) std::memmove(r, s, n * hostemu_nccl_size(t)); return ncclSuccess; }
