// comm.cuh — NCCL plumbing (one process per GPU).  libnccl is dlopen()ed so the library has no link-time
// dependency on it and shares the copy a host framework (torch.distributed) may already have loaded.
#pragma once
#include "common.cuh"
namespace dpr {
int comm_allreduce_sum_f64(double* buf, size_t count);   // in place, on ctx().stream
int comm_allreduce_sum_i64(long long* buf, size_t count);
}
