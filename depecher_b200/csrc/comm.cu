// comm.cu — see comm.cuh.  The only collective on the hot path is the per-iteration allreduce of the
// K x (D+1) partial sums/counts when the rows of X are sharded over ranks; everything else
// (problem-grid split, dAllocate of row shards) needs no data-path communication.
#include "comm.cuh"

#include <dlfcn.h>

#include <cstring>

namespace dpr {
namespace {
typedef struct { char internal[128]; } NcclUniqueId;
typedef void* NcclComm;
enum { kNcclSum = 0, kNcclInt64 = 4, kNcclFloat64 = 8 };  // nccl.h: ncclSum=0, ncclInt64=4, ncclFloat64=8
struct Api {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
} api;

int load(const char* path) {
    if (api.lib) return DPR_OK;
    const char* names[] = {path, "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        if (!n) continue;
        api.lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (api.lib) break;
    }
    if (!api.lib) return fail(DPR_ERR_NCCL, std::string("cannot dlopen NCCL: ") + dlerror());
    api.GetUniqueId = (int (*)(NcclUniqueId*))dlsym(api.lib, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(NcclComm*, int, NcclUniqueId, int))dlsym(api.lib, "ncclCommInitRank");
    api.CommDestroy = (int (*)(NcclComm))dlsym(api.lib, "ncclCommDestroy");
    api.AllReduce = (int (*)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t))dlsym(api.lib, "ncclAllReduce");
    api.GetErrorString = (const char* (*)(int))dlsym(api.lib, "ncclGetErrorString");
    if (!api.GetUniqueId || !api.CommInitRank || !api.CommDestroy || !api.AllReduce)
        return fail(DPR_ERR_NCCL, "NCCL library lacks a required symbol");
    return DPR_OK;
}
int check(int rc, const char* what) {
    if (rc == 0) return DPR_OK;
    return fail(DPR_ERR_NCCL, std::string(what) + ": " + (api.GetErrorString ? api.GetErrorString(rc) : "error"));
}
}  // namespace

int comm_allreduce_sum_f64(double* buf, size_t count) {
    if (ctx().world <= 1) return DPR_OK;
    if (!ctx().nccl_comm) return fail(DPR_ERR_STATE, "dpr_comm_init was not called");
    count_launch();
    return check(api.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, ctx().nccl_comm, ctx().stream), "ncclAllReduce");
}
int comm_allreduce_sum_i64(long long* buf, size_t count) {
    if (ctx().world <= 1) return DPR_OK;
    if (!ctx().nccl_comm) return fail(DPR_ERR_STATE, "dpr_comm_init was not called");
    count_launch();
    return check(api.AllReduce(buf, buf, count, kNcclInt64, kNcclSum, ctx().nccl_comm, ctx().stream), "ncclAllReduce");
}
}  // namespace dpr

using namespace dpr;
extern "C" {
int dpr_comm_unique_id(void* id128, const char* nccl_path) {
    if (!id128) return fail(DPR_ERR_INVALID, "null id buffer");
    DPR_TRY(load(nccl_path));
    NcclUniqueId id;
    DPR_TRY(check(api.GetUniqueId(&id), "ncclGetUniqueId"));
    memcpy(id128, &id, sizeof(id));
    return DPR_OK;
}
int dpr_comm_init(int32_t rank, int32_t world, const void* id128, const char* nccl_path) {
    if (world < 1 || rank < 0 || rank >= world) return fail(DPR_ERR_INVALID, "bad rank/world");
    DPR_TRY(ensure_ready());
    if (world == 1) {
        ctx().rank = 0;
        ctx().world = 1;
        return DPR_OK;
    }
    if (!id128) return fail(DPR_ERR_INVALID, "null unique id");
    DPR_TRY(load(nccl_path));
    if (ctx().nccl_comm) {   // re-initialisation: drop the previous communicator first
        api.CommDestroy(ctx().nccl_comm);
        ctx().nccl_comm = nullptr;
    }
    NcclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    NcclComm comm = nullptr;
    DPR_TRY(check(api.CommInitRank(&comm, world, id, rank), "ncclCommInitRank"));
    ctx().nccl_comm = comm;
    ctx().rank = rank;
    ctx().world = world;
    return DPR_OK;
}
int dpr_comm_destroy(void) {
    if (ctx().nccl_comm) {
        api.CommDestroy(ctx().nccl_comm);
        ctx().nccl_comm = nullptr;
    }
    ctx().rank = 0;
    ctx().world = 1;
    return DPR_OK;
}
int dpr_comm_info(int32_t* rank, int32_t* world) {
    if (rank) *rank = ctx().rank;
    if (world) *world = ctx().world;
    return DPR_OK;
}
}
