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

// ---- peer-memory mailboxes --------------------------------------------------------------------------------
namespace {
struct Peers {
    void* own = nullptr;                       // cudaMalloc'ed: tables, then flags, then the failure word
    void* opened[kMaxPeers] = {nullptr};       // peers' mailboxes (cudaIpcOpenMemHandle)
    bool ready = false;
    unsigned long long seq = 0;
} peers;
constexpr size_t kMailTableBytes = 2 * (size_t)kMaxPeers * kPeerDoubles * sizeof(double);
constexpr size_t kMailFlagBytes = 2 * (size_t)kMaxPeers * kPeerFlags * sizeof(unsigned long long);
constexpr size_t kMailBytes = kMailTableBytes + kMailFlagBytes + 256;
void peers_close() {
    for (int q = 0; q < kMaxPeers; ++q)
        if (peers.opened[q]) {
            cudaIpcCloseMemHandle(peers.opened[q]);
            peers.opened[q] = nullptr;
        }
    if (peers.own) {
        cudaFree(peers.own);
        peers.own = nullptr;
    }
    peers.ready = false;
}
}  // namespace

bool peer_exchange_ready() { return peers.ready && ctx().world > 1 && ctx().world <= kMaxPeers; }
int peer_exchange_next(PeerExchange* px) {
    if (!peer_exchange_ready()) return fail(DPR_ERR_STATE, "peer mailboxes are not attached");
    px->world = ctx().world;
    px->rank = ctx().rank;
    px->seq = ++peers.seq;
    for (int q = 0; q < kMaxPeers; ++q) {
        char* base = reinterpret_cast<char*>(q == ctx().rank ? peers.own : peers.opened[q]);
        px->tables[q] = q < ctx().world ? reinterpret_cast<double*>(base) : nullptr;
        px->flags[q] = q < ctx().world ? reinterpret_cast<unsigned long long*>(base + kMailTableBytes) : nullptr;
    }
    px->failed = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(peers.own) + kMailTableBytes + kMailFlagBytes);
    return DPR_OK;
}
int peer_exchange_check() {
    if (!peers.ready) return DPR_OK;
    int32_t flag = 0;
    DPR_CUDA(cudaMemcpyAsync(&flag, reinterpret_cast<char*>(peers.own) + kMailTableBytes + kMailFlagBytes, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    if (flag) return fail(DPR_ERR_NCCL, "peer-memory exchange: a rank's tables did not arrive (is every rank running the same sequence of calls?)");
    return DPR_OK;
}
}  // namespace dpr

using namespace dpr;
extern "C" {
// this rank's mailbox: allocated here, its 64-byte CUDA IPC handle handed to the caller, who gathers the handles of all ranks
int dpr_comm_peer_handle(void* handle64) {
    if (!handle64) return fail(DPR_ERR_INVALID, "null handle buffer");
    DPR_TRY(ensure_ready());
    peers_close();
    DPR_CUDA(cudaMalloc(&peers.own, kMailBytes));
    DPR_CUDA(cudaMemset(peers.own, 0, kMailBytes));
    cudaIpcMemHandle_t h;
    DPR_CUDA(cudaIpcGetMemHandle(&h, peers.own));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(handle64, &h, sizeof(h));
    return DPR_OK;
}
// handles: world x 64 bytes in rank order (this rank's own entry is ignored).  After this call the per-iteration exchange of a
// row-sharded fit goes through the mailboxes instead of ncclAllReduce.
int dpr_comm_peer_attach(const void* handles, int32_t world) {
    if (!handles || world != ctx().world || world < 2) return fail(DPR_ERR_INVALID, "peer_attach: handles of the current communicator's ranks expected");
    if (world > kMaxPeers) return fail(DPR_ERR_INVALID, "peer_attach: more ranks than mailbox slots");
    if (!peers.own) return fail(DPR_ERR_STATE, "peer_attach: dpr_comm_peer_handle was not called");
    for (int q = 0; q < world; ++q) {
        if (q == ctx().rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, reinterpret_cast<const char*>(handles) + (size_t)q * 64, 64);
        DPR_CUDA(cudaIpcOpenMemHandle(&peers.opened[q], h, cudaIpcMemLazyEnablePeerAccess));
    }
    peers.seq = 0;
    peers.ready = true;
    return DPR_OK;
}
int dpr_comm_peer_detach(void) {
    peers_close();
    return DPR_OK;
}
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
    peers_close();
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
