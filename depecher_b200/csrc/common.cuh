// common.cuh — shared declarations of the depecher_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <cstdlib>
#include <string>
#include <utility>
#include <vector>

#include "../../include/depecher_b200.h"

namespace dpr {

// ---- error plumbing -----------------------------------------------------------------------
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
#define DPR_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return ::dpr::fail(DPR_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define DPR_TRY(expr)                \
    do {                             \
        int _rc = (expr);            \
        if (_rc != DPR_OK) return _rc; \
    } while (0)

// ---- geometry of the fused pass --------------------------------------------------------------
constexpr int kSlotCells = 128;   // cells per tile (= one warp slot of the FFMA pass = the M of one tcgen05 product)
constexpr int kGroupFloats = kSlotCells * 4 + 4;   // one 4-marker group of a tile: 128 cells x 4 markers, + 16 B so that groups shift by 4 banks
constexpr int kMaxKC = 32;        // centres per register-tile chunk (4 cells x 32 centres = 128 accumulators)
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMaxChunks = 16;    // => the FFMA pass handles up to 480 centres (depecheAllData fits K = 3k: depeche(k = 100) needs 300, R/depecheAllData.R:30-31)
constexpr int kSmemBudget = 227 * 1024;

// Resident layout of a cell matrix ("tile-major"): cells are grouped in tiles of 128 consecutive cells, a tile is
// contiguous in HBM (one bulk-async copy moves it) and holds ceil(d/4) marker groups; inside a group each cell owns
// 16 consecutive bytes = its 4 markers (markers beyond d are zero).  This is exactly the K-major, no-swizzle
// canonical operand layout of tcgen05.mma (core matrix = 8 cells x 16 B, 128 B between 8-cell blocks, one group
// between the two halves of a K=8 step), so the tensor core reads a landed tile as it is; the FFMA pass reads a
// cell's 4 markers with one LDS.128 (lane l owns cells l, l+32, l+64, l+96: conflict-free), and the transposed
// reads of the sum phase (lane = marker, fixed cell) hit 32 different banks thanks to the 16 B group padding.
__host__ __device__ __forceinline__ int tile_groups(int d) { return (d + 3) >> 2; }
__host__ __device__ __forceinline__ size_t tile_floats(int d) { return (size_t)tile_groups(d) * kGroupFloats; }
__host__ __device__ __forceinline__ int x_slot_index(int cell_in_slot, int j) { return (j >> 2) * kGroupFloats + cell_in_slot * 4 + (j & 3); }
__host__ __device__ __forceinline__ size_t x_index(int64_t cell, int j, int d) {
    return (size_t)(cell >> 7) * tile_floats(d) + (size_t)x_slot_index((int)(cell & 127), j);
}

// Centre table of one problem as the fused pass reads it (global memory, built on device by
// finalize_kernel / prepare_centres_kernel).  Floats are laid out per chunk:
//   cc[kcp]            ||c||^2 of each slot (dummy slots: +huge)
//   w[d4][kcp]         -2 * c (marker-major, kcp = kc rounded up to 4, d4 = d rounded up to 4 with zero rows)
struct CentreHeader {
    int32_t k;            // rows of mu
    int32_t k_act;        // active centres (no_zero: non-zero rows; else all rows, zero rows deduped)
    int32_t n_slots;      // k_act + the dummy slots that pad each chunk to an even size
    int32_t n_chunks;
    int32_t trivial;      // fewer than two active rows: every label 0 (Clusterer.cpp:50-53)
    int32_t force_exact;  // a centre holds NaN/Inf: decide every cell in fp64
    int32_t table_floats; // floats used in the chunk table
    int32_t tc_n;         // tensor-core path: active centres rounded up to a multiple of 16 (the N of the products)
    float tc_scale_x;     // power of two the cells are multiplied by before the fp16 split
    float tc_unscale;     // 1 / (cell scale * centre scale): brings a product back to -2 x.c
    float tc_abs1, tc_abs2;   // absolute part of the error bound: tau += tc_abs1 * sqrt(xx_tile) + tc_abs2
    float cc_max;         // max ||c||^2 over active centres
    int32_t chunk_kc[kMaxChunks];
    int32_t chunk_off[kMaxChunks];   // float offset of the chunk inside the table
    int32_t chunk_base[kMaxChunks];  // first active slot of the chunk
};

// One problem = one data matrix + one centre set (+ the state of its Lloyd loop).  Device-resident array.
struct Problem {
    const float* x;       // resident fp32 cells, tile-major (see x_index)
    const float* tile_xx; // per tile: an upper bound of ||x||^2 over its cells (enters the near-tie threshold)
    const float* xx_max;  // one float: the largest finite tile bound of the matrix the cells come from (fp16 scaling of the tensor-core pass)
    const float* xrows;   // nullable: the same cells cell-major (row c at c * 4 * tile_groups(d) floats) — what delta_sums_kernel gathers moved cells from
    int64_t n;            // cells
    int32_t d;
    int32_t k;
    int32_t* labels;      // n; read as "previous labels" and overwritten (nullable when not needed)
    double* mu;           // k x d row-major fp64 (exact centres; the recheck path reads these)
    CentreHeader* hdr;
    float* table;         // chunk table (table_cap floats)
    int32_t* orig;        // active slot -> row of mu (n_slots entries, dummies -1)
    float* tc_table;      // tensor-core path (nullable): cc[tc_npad], then the scaled -2*c operand split in fp16 high / low parts
    int32_t* tc_orig;     // tensor-core path: column -> row of mu (tc_npad entries, padding -1)
    int32_t tc_npad;      // columns reserved per problem (k_max rounded up to 32)
    int32_t first_tile;   // index of this problem's first tile in the global tile order
    int32_t n_tiles;
    int32_t bundle;       // assignment-only sessions of the tensor-core pass: this problem and the next bundle - 1 ones score the SAME
                          //   cells against different centre sets; the leader owns the tiles (members: n_tiles = 0, first_tile = the
                          //   leader's end), a landed tile is split once and multiplied with every set's centres.  0 / 1: on its own
    // Lloyd-loop state (find_centers, Clusterer.cpp:244-252)
    double reg;           // full penalty
    int32_t no_zero;
    int32_t done;         // converged (labels unchanged and i > 20) or iteration cap reached
    int32_t n_assign;     // assignments performed so far
    int32_t delta;        // next Lloyd pass accumulates only the cells whose label changes (sums += delta)
    long long changed;    // labels changed in the last assignment
    double* sums;         // k x d  (this rank's, then all ranks')
    long long* counts;    // k
};

// ---- process-wide context ----------------------------------------------------------------------
struct Context {
    bool ready = false;
    int device = 0;
    int sms = 0;
    cudaStream_t stream = nullptr;       // stream every launch goes to
    cudaStream_t own_stream = nullptr;
    int ari_mode = DPR_ARI_EXACT;
    int64_t launches = 0;
    unsigned long long* d_recheck = nullptr;   // cells decided by the exact fp64 path since the last reset (diagnostics)
    bool force_ffma = false;             // diagnostics: keep the tensor-core pass off (dpr_set_pass_kernel)
    int32_t* pinned_word = nullptr;      // pinned host word the Lloyd loop polls the active-problem count through
    // optional per-launch timing of the fused pass (dpr_pass_timing)
    bool time_passes = false;
    std::vector<std::pair<cudaEvent_t, cudaEvent_t>> pass_events;
    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
};
Context& ctx();
int ensure_ready();
int dev_alloc(void** p, size_t bytes);   // stream-ordered, pooled (api.cu)
void dev_free(void* p);   // picks the device, checks sm_100, creates the stream

inline void count_launch(int n = 1) { ctx().launches += n; }

// Launch with programmatic stream serialization: the kernels of a Lloyd step (pass -> sums of the moved cells -> reduce + finalize -> next
// pass) each begin with griddepcontrol.launch_dependents + griddepcontrol.wait, so the next kernel's CTAs are placed on the SMs as the
// previous kernel's CTAs leave them and only wait for its completion (and the visibility of its writes) there: the launch latency between
// dependent kernels is hidden.  DPR_NO_PDL=1: plain launches (A/B aid).
inline bool pdl_enabled() {
    static const bool on = getenv("DPR_NO_PDL") == nullptr;
    return on;
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define DPR_GRID_DEPENDENCY() asm volatile("griddepcontrol.launch_dependents;\n\tgriddepcontrol.wait;" ::: "memory")

}  // namespace dpr

struct dpr_dataset {
    int64_t n = 0;
    int32_t d = 0;
    int64_t n_pad = 0;         // n rounded up to whole tiles
    float* x = nullptr;        // device, tile-major (dpr::x_index), (n_pad / 128) * tile_floats(d) floats
    float* tile_xx = nullptr;  // device, n_pad / 128 + 1 floats: max ||x||^2 of each tile (rounded up), then the largest finite one
    int32_t* labels = nullptr; // device, n_pad entries (lazily allocated)
    float* rows = nullptr;     // device, n_pad x 4 * tile_groups(d) floats: a cell-major copy for row gathers, built on the first Lloyd fit of a
                               //   large matrix (a moved cell's row is one contiguous 16 * tile_groups(d) bytes there, tile_groups(d) scattered sectors in x)
    bool rows_tried = false;
};
