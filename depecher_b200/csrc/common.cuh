// common.cuh — shared declarations of the depecher_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
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
constexpr int kSlotCells = 128;   // cells per warp slot (32 lanes x 4 cells, one LDS.128 per marker)
constexpr int kRowStride = 132;   // floats per marker row of a slot (528 B: 16 B aligned, bank-skewed)
constexpr int kMaxKC = 32;        // centres per register-tile chunk (4 cells x 32 centres = 128 accumulators)
constexpr int kWarpsPerCta = 8;
constexpr int kThreads = kWarpsPerCta * 32;
constexpr int kMaxChunks = 8;     // => fast path handles up to 256 active centres
constexpr int kSmemBudget = 227 * 1024;

// Centre table of one problem as the fused pass reads it (global memory, built on device by
// finalize_kernel / prepare_centres_kernel).  Floats are laid out per chunk:
//   cc[kcp]            ||c||^2 of each slot (dummy slots: +huge)
//   w[d][kcp]          -2 * c (marker-major, kcp = kc rounded up to 4)
struct CentreHeader {
    int32_t k;            // rows of mu
    int32_t k_act;        // active centres (no_zero: non-zero rows; else all rows, zero rows deduped)
    int32_t n_slots;      // k_act + the dummy slots that pad each chunk to an even size
    int32_t n_chunks;
    int32_t trivial;      // fewer than two active rows: every label 0 (Clusterer.cpp:50-53)
    int32_t force_exact;  // a centre holds NaN/Inf: decide every cell in fp64
    int32_t table_floats; // floats used in the chunk table
    float cc_max;         // max ||c||^2 over active centres
    int32_t chunk_kc[kMaxChunks];
    int32_t chunk_off[kMaxChunks];   // float offset of the chunk inside the table
    int32_t chunk_base[kMaxChunks];  // first active slot of the chunk
};

// One problem = one data matrix + one centre set (+ the state of its Lloyd loop).  Device-resident array.
struct Problem {
    const float* x;       // column-major fp32, column j at x + j*ld
    int64_t ld;           // floats per column (multiple of kSlotCells)
    int64_t n;            // cells
    int32_t d;
    int32_t k;
    int32_t* labels;      // n; read as "previous labels" and overwritten (nullable when not needed)
    double* mu;           // k x d row-major fp64 (exact centres; the recheck path reads these)
    CentreHeader* hdr;
    float* table;         // chunk table (table_cap floats)
    int32_t* orig;        // active slot -> row of mu (n_slots entries, dummies -1)
    int32_t first_tile;   // index of this problem's first tile in the global tile order
    int32_t n_tiles;
    // Lloyd-loop state (find_centers, Clusterer.cpp:244-252)
    double reg;           // full penalty
    int32_t no_zero;
    int32_t done;         // converged (labels unchanged and i > 20) or iteration cap reached
    int32_t n_assign;     // assignments performed so far
    int32_t pad_;
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
    // multi-GPU
    void* nccl_comm = nullptr;
    int rank = 0, world = 1;
};
Context& ctx();
int ensure_ready();   // picks the device, checks sm_100, creates the stream

inline void count_launch(int n = 1) { ctx().launches += n; }

}  // namespace dpr

struct dpr_dataset {
    int64_t n = 0;
    int32_t d = 0;
    int64_t ld = 0;
    float* x = nullptr;        // device
    int32_t* labels = nullptr; // device, ld entries (lazily allocated)
};
