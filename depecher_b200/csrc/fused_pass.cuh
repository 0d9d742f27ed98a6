// fused_pass.cuh — launch interface of the fused assignment / Lloyd pass (fused_pass.cu).
#pragma once
#include "common.cuh"

namespace dpr {

constexpr int kPassAssign = 0;  // labels only           (allocate_clusters, Clusterer.cpp:34-87)
constexpr int kPassLloyd = 1;   // labels + sums/counts   (+ the scatter-add of reevaluate_centers, :99-102)
constexpr int kPassSums = 2;    // sums/counts of the labels already stored (reevaluate_centers for given labels)

constexpr int kPassLloydDelta = 3;  // kernel instance only: kPassLloyd when every running problem accumulates deltas (all but a fit's first pass)

struct PassParams {
    const Problem* problems;
    int32_t n_problems;
    const int32_t* cta_tile_start;  // [grid + 1]: contiguous tile range of each CTA in the global tile order
    int32_t mode;
    int32_t compare_old;            // count labels that differ from the previous ones
    int32_t all_delta;              // Lloyd mode: no running problem needs full sums (the host knows: a finalize has run since the fits started) -> the delta-only instance
    int32_t bundles;                // some problems share their cells with the next ones (Problem::bundle > 1): the tensor-core pass runs its multi-set instance
    int32_t k_max;                  // rows of the widest centre matrix (table strides)
    int64_t part_stride;            // doubles per sum table (k_max * d)
    double* warp_sums;              // [grid * warps][k_max * d]   private per-warp tables (Lloyd / sums modes)
    long long* warp_counts;         // [grid * warps][k_max]
    double* cta_tables;             // [(grid + n_problems)][k_max*d + k_max + 1]: sums, counts, changed of each (CTA, problem)
    unsigned long long* recheck_count;  // diagnostics: cells decided by the fp64 path (nullable)
    // tensor-core pass, delta iterations: every consumer warp appends the cells whose label changed to its own list
    // (cell, new label | old label << 16), in tile order; delta_sums_kernel turns the lists into sum tables
    uint2* move_lists;              // [grid * warps][move_cap]
    int32_t* move_counts;           // [(grid + n_problems) * warps]: entries of each (CTA, problem, warp)
    int64_t move_cap;               // entries per warp
    int32_t* move_overflow;         // set when a list ran out of room (a sizing bug: the host fails loudly)
    // scoring pass with fused contingency counts (score_pass.cu): problems 2c, 2c + 1 are the (ret1, ret2) pair of penalty cell c;
    // the pass adds the (label under 2c, label under 2c + 1) counts of every cell into table c ([pair_kk x pair_kk] each)
    unsigned long long* pair_tables;   // nullable
    // filled from the plan
    int32_t warps, slots_per_warp, slot_floats, off_hdr, off_table, off_orig, off_scratch, scratch_bytes, off_slots;
};

struct PassPlan {
    int32_t warps, slots_per_warp, slot_floats, off_hdr, off_table, off_orig, off_scratch, scratch_bytes, off_slots;
    int32_t table_floats_cap;
    size_t smem_bytes;
};

// tensor-core pass (tc_pass.cu): shared-memory / tensor-memory plan
struct TcPlan {
    int32_t d, g8;          // markers; groups of 8 markers rounded up to whole K=16 steps
    int32_t wait_hint;      // suspend-time hint (ns) of the mbarrier waits
    uint32_t key_mul;       // 16 as a run-time value (top2_block16)
    int32_t groups;         // consumer groups per CTA (2..4: as many as tensor memory holds)
    int32_t npad;           // columns reserved per centre set (k_max rounded up to 32)
    int32_t tmem_cols;      // allocation: groups x (accumulator 2*npad + xh 4*g8 + xl 4*g8 columns), power of two
    int32_t stages, stage_floats;
    int32_t off_hdr, off_cc, off_orig, off_scratch, scratch_bytes, off_b, off_stages;
    int32_t mu_in_smem, off_mu;   // fp64 centres in shared memory for the exact re-decisions
    int32_t sets, set_bytes;      // centre sets (problems of a bundle) resident at once; hdr/cc/orig/b/mu of set s sit s * set_bytes further
    int32_t gtab, gtab_bytes, glist_bytes, off_gtab, off_glist;   // Lloyd tables of the consumer groups in shared memory (gtab = 0: per-warp tables in global memory)
    size_t smem_bytes;
    size_t table_floats;    // floats of Problem::tc_table
};
// returns 1 when the shape fits the tensor-core pass (k_max <= 128, >= 2 consumer groups in the 512 TMEM columns, >= 3 stages)
// sets_wanted: 1 for Lloyd sessions; assignment-only sessions may hold up to kTcMaxSets centre sets that share the cells
constexpr int kTcMaxSets = 4;
int plan_tc_pass(int d, int k_max, int sets_wanted, TcPlan* plan);
int launch_tc_pass(const TcPlan& plan, const PassParams& P, int grid, cudaStream_t stream);
// sums/counts of the moved cells recorded by a tensor-core Lloyd pass in delta mode (tc_pass.cu)
int launch_delta_sums(const TcPlan& plan, const PassParams& P, int grid, cudaStream_t stream);

// scoring pass (score_pass.cu): up to 4 centre sets stacked along N per tile, role-split CTA (split / products / reduce)
constexpr int kScMaxStages = 10;
constexpr int kScMaxSets = 4;
struct ScorePlan {
    int32_t d, k16;          // markers; K = 16 * k16 per product chain
    int32_t ncol_max;        // accumulator columns reserved per set (k_max rounded up to 16)
    int32_t sets;            // centre sets per bundle: 4 (ncol_max <= 32) or 2; 1 for a plain assignment
    int32_t b_rows;          // rows of the stacked B operands = sets * ncol_max
    int32_t acols;           // TMEM columns of one fp16 A operand
    int32_t wait_hint;
    int32_t kk;              // width of the fused contingency tables (k_max + 1); 0: not fused
    uint32_t key_mul;        // 16, as a run-time value: ptxas turns a multiplication by a constant 16 into LEA (the half-rate ALU pipe, which
                             //   the reduction saturates first); a register factor stays an IMAD on the FMA pipe
    int32_t stages, stage_floats;
    int32_t off_sets, off_cc, off_orig, off_bh, off_bl, off_pairs, off_stages;
    size_t smem_bytes;
};
// returns 1 when the shape fits (d <= 64, k_max <= 64); fuse_kk > 0 asks for the contingency tables
// of the pairs in shared memory (plan->kk stays 0 when they do not fit)
int plan_score_pass(int d, int k_max, int sets_wanted, int fuse_kk, ScorePlan* plan);
int launch_score_pass(const ScorePlan& plan, const PassParams& P, int grid, cudaStream_t stream);

// returns 1 and fills the plan when a d-marker slot per warp fits in shared memory, 0 otherwise
int plan_pass(int d, int k_max, PassPlan* plan);
int launch_fused_pass(const PassPlan& plan, PassParams P, int grid, cudaStream_t stream);

}  // namespace dpr
