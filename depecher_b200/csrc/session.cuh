// session.cuh — host-side owner of everything one batch of problems needs on the device.
//
// A Session holds P problems (data pointer + centre set each), the partition of their tiles over the
// persistent CTAs, the partial-sum tables and the Lloyd-loop state.  It is the B200 counterpart of one
// `Clusterer` instance of the reference (src/Clusterer.h:43-73), batched: the reference runs one fit at a
// time per R worker process (R/depechePenaltyOpt.R:50-54), here every concurrent fit shares each launch.
#pragma once
#include <vector>

#include "centres.cuh"
#include "common.cuh"
#include "fused_pass.cuh"

namespace dpr {

struct ProblemSpec {
    const float* x;        // tile-major resident cells
    const float* tile_xx;  // per-tile bound of ||x||^2
    const float* xx_max;   // one float: largest finite tile bound of the source matrix
    int64_t n;
    int32_t k;
    double reg;
    int32_t no_zero;
    int32_t* labels;   // device, >= n ints (nullable: session allocates when want_labels)
    const float* rows = nullptr;   // nullable: cell-major copy of x (dpr_dataset::rows)
};

class Session {
public:
    Session() = default;
    ~Session();
    Session(const Session&) = delete;
    Session& operator=(const Session&) = delete;

    // d is common to all problems.  want_sums: allocate the Lloyd partial tables.
    // pair_kk > 0 (assignment-only sessions whose problems 2c, 2c + 1 are the pairs of a penalty search): ask for the contingency
    // counts of each pair to be taken inside the pass, tables pair_kk wide (fuse_pairs tells whether the plan granted it)
    int init(const std::vector<ProblemSpec>& specs, int d, bool want_sums, int pair_kk = 0);
    int set_centres(int p, const double* mu_rowmajor_host);        // H2D of one centre matrix
    int set_centres_device(int p, const double* mu_rowmajor_dev);   // D2D
    int get_centres(int p, double* mu_rowmajor_host);
    int prepare();                                                   // chunk tables for the current centres
    int pass(int mode, bool compare_old);                            // one fused pass over every problem
    int reduce_and_finalize(int iter, int max_iter);                 // sums -> centres (+ convergence)
    // find_centers loop (Clusterer.cpp:244-252) for all problems; labels must be zeroed by the caller
    int lloyd(int max_iter);
    int fetch_state();                                               // d_problems -> h_problems (synchronises)
    void mark_restart() { next_pass_full_ = true; }                  // the caller has reset the problems' delta flags: the next Lloyd pass sums all cells again
    int check_lists();                                               // fails when a moved-cell list overflowed (synchronises)
    int zero_labels();

    int n_problems() const { return (int)h_problems.size(); }
    int d() const { return d_; }
    int k_max() const { return k_max_; }
    std::vector<Problem> h_problems;
    Problem* d_problems = nullptr;
    unsigned long long* d_recheck = nullptr;   // cells decided by the fp64 path (diagnostics)
    int grid = 0;
    PassPlan plan{};
    TcPlan tc_plan{};
    bool use_tc = false;       // shape fits the tcgen05 pass (tc_pass.cu); otherwise the FFMA pass runs
    ScorePlan score_plan{};
    bool use_score = false;    // assignment-only session with bundles that fits the scoring pass (score_pass.cu)
    bool fuse_pairs = false;   // ... which also takes the pairs' contingency counts
    unsigned long long* pair_tables = nullptr;   // device, [n_problems / 2][pair_kk^2], zeroed by the caller (fuse_pairs)
    bool allow_delta = true;   // let finalize switch a problem to delta accumulation when few labels change
    // Several ranks (dpr_comm_init): the problems of THIS session are row shards of fits that every rank runs in step, so each
    // iteration's tables are summed over the ranks.  Opt-in: only the row-sharded entry points (dpr_dataset_find_centers_from,
    // dpr_dataset_lloyd_iterations, dpr_dataset_reevaluate_centers) set it; every other session — bootstrap fits of a penalty search,
    // scoring, a rank-local sparse_k_means — stays local whatever the communicator (summing unrelated fits over ranks).
    bool row_sharded = false;
    bool exchanges() const { return row_sharded && ctx().world > 1; }

    // raw device buffers (exposed for the multi-GPU allreduce between reduce and finalize)
    double* seg = nullptr;
    int n_seg = 1;
    int64_t seg_entries = 0;   // per problem per segment

private:
    int d_ = 0, k_max_ = 0, total_tiles_ = 0;
    bool want_sums_ = false, has_bundles_ = false;
    bool next_pass_full_ = true;     // no finalize yet since the fits (re)started: their next Lloyd pass is a full one
    std::vector<void*> owned_;
    int32_t* d_cta_start_ = nullptr;
    double* warp_sums_ = nullptr;
    long long* warp_counts_ = nullptr;
    double* cta_tables_ = nullptr;
    int32_t* d_n_active_ = nullptr;
    uint2* move_lists_ = nullptr;      // tensor-core pass, delta iterations: moved cells per consumer warp
    int32_t* move_counts_ = nullptr;
    int32_t* move_overflow_ = nullptr;
    int64_t move_cap_ = 0;
    int32_t* h_n_active_ = nullptr;  // pinned
    int last_active_slot_ = 0;       // which of the two device counters the last finalize counted in
    bool fused_finalize() const;     // single GPU and the tables fit: reduce + finalize in one launch
    template <typename T> int alloc(T** p, size_t count, bool zero);
};

}  // namespace dpr
