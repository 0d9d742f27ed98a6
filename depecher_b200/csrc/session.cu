// session.cu — see session.cuh
#include "session.cuh"

#include <algorithm>
#include <cstring>
#include <cstdlib>

#include "comm.cuh"

namespace dpr {

Session::~Session() {
    for (void* p : owned_) dev_free(p);
}

template <typename T>
int Session::alloc(T** p, size_t count, bool zero) {
    void* q = nullptr;
    const size_t bytes = std::max<size_t>(count, 1) * sizeof(T);
    DPR_TRY(dev_alloc(&q, bytes));
    owned_.push_back(q);
    if (zero) DPR_CUDA(cudaMemsetAsync(q, 0, bytes, ctx().stream));
    *p = reinterpret_cast<T*>(q);
    return DPR_OK;
}

int Session::init(const std::vector<ProblemSpec>& specs, int d, bool want_sums, int pair_kk) {
    DPR_TRY(ensure_ready());
    if (specs.empty()) return fail(DPR_ERR_INVALID, "no problems");
    d_ = d;
    want_sums_ = want_sums;
    k_max_ = 0;
    for (const auto& s : specs) {
        if (s.k < 1 || s.n < 1) return fail(DPR_ERR_INVALID, "problem with k < 1 or n < 1");
        k_max_ = std::max(k_max_, (int)s.k);
    }
    if (!plan_pass(d, k_max_, &plan))
        return fail(DPR_ERR_INVALID, "shape not supported by the sm_100a kernels: d = " + std::to_string(d) + ", k = " +
                                         std::to_string(k_max_) + " (one 128-cell slot of d markers plus the centre table must fit in 227 KB)");
    // assignment-only sessions: how many consecutive problems score the same cells (they can share the landed tiles: bundles)
    int max_run = 1;
    if (!want_sums)
        for (size_t p = 1, run = 1; p < specs.size(); ++p) {
            run = (specs[p].x == specs[p - 1].x && specs[p].n == specs[p - 1].n && specs[p].tile_xx == specs[p - 1].tile_xx && specs[p].xx_max == specs[p - 1].xx_max) ? run + 1 : 1;
            max_run = std::max(max_run, (int)run);
        }
    use_tc = !ctx().force_ffma && plan_tc_pass(d, k_max_, std::min(max_run, kTcMaxSets), &tc_plan) != 0;
    // several sets over the same cells: the scoring pass (its own bundle size; it reads the tensor-core tables of the problems)
    use_score = use_tc && !want_sums && plan_score_pass(d, k_max_, max_run, pair_kk, &score_plan) != 0;
    fuse_pairs = use_score && pair_kk > 0 && score_plan.kk == pair_kk && specs.size() % 2 == 0;
    const int bundle_sets = use_score ? score_plan.sets : tc_plan.sets;
    const int P = (int)specs.size();
    h_problems.assign(P, Problem{});
    // pooled per-problem storage
    const size_t mu_sz = (size_t)k_max_ * d, tab_sz = (size_t)plan.table_floats_cap, orig_sz = (size_t)k_max_ + 2 * kMaxChunks;
    double *mu_pool, *sums_pool;
    long long* cnt_pool;
    CentreHeader* hdr_pool;
    float* tab_pool;
    int32_t* orig_pool;
    DPR_TRY(alloc(&mu_pool, mu_sz * P, true));
    DPR_TRY(alloc(&sums_pool, mu_sz * P, true));
    DPR_TRY(alloc(&cnt_pool, (size_t)k_max_ * P, true));
    DPR_TRY(alloc(&hdr_pool, (size_t)P, true));
    DPR_TRY(alloc(&tab_pool, tab_sz * P, true));
    DPR_TRY(alloc(&orig_pool, orig_sz * P, true));
    float* tc_tab_pool = nullptr;
    int32_t* tc_orig_pool = nullptr;
    if (use_tc) {
        DPR_TRY(alloc(&tc_tab_pool, tc_plan.table_floats * P, true));
        DPR_TRY(alloc(&tc_orig_pool, (size_t)tc_plan.npad * P, true));
    }
    int tile = 0, leader = 0, bundle_left = 0;
    for (int p = 0; p < P; ++p) {
        Problem& pr = h_problems[p];
        const ProblemSpec& s = specs[p];
        pr.x = s.x;
        pr.tile_xx = s.tile_xx;
        pr.xx_max = s.xx_max;
        pr.xrows = s.rows;
        pr.n = s.n;
        pr.d = d;
        pr.k = s.k;
        pr.labels = s.labels;
        pr.mu = mu_pool + mu_sz * p;
        pr.hdr = hdr_pool + p;
        pr.table = tab_pool + tab_sz * p;
        pr.orig = orig_pool + orig_sz * p;
        if (use_tc) {
            pr.tc_table = tc_tab_pool + tc_plan.table_floats * p;
            pr.tc_orig = tc_orig_pool + (size_t)tc_plan.npad * p;
            pr.tc_npad = tc_plan.npad;
        }
        // assignment-only sessions of the tensor-core pass: consecutive problems over the same cells form a bundle — the leader
        // owns the tiles, the others sit (empty) at its end and ride along in the kernel (Problem::bundle)
        const bool member = use_tc && !want_sums && p > 0 && bundle_left > 0 && s.x == specs[p - 1].x && s.n == specs[p - 1].n &&
                            s.tile_xx == specs[p - 1].tile_xx && s.xx_max == specs[p - 1].xx_max;
        if (member) {
            pr.first_tile = tile;
            pr.n_tiles = 0;
            pr.bundle = 0;
            ++h_problems[leader].bundle;
            has_bundles_ = true;
            --bundle_left;
        } else {
            leader = p;
            bundle_left = use_tc && !want_sums ? bundle_sets - 1 : 0;
            pr.bundle = 1;
            pr.first_tile = tile;
            pr.n_tiles = (int)((s.n + kSlotCells - 1) / kSlotCells);
            tile += pr.n_tiles;
        }
        pr.reg = s.reg;
        pr.no_zero = s.no_zero;
        pr.sums = sums_pool + mu_sz * p;
        pr.counts = cnt_pool + (size_t)k_max_ * p;
    }
    total_tiles_ = tile;
    DPR_TRY(alloc(&d_problems, (size_t)P, false));
    DPR_CUDA(cudaMemcpyAsync(d_problems, h_problems.data(), sizeof(Problem) * P, cudaMemcpyHostToDevice, ctx().stream));
    // static partition: contiguous, equal tile counts per CTA
    const int per_cta_min = use_tc ? 4 * tc_plan.groups : plan.warps;  // a few tiles per CTA before using another one
    grid = std::min(ctx().sms, std::max(1, (total_tiles_ + per_cta_min - 1) / per_cta_min));
    std::vector<int32_t> start(grid + 1);
    for (int c = 0; c <= grid; ++c) start[c] = (int32_t)((int64_t)total_tiles_ * c / grid);
    DPR_TRY(alloc(&d_cta_start_, (size_t)grid + 1, false));
    DPR_CUDA(cudaMemcpyAsync(d_cta_start_, start.data(), sizeof(int32_t) * (grid + 1), cudaMemcpyHostToDevice, ctx().stream));
    if (!ctx().d_recheck) {   // one process-wide diagnostics counter (dpr_recheck_count)
        DPR_CUDA(cudaMalloc(&ctx().d_recheck, sizeof(unsigned long long)));
        DPR_CUDA(cudaMemsetAsync(ctx().d_recheck, 0, sizeof(unsigned long long), ctx().stream));
    }
    d_recheck = ctx().d_recheck;
    n_seg = std::min(16, grid);
    seg_entries = (int64_t)k_max_ * d + k_max_ + 1;
    if (want_sums) {
        const size_t wtabs = (size_t)grid * std::max(plan.warps, 16);   // either pass kernel: one private table per (CTA, warp)
        DPR_TRY(alloc(&warp_sums_, wtabs * mu_sz, true));
        DPR_TRY(alloc(&warp_counts_, wtabs * k_max_, true));
        DPR_TRY(alloc(&cta_tables_, (size_t)(grid + P) * seg_entries, true));
        DPR_TRY(alloc(&seg, (size_t)P * n_seg * seg_entries, true));
        DPR_TRY(alloc(&d_n_active_, (size_t)2, true));   // [0] alone on the two-kernel path; the fused reduce + finalize alternates between the two
        if (use_tc) {
            // lists of moved cells of the delta iterations: one per consumer warp, sized for "every cell of the warp's tiles moves".
            // A warp sees at most ceil(tiles of a problem range / groups) tiles per (CTA, problem) range, 32 cells of each.
            int64_t min_tiles = total_tiles_, max_n = 0;
            for (const auto& pr : h_problems) {
                min_tiles = std::min<int64_t>(min_tiles, pr.n_tiles);
                max_n = std::max<int64_t>(max_n, pr.n);
            }
            const int64_t range_tiles = (total_tiles_ + grid - 1) / grid + 1;
            const int64_t range_problems = std::min<int64_t>(P, range_tiles / std::max<int64_t>(min_tiles, 1) + 2);
            move_cap_ = (range_tiles / tc_plan.groups + range_problems + 1) * 32;
            if (max_n < (int64_t)1 << 32 && k_max_ <= 65535) {
                const size_t lists = (size_t)grid * 4 * tc_plan.groups;
                DPR_TRY(alloc(&move_lists_, lists * (size_t)move_cap_, false));
                DPR_TRY(alloc(&move_counts_, (size_t)(grid + P) * 4 * tc_plan.groups, true));
                DPR_TRY(alloc(&move_overflow_, (size_t)1, true));
            } else {
                allow_delta = false;   // cell indices would not fit the 32-bit list entries: every iteration sums all cells
            }
        }
        if (!ctx().pinned_word) DPR_CUDA(cudaMallocHost(&ctx().pinned_word, sizeof(int32_t)));   // kept for the life of the process
        h_n_active_ = ctx().pinned_word;
    }
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));  // `start`/h_problems are stack/heap temporaries of this call
    return DPR_OK;
}

int Session::set_centres(int p, const double* mu) {
    DPR_CUDA(cudaMemcpyAsync(h_problems[p].mu, mu, sizeof(double) * h_problems[p].k * d_, cudaMemcpyHostToDevice, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}
int Session::set_centres_device(int p, const double* mu) {
    DPR_CUDA(cudaMemcpyAsync(h_problems[p].mu, mu, sizeof(double) * h_problems[p].k * d_, cudaMemcpyDeviceToDevice, ctx().stream));
    return DPR_OK;
}
int Session::get_centres(int p, double* mu) {
    DPR_CUDA(cudaMemcpyAsync(mu, h_problems[p].mu, sizeof(double) * h_problems[p].k * d_, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}

int Session::prepare() { return launch_prepare_centres(d_problems, n_problems(), plan.table_floats_cap, ctx().stream); }

int Session::pass(int mode, bool compare_old) {
    if (mode != kPassAssign && !want_sums_) return fail(DPR_ERR_STATE, "session was created without sum tables");
    PassParams P{};
    P.problems = d_problems;
    P.n_problems = n_problems();
    P.cta_tile_start = d_cta_start_;
    P.mode = mode;
    P.compare_old = compare_old ? 1 : 0;
    P.bundles = has_bundles_ ? 1 : 0;
    P.all_delta = (mode == kPassLloyd && !next_pass_full_ && allow_delta && move_lists_) ? 1 : 0;   // every running problem has had a finalize since its fit began
    P.k_max = k_max_;
    P.part_stride = (int64_t)k_max_ * d_;
    P.warp_sums = warp_sums_;
    P.warp_counts = warp_counts_;
    P.cta_tables = cta_tables_;
    P.recheck_count = d_recheck;
    P.move_lists = move_lists_;
    P.move_counts = move_counts_;
    P.move_cap = move_cap_;
    P.move_overflow = move_overflow_;
    P.pair_tables = fuse_pairs ? pair_tables : nullptr;
    if (use_score && has_bundles_ && mode == kPassAssign) {   // (one set per tile: tc_pass_kernel is the faster of the two — 0.279 vs 0.303 ms at 10M x 40)
        if (fuse_pairs && !pair_tables) return fail(DPR_ERR_STATE, "scoring session: the pair tables were not set");
        return launch_score_pass(score_plan, P, grid, ctx().stream);
    }
    if (use_tc && mode != kPassSums) {
        DPR_TRY(launch_tc_pass(tc_plan, P, grid, ctx().stream));
        // delta iterations (every Lloyd pass of a fit but the first, where it finds nothing to do): the pass recorded the moved cells, this sums their rows
        if (mode == kPassLloyd && move_lists_ && allow_delta) return launch_delta_sums(tc_plan, P, grid, ctx().stream);
        return DPR_OK;
    }
    return launch_fused_pass(plan, P, grid, ctx().stream);
}

int Session::reduce_and_finalize(int iter, int max_iter) {
    next_pass_full_ = false;   // finalize switches every running problem to delta accumulation (when allowed)
    if (fused_finalize()) {   // one GPU: nothing to exchange between the two steps; several: exchanged inside, through peer memory
        last_active_slot_ = iter & 1;
        return launch_reduce_finalize(d_problems, n_problems(), d_cta_start_, grid, cta_tables_, k_max_, d_, iter, max_iter, plan.table_floats_cap,
                                      allow_delta ? 1 : 0, d_n_active_, exchanges(), ctx().stream);
    }
    static const bool two_kernels = getenv("DPR_TWO_KERNEL_FINALIZE") != nullptr;   // A/B aid
    if (!two_kernels && !exchanges() && reduce_finalize_fits(k_max_, d_)) {   // many problems, one GPU: one CTA per problem reduces and finalizes
        last_active_slot_ = iter & 1;
        return launch_finalize_direct(d_problems, n_problems(), d_cta_start_, grid, cta_tables_, k_max_, d_, iter, max_iter, plan.table_floats_cap,
                                      allow_delta ? 1 : 0, d_n_active_, ctx().stream);
    }
    last_active_slot_ = 0;
    DPR_TRY(launch_reduce_partials(d_problems, n_problems(), d_cta_start_, grid, k_max_, d_, cta_tables_, seg, n_seg, d_n_active_, ctx().stream));
    if (exchanges()) DPR_TRY(comm_allreduce_sum_f64(seg, (size_t)n_problems() * n_seg * seg_entries));
    return launch_finalize(d_problems, n_problems(), seg, n_seg, k_max_, d_, iter, max_iter, plan.table_floats_cap, allow_delta ? 1 : 0, d_n_active_,
                           ctx().stream);
}

bool Session::fused_finalize() const {
    static const bool off = getenv("DPR_TWO_KERNEL_FINALIZE") != nullptr;   // A/B aid
    // a few large problems: one launch fewer per iteration matters and the clusters all run at once.  Many small ones (the 440 bootstrap
    // fits of a penalty-search set): 8 CTAs per problem come in a dozen waves of gang-scheduled clusters — the two plain kernels are faster
    // (measured: 0.137 vs 0.259 s for the penalty search of depeche() on 10M x 40)
    if (off || n_problems() > 16 || !reduce_finalize_fits(k_max_, d_)) return false;
    if (!exchanges()) return true;
    return peer_exchange_ready() && reduce_finalize_peer_fits(n_problems(), k_max_, d_);   // several ranks: the exchange happens inside the launch
}

int Session::zero_labels() {
    for (auto& pr : h_problems)
        if (pr.labels) DPR_CUDA(cudaMemsetAsync(pr.labels, 0, sizeof(int32_t) * pr.n, ctx().stream));
    return DPR_OK;
}

int Session::lloyd(int max_iter) {
    // The convergence rule needs i > 20 (Clusterer.cpp:247): iterations 0..20 run back to back without
    // touching the host; after that the host looks at the active-problem counter once per iteration.
    // Batches of small fits (the bootstrap fits of a penalty search) are launch-latency-bound: an iteration is a few tens of
    // microseconds, a host round trip costs as much, and a converged problem is skipped by every kernel — they look only every
    // fourth iteration (at most three near-empty iterations at the end).  Large problems look every iteration: one more pass
    // over 10M cells costs 20 round trips.
    const int look_every = total_tiles_ >= 20000 ? 1 : 4;
    int last_active = n_problems();
    for (int i = 0; i < max_iter; ++i) {
        DPR_TRY(pass(kPassLloyd, true));
        DPR_TRY(reduce_and_finalize(i, max_iter));
        if (i > 20 && ((i - 21) % look_every == look_every - 1 || i + 1 == max_iter)) {
            DPR_CUDA(cudaMemcpyAsync(h_n_active_, d_n_active_ + last_active_slot_, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx().stream));
            DPR_CUDA(cudaStreamSynchronize(ctx().stream));
            if (*h_n_active_ == 0) break;
            // some fits of the batch have converged: give their CTAs' time to the ones still running
            if (n_problems() > 1 && *h_n_active_ < last_active) {
                DPR_TRY(launch_partition(d_problems, n_problems(), grid, total_tiles_, d_cta_start_, ctx().stream));
                last_active = *h_n_active_;
            }
        }
    }
    return check_lists();
}

int Session::check_lists() {
    if (exchanges()) DPR_TRY(peer_exchange_check());   // (a peer-memory exchange that timed out)
    if (!move_overflow_) return DPR_OK;
    int32_t flag = 0;
    DPR_CUDA(cudaMemcpyAsync(&flag, move_overflow_, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    if (flag) return fail(DPR_ERR_STATE, "internal error: a moved-cell list of the delta iterations overflowed");
    return DPR_OK;
}

int Session::fetch_state() {
    DPR_TRY(check_lists());
    DPR_CUDA(cudaMemcpyAsync(h_problems.data(), d_problems, sizeof(Problem) * h_problems.size(), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}

}  // namespace dpr
