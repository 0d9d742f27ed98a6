// penalty_opt.cu — depechePenaltyOpt's state machine and depecheClusterCenterSelection behind the C ABI (SURVEY f-1).
//
// The reference keeps this logic in R (/root/reference/R/depechePenaltyOpt.R:14-265, R/depecheClusterCenterSelection.R:7-59) and
// returns to R between the sets of the search: every set is a foreach over SOCK workers that each receive the data matrix and run
// grid_search (:50-55).  Here the whole adaptive loop runs inside one call on a resident dataset — every set is ONE batch of problems
// on the device — so a one-process R driver (rpkg/) reaches depeche()'s penalty search with a single .Call.
//
// Two layers:
//   dpr_penalty_opt_{create,used,feed,result,solutions,destroy}   the state machine alone: pure host code, no device needed.  A
//       caller that produces the sets itself (the CPU tests with the oracle engine; a grid split over processes) drives it step by step.
//   dpr_dataset_penalty_opt / dpr_dataset_select_solution         the loop around dpr_dataset_grid_search_each, and the selection of
//       the most central solution (allocate the selection set to every stored solution, all-pairs ARI).
#include <algorithm>
#include <cmath>
#include <memory>
#include <vector>

#include "api_internal.cuh"
#include "comm.cuh"
#include "common.cuh"

using namespace dpr;

struct dpr_penalty_opt {
    // arguments of depechePenaltyOpt (:14-17)
    std::vector<double> penalties, real, round1;
    int32_t k = 0, d = 0, n_cores = 1, max_iter = 100;
    uint32_t sample_size = 0;
    double min_ari_improvement = 0.01, optim_ari = 0.95;
    // loop state (:31-45)
    int32_t iter = 1, iter_times_ncores = 1;
    double last_std = 1.0, std_change = 1.0;
    std::vector<int32_t> used;   // usedPositions (0-based)
    // optimListFull: one entry per worker-iteration; ari / nclust / the two fitted centre sets per penalty the entry ran
    struct Entry {
        std::vector<int32_t> pos;
        std::vector<double> ari, nclust, centres;   // centres: [pos][2][k x d col-major]
    };
    std::vector<Entry> full;
    std::vector<double> mean_ari, std_opt;
    bool closed = false;   // result() has been computed
    int32_t best_pos = -1, warn = 0;
    std::vector<double> mean_nclust;
};

namespace {

// R's mean(): long double accumulation, then one refinement pass (src/main/summary.c: rsum / real_mean)
double r_mean(const std::vector<double>& v) {
    if (v.empty()) return NAN;
    long double s = 0.0L;
    for (double x : v) s += x;
    s /= (long double)v.size();
    if (std::isfinite((double)s)) {
        long double t = 0.0L;
        for (double x : v) t += (x - s);
        s += t / (long double)v.size();
    }
    return (double)s;
}
// R's sd(): sqrt(var), two-pass with the mean above (src/library/stats/src/cov.c); NA for fewer than two values
double r_sd(const std::vector<double>& v) {
    if (v.size() < 2) return NAN;
    const long double m = (long double)r_mean(v);
    long double s = 0.0L;
    for (double x : v) s += ((long double)x - m) * ((long double)x - m);
    return std::sqrt((double)(s / (long double)(v.size() - 1)));
}
double nan_max(const std::vector<double>& v) {
    double m = NAN;
    for (double x : v)
        if (!std::isnan(x) && (std::isnan(m) || x > m)) m = x;
    return m;
}
bool loop_continues(const dpr_penalty_opt& o) {   // (:47-48)
    return o.iter_times_ncores < 20 || (o.iter_times_ncores < o.max_iter && o.std_change >= o.min_ari_improvement);
}
// the values of every stored worker-iteration at penalty position p
void values_at(const dpr_penalty_opt& o, int p, bool ari, std::vector<double>& out) {
    out.clear();
    for (const auto& e : o.full)
        for (size_t q = 0; q < e.pos.size(); ++q)
            if (e.pos[q] == p) out.push_back(ari ? e.ari[q] : e.nclust[q]);
}

}  // namespace

extern "C" {

int dpr_penalty_opt_create(const double* penalties, int32_t n_pen, int32_t k, int32_t d, uint32_t sample_size, int32_t n_cores, int32_t max_iter,
                           double min_ari_improvement, double optim_ari, dpr_penalty_opt_t* out) {
    if (!penalties || n_pen < 1 || k < 2 || d < 1 || sample_size < 1 || n_cores < 1 || !out) return fail(DPR_ERR_INVALID, "penalty_opt_create: bad arguments");
    std::unique_ptr<dpr_penalty_opt> o(new dpr_penalty_opt());
    o->penalties.assign(penalties, penalties + n_pen);
    const double constant = ((double)sample_size * std::sqrt((double)d)) / 1450.0;   // penaltyConstant (:22)
    for (int p = 0; p < n_pen; ++p) {
        o->real.push_back(penalties[p] * constant);                                   // realPenalties (:23)
        o->round1.push_back(std::nearbyint(penalties[p] * 10.0) / 10.0);              // round(penalties, digits = 1) (:24)
        o->used.push_back(p);
    }
    o->k = k;
    o->d = d;
    o->sample_size = sample_size;
    o->n_cores = n_cores;
    o->max_iter = max_iter;
    o->min_ari_improvement = min_ari_improvement;
    o->optim_ari = optim_ari;
    o->mean_ari.assign(n_pen, NAN);
    o->std_opt.assign(n_pen, NAN);
    *out = o.release();
    return DPR_OK;
}

int dpr_penalty_opt_destroy(dpr_penalty_opt_t o) {
    delete o;
    return DPR_OK;
}

// the penalties of the next set: positions (0-based) and the scaled values grid_search gets (usedPenalties, :40,143)
int dpr_penalty_opt_used(dpr_penalty_opt_t o, int32_t* n_used_out, int32_t* pos_out, double* reg_out, int32_t* continue_out) {
    if (!o || !n_used_out) return fail(DPR_ERR_INVALID, "penalty_opt_used: null argument");
    *n_used_out = (int32_t)o->used.size();
    for (size_t q = 0; q < o->used.size(); ++q) {
        if (pos_out) pos_out[q] = o->used[q];
        if (reg_out) reg_out[q] = o->real[o->used[q]];
    }
    if (continue_out) *continue_out = loop_continues(*o) ? 1 : 0;
    return DPR_OK;
}

// one set = n_cores worker results over the used penalties (what the foreach of :50-55 returns): ari_each / nclust_each hold
// n_cores rows of n_used values (worker-major: dpr_dataset_grid_search_each's layout for one k), centres (nullable)
// [worker][penalty][2] matrices k x d col-major.  Runs the rest of the loop body (:75-165).
int dpr_penalty_opt_feed(dpr_penalty_opt_t o, const double* ari_each, const double* nclust_each, const double* centres, int32_t* continue_out) {
    if (!o || !ari_each || !nclust_each) return fail(DPR_ERR_INVALID, "penalty_opt_feed: null argument");
    if (o->closed) return fail(DPR_ERR_STATE, "penalty_opt_feed: the search has been closed by dpr_penalty_opt_result");
    const int n_pen = (int)o->penalties.size(), nu = (int)o->used.size();
    const size_t kd = (size_t)o->k * o->d;
    for (int w = 0; w < o->n_cores; ++w) {
        dpr_penalty_opt::Entry e;
        e.pos = o->used;
        e.ari.assign(ari_each + (size_t)w * nu, ari_each + (size_t)(w + 1) * nu);
        e.nclust.assign(nclust_each + (size_t)w * nu, nclust_each + (size_t)(w + 1) * nu);
        for (int q = 0; q < nu; ++q)
            if (e.nclust[q] == 1.0) e.ari[q] = 0.0;   // trivial solutions are practically eliminated (:75-78)
        if (centres) e.centres.assign(centres + (size_t)w * nu * 2 * kd, centres + (size_t)(w + 1) * nu * 2 * kd);
        o->full.push_back(std::move(e));
    }
    std::vector<double> v;
    for (int p = 0; p < n_pen; ++p) {   // mean and sd of the ARI per penalty over everything stored so far (:89-100)
        values_at(*o, p, true, v);
        o->mean_ari[p] = r_mean(v);
        o->std_opt[p] = r_sd(v);
    }
    int max_pos = -1;   // which.max: first maximum, NA skipped (:103)
    for (int p = 0; p < n_pen; ++p)
        if (!std::isnan(o->mean_ari[p]) && (max_pos < 0 || o->mean_ari[p] > o->mean_ari[max_pos])) max_pos = p;
    if (max_pos < 0) return fail(DPR_ERR_INVALID, "penalty_opt_feed: every ARI is NaN");
    const double mean_minus_2std_max = o->mean_ari[max_pos] - 2.0 * o->std_opt[max_pos];   // (:107-108)
    const int local_iter = o->iter < 3 ? o->iter : 3;                                       // (:123-125)
    const double k1 = local_iter == 1 ? 2.0 : local_iter == 2 ? 1.0 : 0.0;
    const double k2 = local_iter == 1 ? 4.0 : local_iter == 2 ? 3.0 : 2.0;
    const double best = nan_max(o->mean_ari);
    std::vector<int32_t> keep;
    for (int32_t p : o->used) {   // (:126-133; comparisons with NA are false)
        const bool overlaps = (o->mean_ari[p] + 2.0 * o->std_opt[p]) >= (mean_minus_2std_max - k1 * o->std_opt[max_pos]);
        const bool similar = o->mean_ari[p] >= best - (1.0 - o->optim_ari) * k2;
        if (overlaps || similar) keep.push_back(p);
    }
    o->used = keep;
    if (o->used.size() > 1) {   // (:145-154)
        if (o->last_std != 0.0) {
            o->std_change = ((o->last_std - o->std_opt[max_pos]) / o->last_std) / std::sqrt((double)o->iter * (double)o->n_cores);
            if (std::isnan(o->std_change)) o->std_change = 0.0;
        } else {
            o->std_change = 0.0;
        }
    } else {
        o->std_change = 0.0;
    }
    o->last_std = o->std_opt[max_pos];
    o->iter_times_ncores = o->iter * o->n_cores;
    o->iter += 1;
    if (continue_out) *continue_out = loop_continues(*o) ? 1 : 0;
    return DPR_OK;
}

// After the loop (:173-264): the best penalty = the median position of the near-optimal set, warnings, the per-penalty means, and
// how many solutions were stored at the best penalty.  warn_flags: bit 0 "maximum number of iterations reached" (:173-177),
// bit 1 / bit 2 the best penalty is the lowest / highest of the range (:196-207).
int dpr_penalty_opt_result(dpr_penalty_opt_t o, double* best_penalty_out, int32_t* best_pos_out, double* mean_ari_out, double* mean_nclust_out,
                           int32_t* worker_iterations_out, int32_t* n_solutions_out, int32_t* warn_flags_out) {
    if (!o) return fail(DPR_ERR_INVALID, "penalty_opt_result: null handle");
    if (o->full.empty()) return fail(DPR_ERR_STATE, "penalty_opt_result: no set has been fed");
    const int n_pen = (int)o->penalties.size();
    if (!o->closed) {
        o->warn = 0;
        if (o->iter * o->n_cores >= o->max_iter && o->std_change > o->min_ari_improvement) o->warn |= 1;
        const double best = nan_max(o->mean_ari);
        std::vector<int32_t> optimal;   // (:184-185)
        for (int p = 0; p < n_pen; ++p)
            if (o->mean_ari[p] >= best - (1.0 - o->optim_ari)) optimal.push_back(p);
        if (optimal.empty()) return fail(DPR_ERR_INVALID, "penalty_opt_result: no finite ARI");
        // round(mean(seq_along(optimal))): R rounds half to even (:192)
        const double centre = ((double)optimal.size() + 1.0) / 2.0;
        o->best_pos = optimal[(size_t)std::nearbyint(centre) - 1];
        if (n_pen > 1) {
            if (o->round1[o->best_pos] == o->round1[0]) o->warn |= 2;
            if (o->round1[o->best_pos] == o->round1[n_pen - 1]) o->warn |= 4;
        }
        o->mean_nclust.assign(n_pen, NAN);
        std::vector<double> v, w;
        for (int p = 0; p < n_pen; ++p) {   // mean(..., na.rm = TRUE) (:242-249)
            values_at(*o, p, false, v);
            w.clear();
            for (double x : v)
                if (!std::isnan(x)) w.push_back(x);
            o->mean_nclust[p] = r_mean(w);
        }
        o->closed = true;
    }
    if (best_penalty_out) *best_penalty_out = o->round1[o->best_pos];
    if (best_pos_out) *best_pos_out = o->best_pos;
    for (int p = 0; p < n_pen; ++p) {
        if (mean_ari_out) mean_ari_out[p] = o->mean_ari[p];
        if (mean_nclust_out) mean_nclust_out[p] = o->mean_nclust[p];
    }
    if (worker_iterations_out) *worker_iterations_out = (o->iter - 1) * o->n_cores;
    if (n_solutions_out) {
        int m = 0;
        for (const auto& e : o->full)
            if (!e.centres.empty() && std::find(e.pos.begin(), e.pos.end(), o->best_pos) != e.pos.end()) m += 2;
        *n_solutions_out = m;
    }
    if (warn_flags_out) *warn_flags_out = o->warn;
    return DPR_OK;
}

// bestClusterCenters (:255-258): every centre matrix fitted at the best penalty, in worker-iteration order, ret1 then ret2;
// n_solutions matrices k x d col-major back to back
int dpr_penalty_opt_solutions(dpr_penalty_opt_t o, double* centres_out) {
    if (!o || !centres_out) return fail(DPR_ERR_INVALID, "penalty_opt_solutions: null argument");
    if (!o->closed) return fail(DPR_ERR_STATE, "penalty_opt_solutions: call dpr_penalty_opt_result first");
    const size_t kd = (size_t)o->k * o->d;
    size_t off = 0;
    for (const auto& e : o->full) {
        if (e.centres.empty()) continue;
        for (size_t q = 0; q < e.pos.size(); ++q)
            if (e.pos[q] == o->best_pos) {
                std::copy(e.centres.begin() + q * 2 * kd, e.centres.begin() + (q + 1) * 2 * kd, centres_out + off);
                off += 2 * kd;
            }
    }
    return DPR_OK;
}

// The whole search on a resident dataset.  Set `it` draws from seed + 7919 * it, so repeated sets differ while a given seed
// reproduces the search.  split_over_ranks != 0 (after dpr_comm_init, every rank holding the SAME matrix): the (worker, penalty)
// problems of each set are dealt over the ranks in contiguous blocks — they are independent (src/Clusterer.cpp:349-384), random
// streams are addressed by the problem's global number, so the results do not depend on the number of ranks — and the small
// per-problem results are summed into place with one allreduce per set; every rank then runs the identical state machine.
int dpr_dataset_penalty_opt(dpr_dataset_t ds, dpr_penalty_opt_t o, uint64_t seed, int32_t split_over_ranks) {
    if (!ds || !o) return fail(DPR_ERR_INVALID, "penalty_opt: null argument");
    if (ds->d != o->d) return fail(DPR_ERR_INVALID, "penalty_opt: the dataset's marker count differs from the one given to dpr_penalty_opt_create");
    DPR_TRY(ensure_ready());
    const int world = split_over_ranks ? ctx().world : 1, rank = split_over_ranks ? ctx().rank : 0;
    const size_t kd = (size_t)o->k * o->d;
    std::vector<double> ari, ncl, cen, reg;
    while (loop_continues(*o)) {
        const int nu = (int)o->used.size(), cells = o->n_cores * nu;
        reg.resize(nu);
        for (int q = 0; q < nu; ++q) reg[q] = o->real[o->used[q]];
        ari.assign((size_t)cells, 0.0);
        ncl.assign((size_t)cells, 0.0);
        cen.assign((size_t)cells * 2 * kd, 0.0);
        const int c_lo = (int)((int64_t)cells * rank / world), c_hi = (int)((int64_t)cells * (rank + 1) / world);
        const uint64_t set_seed = seed + 7919ull * (uint64_t)o->iter;
        if (c_hi > c_lo)
            DPR_TRY(grid_search_cells(ds, o->k, reg.data(), nu, (uint32_t)o->n_cores, o->sample_size, set_seed, c_lo, c_hi, ari.data() + c_lo, ncl.data() + c_lo,
                                      cen.data() + (size_t)c_lo * 2 * kd));
        if (world > 1) DPR_TRY(comm_allreduce_host_f64({&ari, &ncl, &cen}));
        DPR_TRY(dpr_penalty_opt_feed(o, ari.data(), ncl.data(), cen.data(), nullptr));
    }
    return DPR_OK;
}

// depecheClusterCenterSelection (R/depecheClusterCenterSelection.R:11-37): allocate the selection set to each of the m stored
// solutions (dAllocate drops their zero-sum rows, R/dAllocate.R:99-103: masked here), all-pairs rand_index, the solution with the
// highest mean ARI against the others wins (first maximum).  solutions: m matrices k x d col-major back to back.
int dpr_dataset_select_solution(dpr_dataset_t sel, const double* solutions, int32_t m, int32_t k, int32_t* best_out, double* mean_ari_out) {
    if (!sel || !solutions || m < 1 || k < 1 || !best_out) return fail(DPR_ERR_INVALID, "select_solution: bad arguments");
    const int d = sel->d;
    const size_t kd = (size_t)k * d;
    std::vector<double> masked(solutions, solutions + (size_t)m * kd);
    for (int s = 0; s < m; ++s)
        for (int r = 0; r < k; ++r) {
            double row_sum = 0.0;   // rowSums(centers) != 0
            for (int j = 0; j < d; ++j) row_sum += masked[(size_t)s * kd + (size_t)j * k + r];
            if (row_sum == 0.0)
                for (int j = 0; j < d; ++j) masked[(size_t)s * kd + (size_t)j * k + r] = 0.0;
        }
    // dAllocate also drops the columns with colSums(centers) == 0 (:99-103).  A resident matrix keeps its columns, so such a column
    // of the centres is zeroed instead (it matters only when it sums to zero by cancellation): it then adds the same x_j^2 to every
    // distance of a cell
    for (int s = 0; s < m; ++s)
        for (int j = 0; j < d; ++j) {
            double col_sum = 0.0;
            for (int r = 0; r < k; ++r) col_sum += solutions[(size_t)s * kd + (size_t)j * k + r];
            if (col_sum == 0.0)
                for (int r = 0; r < k; ++r) masked[(size_t)s * kd + (size_t)j * k + r] = 0.0;
        }
    std::vector<double> ari((size_t)m * m);
    DPR_TRY(dpr_dataset_allocate_many(sel, masked.data(), m, k, 1, nullptr, ari.data()));
    int best = -1;
    double best_v = NAN;
    std::vector<double> col(m);
    for (int i = 0; i < m; ++i) {   // mean over j of rand_index(alloc[j], alloc[i]) (:24-29)
        for (int j = 0; j < m; ++j) col[j] = ari[(size_t)i * m + j];
        const double v = r_mean(col);
        if (mean_ari_out) mean_ari_out[i] = v;
        if (!std::isnan(v) && (best < 0 || v > best_v)) {   // which(meanARI == max(meanARI))[1] (:35-37)
            best = i;
            best_v = v;
        }
    }
    if (best < 0) return fail(DPR_ERR_INVALID, "select_solution: every mean ARI is NaN");
    *best_out = best;
    return DPR_OK;
}

}  // extern "C"
