/*
 * clusterer_oracle.cpp — CPU parity oracle (TEST INFRASTRUCTURE, see dpr_oracle.h).
 *
 * A restatement, in plain C++ without Eigen or Rcpp, of the numerical engine in
 * /root/reference/src/Clusterer.cpp.  Every function names the reference lines it follows.
 * The operation ORDER is the reference's (sequential double sums, sqrt before the argmin,
 * strict-< first minimum, float ramp factor), and so is its COST STRUCTURE (row-major double
 * data, one sweep over X per centre, a materialised n x k distance matrix): this file is also
 * what bench.py times as the CPU baseline.
 *
 * Randomness is injected (struct Rng): either libc rand() in the reference's call order, or the
 * Philox streams of include/dpr_philox.h that the CUDA library uses.
 */
#include "dpr_oracle.h"
#include "../include/dpr_philox.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

namespace {

typedef std::vector<double> dvec;
typedef std::vector<int32_t> ivec;

// ---------------------------------------------------------------------------------------------
// randomness
// ---------------------------------------------------------------------------------------------
struct Rng {
    int mode = ORACLE_RNG_PHILOX;
    uint64_t seed = 0;
    // row number for bootstrap draw i of a fit (Clusterer.cpp:314-316: Random(brows) then % xrows)
    uint64_t boot_row(uint32_t fit, uint32_t i, uint64_t n) const {
        if (mode == ORACLE_RNG_LIBC) return (uint64_t)((unsigned int)std::rand() % (unsigned int)n);
        return dpr_rng_draw(seed, DPR_RNG_BOOT, fit, i, 0).w[0] % n;
    }
    // first centre (Clusterer.cpp:156: std::rand() % X_rows)
    uint64_t first_row(uint32_t fit, uint64_t n) const {
        if (mode == ORACLE_RNG_LIBC) return (uint64_t)(std::rand() % (int)n);
        return dpr_rng_draw(seed, DPR_RNG_SEED, fit, 0, 0).w[0] % n;
    }
    // u for the t-th weighted pick (Clusterer.cpp:136: (double) rand() / (RAND_MAX))
    double unit(uint32_t fit, uint32_t t) const {
        if (mode == ORACLE_RNG_LIBC) return (double)std::rand() / (RAND_MAX);
        return dpr_rng_unit(dpr_rng_draw(seed, DPR_RNG_SEED, fit, t, 0).w[0]);
    }
    // one candidate pair (Clusterer.cpp:292-294: Matrix<unsigned,2,1>::Random(), each % intSize)
    void pair(uint32_t pair_id, uint32_t num, uint32_t attempt, uint64_t n, uint64_t* i, uint64_t* j) const {
        if (mode == ORACLE_RNG_LIBC) {
            unsigned int r0 = (unsigned int)std::rand();
            unsigned int r1 = (unsigned int)std::rand();
            *i = r0 % (unsigned int)n;
            *j = r1 % (unsigned int)n;
            return;
        }
        dpr_u32x4 w = dpr_rng_draw(seed, DPR_RNG_PAIR, pair_id, num, attempt);
        *i = w.w[0] % n;
        *j = w.w[1] % n;
    }
};

Rng g_rng;
int g_dead_work = 1;

// ---------------------------------------------------------------------------------------------
// layout helpers: R column-major <-> the reference's row-major working copy
// (numeric_to_eigen / eigen_to_numeric, src/InterfaceUtils.cpp:23-33, 53-63)
// ---------------------------------------------------------------------------------------------
dvec to_row_major(const double* cm, size_t rows, size_t cols) {
    dvec out(rows * cols);
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) out[i * cols + j] = cm[j * rows + i];
    return out;
}
void to_col_major(const dvec& rm, size_t rows, size_t cols, double* cm) {
    for (size_t i = 0; i < rows; ++i)
        for (size_t j = 0; j < cols; ++j) cm[j * rows + i] = rm[i * cols + j];
}

// ---------------------------------------------------------------------------------------------
// allocate_clusters — src/Clusterer.cpp:34-87
// ---------------------------------------------------------------------------------------------
// Active centres: all, or (no_zero) those with any coefficient != 0.0 (isZero(0), :41-45; a NaN
// coefficient is "not zero").  Fewer than two active => every label 0 (:50-53).
std::vector<unsigned> active_centres(const dvec& mu, unsigned k, unsigned d, bool no_zero) {
    std::vector<unsigned> act;
    for (unsigned c = 0; c < k; ++c) {
        bool use = true;
        if (no_zero) {
            use = false;
            for (unsigned t = 0; t < d; ++t) {
                double v = mu[(size_t)c * d + t];
                if (!(std::fabs(v) <= 0.0)) { use = true; break; }
            }
        }
        if (use) act.push_back(c);
    }
    return act;
}

// One column of the distance matrix per active centre, filled by a full sweep over X
// (:57-59; Euclidean norm = sqrt of the left-to-right sum of squares of (-x)+mu), then a
// row-wise first minimum (:63-66, minCoeff keeps the earliest smallest value).
void allocate_clusters(const dvec& X, size_t n, unsigned d, const dvec& mu, unsigned k, bool no_zero,
                       ivec& out, dvec* margin = nullptr) {
    out.assign(n, 0);
    std::vector<unsigned> act = active_centres(mu, k, d, no_zero);
    if (margin) margin->assign(n, INFINITY);
    if (act.size() < 2) return;
    const size_t na = act.size();
    dvec dist(n * na, 0.0);
    for (size_t a = 0; a < na; ++a) {
        const double* m = &mu[(size_t)act[a] * d];
        for (size_t i = 0; i < n; ++i) {
            const double* x = &X[i * d];
            double s = 0.0;
            for (unsigned t = 0; t < d; ++t) {
                double diff = (-x[t]) + m[t];
                s += diff * diff;
            }
            dist[i * na + a] = std::sqrt(s);
        }
    }
    for (size_t i = 0; i < n; ++i) {
        const double* row = &dist[i * na];
        size_t best = 0;
        double bv = row[0];
        for (size_t a = 1; a < na; ++a)
            if (row[a] < bv) { bv = row[a]; best = a; }
        out[i] = (int32_t)act[best];
        if (margin) {
            double second = INFINITY;
            for (size_t a = 0; a < na; ++a)
                if (a != best && row[a] < second) second = row[a];
            (*margin)[i] = bv > 0 ? (second - bv) / bv : (second > 0 ? INFINITY : 0.0);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// reevaluate_centers — src/Clusterer.cpp:89-108
// ---------------------------------------------------------------------------------------------
// Sums in row order, int counts, then mu = min((S+reg/2)/n, max((S-reg/2)/n, 0)).
// min/max follow the SSE2 semantics Eigen's packet path has on x86-64 (_mm_max_pd/_mm_min_pd
// return the SECOND operand when either is NaN), so an empty cluster (0/0 when reg == 0,
// -+inf when reg > 0) becomes an exact 0 row — a dead centre (SURVEY A-6, A-7).
inline double sse_max(double a, double b) { return a > b ? a : b; }
inline double sse_min(double a, double b) { return a < b ? a : b; }

void reevaluate_centers(const dvec& X, size_t n, unsigned d, const ivec& idx, unsigned k, double reg,
                        dvec& mu, std::vector<int64_t>* counts_out = nullptr) {
    dvec sums((size_t)k * d, 0.0);
    std::vector<int> nums(k, 0);
    for (size_t i = 0; i < n; ++i) {
        double* s = &sums[(size_t)idx[i] * d];
        const double* x = &X[i * d];
        for (unsigned t = 0; t < d; ++t) s[t] += x[t];
        nums[idx[i]] += 1;
    }
    mu.assign((size_t)k * d, 0.0);
    for (unsigned c = 0; c < k; ++c)
        for (unsigned t = 0; t < d; ++t) {
            double S = sums[(size_t)c * d + t];
            double lo = (S - reg / 2) / nums[c];
            double hi = (S + reg / 2) / nums[c];
            mu[(size_t)c * d + t] = sse_min(hi, sse_max(lo, 0.0));
        }
    if (counts_out) {
        counts_out->assign(k, 0);
        for (unsigned c = 0; c < k; ++c) (*counts_out)[c] = nums[c];
    }
}

// ---------------------------------------------------------------------------------------------
// element_from_vector + initialize_mu — src/Clusterer.cpp:123-176 (k-means++ D^2 sampling)
// ---------------------------------------------------------------------------------------------
// Cumulative sums of the strictly positive weights keyed in an ordered map; x = u * total;
// x < total -> first key strictly greater than x (upper_bound); else the entry at key x (:138-143;
// operator[] would insert 0 if absent, which is what the reference does too).
unsigned element_from_vector(const dvec& w, double u) {
    std::map<double, unsigned> cum_to_row;
    double cum = 0;
    for (size_t i = 0; i < w.size(); ++i)
        if (w[i] > 0) {
            cum += w[i];
            cum_to_row[cum] = (unsigned)i;
        }
    double x = u * cum;
    if (x < cum) return cum_to_row.upper_bound(x)->second;
    return cum_to_row[x];
}

void sq_dist_to_row(const dvec& X, size_t n, unsigned d, const double* m, dvec& out) {
    out.resize(n);
    for (size_t i = 0; i < n; ++i) {
        const double* x = &X[i * d];
        double s = 0.0;
        for (unsigned t = 0; t < d; ++t) {
            double diff = (-x[t]) + m[t];
            s += diff * diff;
        }
        out[i] = s;
    }
}

void initialize_mu(const dvec& X, size_t n, unsigned d, unsigned k, uint32_t fit, dvec& mu,
                   std::vector<int64_t>* rows_out = nullptr) {
    mu.assign((size_t)k * d, 0.0);
    if (rows_out) rows_out->assign(k, -1);
    size_t r0 = (size_t)g_rng.first_row(fit, n);
    std::copy(&X[r0 * d], &X[r0 * d] + d, &mu[0]);
    if (rows_out) (*rows_out)[0] = (int64_t)r0;
    dvec dists;
    sq_dist_to_row(X, n, d, &mu[0], dists);
    unsigned pick = element_from_vector(dists, g_rng.unit(fit, 1));
    std::copy(&X[(size_t)pick * d], &X[(size_t)pick * d] + d, &mu[d]);
    if (rows_out) (*rows_out)[1] = pick;
    dvec tmp;
    for (unsigned c = 2; c < k; ++c) {
        sq_dist_to_row(X, n, d, &mu[(size_t)(c - 1) * d], tmp);
        for (size_t i = 0; i < n; ++i)
            if (tmp[i] < dists[i]) dists[i] = tmp[i];
        pick = element_from_vector(dists, g_rng.unit(fit, c));
        std::copy(&X[(size_t)pick * d], &X[(size_t)pick * d] + d, &mu[(size_t)c * d]);
        if (rows_out) (*rows_out)[c] = pick;
    }
}

// ---------------------------------------------------------------------------------------------
// cluster_norm — src/Clusterer.cpp:198-211: SSQ + reg * (signed) sum of the centre coefficients,
// centres traversed column by column (:204-208).
// ---------------------------------------------------------------------------------------------
double cluster_norm(const dvec& X, size_t n, unsigned d, const dvec& mu, unsigned k, const ivec& idx,
                    double reg) {
    double ssq = 0;
    for (size_t i = 0; i < n; ++i) {
        const double* x = &X[i * d];
        const double* m = &mu[(size_t)idx[i] * d];
        double s = 0.0;
        for (unsigned t = 0; t < d; ++t) {
            double diff = x[t] - m[t];
            s += diff * diff;
        }
        ssq += s;
    }
    for (unsigned t = 0; t < d; ++t)
        for (unsigned c = 0; c < k; ++c) ssq += mu[(size_t)c * d + t] * reg;
    return ssq;
}

// ---------------------------------------------------------------------------------------------
// find_centers — src/Clusterer.cpp:213-265
// ---------------------------------------------------------------------------------------------
struct Fit {
    ivec idx;
    dvec mu;
    double norm = 0;
    int n_assign = 0;
};

// The Lloyd loop (:244-252): assign; stop if labels unchanged AND i > 20; update with the penalty
// ramped as reg * ((float)i / (float)20), capped at reg.  Old labels start as zeros (:226).
void lloyd(const dvec& X, size_t n, unsigned d, unsigned k, double reg, bool no_zero, int max_iter, Fit& f) {
    ivec old(n, 0);
    f.idx.assign(n, 1);
    f.n_assign = 0;
    for (int i = 0; i < max_iter; ++i) {
        allocate_clusters(X, n, d, f.mu, k, no_zero, f.idx);
        f.n_assign = i + 1;
        if (i > 20 && f.idx == old) break;
        old = f.idx;
        double ramp = reg * ((float)i / (float)20);
        reevaluate_centers(X, n, d, f.idx, k, std::min(reg, ramp), f.mu);
    }
    f.norm = cluster_norm(X, n, d, f.mu, k, f.idx, reg);
}

void find_centers(const dvec& X, size_t n, unsigned d, unsigned k, double reg, bool no_zero, uint32_t fit,
                  Fit& f) {
    initialize_mu(X, n, d, k, fit, f.mu);
    lloyd(X, n, d, k, reg, no_zero, 1000, f);
}

// ---------------------------------------------------------------------------------------------
// cluster_distance — src/Clusterer.cpp:267-305 (Monte-Carlo adjusted Rand index)
// ---------------------------------------------------------------------------------------------
// Populations accumulate 1.0/size per member (:275-278); chance terms false1/false2 (:281-282);
// Rand index from 10 000 random pairs i != j (:286-302); (RI - E)/(1 - E) (:304).
// Tables are sized k+1 so that R's 1-based labels with k = max label stay in range (SURVEY A-9;
// the reference indexes out of bounds there).
double cluster_distance(const int32_t* c1, const int32_t* c2, size_t n, unsigned k, uint32_t pair_id) {
    const double size = (double)(unsigned int)n;
    dvec p1(k + 1, 0.0), p2(k + 1, 0.0);
    for (size_t i = 0; i < n; ++i) {
        p1[c1[i]] += 1.0 / size;
        p2[c2[i]] += 1.0 / size;
    }
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0;
    for (unsigned c = 0; c <= k; ++c) {  // entry k is 0 (and adds 0) whenever labels stay below k
        a1 += p1[c] * (p1[c] - 1.0 / size) * (size / (size - 1));
        a2 += p2[c] * (p2[c] - 1.0 / size) * (size / (size - 1));
        b1 += p1[c] * (1 - (p1[c] - 1.0 / size) * (size / (size - 1)));
        b2 += p2[c] * (1 - (p2[c] - 1.0 / size) * (size / (size - 1)));
    }
    const double false1 = a1 * a2;
    const double false2 = b1 * b2;
    double stability = 0;
    const unsigned pairs = 10000;
    for (unsigned num = 0; num < pairs; ++num) {
        uint64_t i = 0, j = 0;
        uint32_t attempt = 0;
        while (i == j) g_rng.pair(pair_id, num, attempt++, n, &i, &j);
        if (c1[i] == c1[j] && c2[i] == c2[j]) stability += 1;
        else if (c1[i] != c1[j] && c2[i] != c2[j]) stability += 1;
    }
    return ((stability / pairs) - (false1 + false2)) / (1 - (false1 + false2));
}

// Expectation of the estimator above over the pair draw: every ordered pair i != j equally likely.
double cluster_distance_exact(const int32_t* c1, const int32_t* c2, size_t n, unsigned k) {
    const unsigned kk = k + 1;
    std::vector<int64_t> tab((size_t)kk * kk, 0), r(kk, 0), c(kk, 0);
    for (size_t i = 0; i < n; ++i) {
        tab[(size_t)c1[i] * kk + c2[i]]++;
        r[c1[i]]++;
        c[c2[i]]++;
    }
    const double N = (double)n, pairs = N * (N - 1);
    double same12 = 0, s1 = 0, s2 = 0;
    for (size_t t = 0; t < tab.size(); ++t) same12 += (double)tab[t] * (double)(tab[t] - 1);
    for (unsigned t = 0; t < kk; ++t) {
        s1 += (double)r[t] * (double)(r[t] - 1);
        s2 += (double)c[t] * (double)(c[t] - 1);
    }
    const double A = s1 / pairs, B = s2 / pairs, both = same12 / pairs;
    const double ri = both + (1 - A - B + both);
    const double e = A * B + (1 - A) * (1 - B);
    return (ri - e) / (1 - e);
}

// bootstrap_data — src/Clusterer.cpp:308-321
void bootstrap_rows(size_t n, unsigned boot_n, uint32_t fit, std::vector<int64_t>& rows) {
    rows.resize(boot_n);
    for (unsigned i = 0; i < boot_n; ++i) rows[i] = (int64_t)g_rng.boot_row(fit, i, n);
}
void gather_rows(const dvec& X, unsigned d, const std::vector<int64_t>& rows, dvec& out) {
    out.resize(rows.size() * d);
    for (size_t i = 0; i < rows.size(); ++i)
        std::copy(&X[(size_t)rows[i] * d], &X[(size_t)rows[i] * d] + d, &out[i * d]);
}

// n_used_clusters — src/Clusterer.cpp:323-332
unsigned n_used_clusters(unsigned k, const int32_t* idx, size_t n) {
    std::vector<int> seen(k + 1, 0);
    for (size_t i = 0; i < n; ++i) seen[idx[i]] = 1;
    unsigned s = 0;
    for (unsigned c = 0; c <= k; ++c) s += seen[c];
    return s;
}

}  // namespace

// =================================================================================================
// C interface
// =================================================================================================
extern "C" {

void oracle_set_rng(int mode, uint64_t seed) {
    g_rng.mode = mode;
    g_rng.seed = seed;
}
void oracle_srand(unsigned int seed32) { std::srand(seed32); }
void oracle_set_dead_work(int on) { g_dead_work = on; }

int oracle_allocate_points(const double* X, int64_t n, int32_t d, const double* mu, int32_t k,
                           int32_t no_zero, int32_t* idx_out) {
    if (!X || !mu || !idx_out || n < 1 || d < 1 || k < 1) return -1;
    dvec Xr = to_row_major(X, n, d), M = to_row_major(mu, k, d);
    ivec idx;
    allocate_clusters(Xr, n, d, M, k, no_zero != 0, idx);
    std::memcpy(idx_out, idx.data(), sizeof(int32_t) * n);
    return 0;
}

int oracle_assignment_margin(const double* X, int64_t n, int32_t d, const double* mu, int32_t k,
                             int32_t no_zero, double* margin_out) {
    dvec Xr = to_row_major(X, n, d), M = to_row_major(mu, k, d), mg;
    ivec idx;
    allocate_clusters(Xr, n, d, M, k, no_zero != 0, idx, &mg);
    std::memcpy(margin_out, mg.data(), sizeof(double) * n);
    return 0;
}

// sparse_k_means — src/InterfaceUtils.cpp:66-94
int oracle_sparse_k_means(const double* X, int64_t n, int32_t d, uint32_t k, double reg, int32_t no_zero,
                          uint32_t fit_id, int32_t* idx_out, double* centers_out, double* norm_out,
                          int32_t* n_iter_out) {
    if (!X || n < 1 || d < 1 || k < 2) return -1;
    dvec Xr = to_row_major(X, n, d);
    Fit f;
    find_centers(Xr, n, d, k, reg, no_zero != 0, fit_id, f);
    if (idx_out) std::memcpy(idx_out, f.idx.data(), sizeof(int32_t) * n);
    if (centers_out) to_col_major(f.mu, k, d, centers_out);
    if (norm_out) *norm_out = f.norm;
    if (n_iter_out) *n_iter_out = f.n_assign;
    return 0;
}

int oracle_find_centers_from(const double* X, int64_t n, int32_t d, const double* init_mu, uint32_t k,
                             double reg, int32_t no_zero, int32_t max_iter, int32_t* idx_out,
                             double* centers_out, double* norm_out, int32_t* n_iter_out) {
    if (!X || !init_mu || n < 1 || d < 1 || k < 2) return -1;
    dvec Xr = to_row_major(X, n, d);
    Fit f;
    f.mu = to_row_major(init_mu, k, d);
    lloyd(Xr, n, d, k, reg, no_zero != 0, max_iter, f);
    if (idx_out) std::memcpy(idx_out, f.idx.data(), sizeof(int32_t) * n);
    if (centers_out) to_col_major(f.mu, k, d, centers_out);
    if (norm_out) *norm_out = f.norm;
    if (n_iter_out) *n_iter_out = f.n_assign;
    return 0;
}

// optimize_param — src/Clusterer.cpp:334-394, behind grid_search (src/InterfaceUtils.cpp:99-131).
// Per (iteration, k, reg): two bootstraps, two fits with zero-centre removal (:354-362), the full
// data allocated to both centre sets WITHOUT zero removal (:370-371, no_zero defaults to false,
// Clusterer.h:49), two more allocations whose results are discarded (:373-374), ARI of the pair
// (:376) and the distinct labels of the two bootstrap fits (:379-380).
int oracle_grid_search(const double* X, int64_t n, int32_t d, const int32_t* k, int32_t nk,
                       const double* reg, int32_t nreg, uint32_t iterations, uint32_t boot_n,
                       double* ari_out, double* nclust_out, double* centers_out) {
    if (!X || !k || !reg || n < 1 || d < 1 || nk < 1 || nreg < 1 || iterations < 1) return -1;
    dvec Xr = to_row_major(X, n, d);
    dvec dist((size_t)nk * nreg, 0.0), found((size_t)nk * nreg, 0.0);
    size_t coff = 0;
    for (uint32_t it = 0; it < iterations; ++it)
        for (int32_t j = 0; j < nk; ++j)
            for (int32_t l = 0; l < nreg; ++l) {
                const uint32_t problem = (it * (uint32_t)nk + (uint32_t)j) * (uint32_t)nreg + (uint32_t)l;
                const unsigned kk = (unsigned)k[j];
                std::vector<int64_t> rows1, rows2;
                dvec b1, b2;
                bootstrap_rows(n, boot_n, 2 * problem, rows1);
                gather_rows(Xr, d, rows1, b1);
                bootstrap_rows(n, boot_n, 2 * problem + 1, rows2);
                gather_rows(Xr, d, rows2, b2);
                Fit f1, f2;
                find_centers(b1, boot_n, d, kk, reg[l], true, 2 * problem, f1);
                find_centers(b2, boot_n, d, kk, reg[l], true, 2 * problem + 1, f2);
                if (centers_out) {
                    to_col_major(f1.mu, kk, d, centers_out + coff);
                    coff += (size_t)kk * d;
                    to_col_major(f2.mu, kk, d, centers_out + coff);
                    coff += (size_t)kk * d;
                }
                ivec ind1, ind2;
                allocate_clusters(Xr, n, d, f1.mu, kk, false, ind1);
                allocate_clusters(Xr, n, d, f2.mu, kk, false, ind2);
                if (g_dead_work) {
                    ivec z1, z2;
                    allocate_clusters(Xr, n, d, f1.mu, kk, true, z1);
                    allocate_clusters(Xr, n, d, f2.mu, kk, true, z2);
                }
                dist[(size_t)j * nreg + l] += cluster_distance(ind1.data(), ind2.data(), n, kk, problem);
                found[(size_t)j * nreg + l] += n_used_clusters(kk, f1.idx.data(), boot_n);
                found[(size_t)j * nreg + l] += n_used_clusters(kk, f2.idx.data(), boot_n);
            }
    for (int32_t j = 0; j < nk; ++j)
        for (int32_t l = 0; l < nreg; ++l) {
            ari_out[(size_t)l * nk + j] = dist[(size_t)j * nreg + l] / iterations;
            nclust_out[(size_t)l * nk + j] = found[(size_t)j * nreg + l] / (iterations * 2);
        }
    return 0;
}

int oracle_rand_index(const int32_t* a, const int32_t* b, int64_t n, uint32_t k, uint32_t pair_id, double* out) {
    if (!a || !b || !out || n < 2) return -1;
    for (int64_t i = 0; i < n; ++i)
        if (a[i] < 0 || b[i] < 0 || (uint32_t)a[i] > k || (uint32_t)b[i] > k) return -1;
    *out = cluster_distance(a, b, (size_t)n, k, pair_id);
    return 0;
}
int oracle_rand_index_exact(const int32_t* a, const int32_t* b, int64_t n, uint32_t k, double* out) {
    if (!a || !b || !out || n < 2) return -1;
    for (int64_t i = 0; i < n; ++i)
        if (a[i] < 0 || b[i] < 0 || (uint32_t)a[i] > k || (uint32_t)b[i] > k) return -1;
    *out = cluster_distance_exact(a, b, (size_t)n, k);
    return 0;
}

int oracle_reevaluate_centers(const double* X, int64_t n, int32_t d, const int32_t* idx, uint32_t k,
                              double reg, double* centers_out, int64_t* counts_out) {
    dvec Xr = to_row_major(X, n, d), mu;
    ivec id(idx, idx + n);
    std::vector<int64_t> cnt;
    reevaluate_centers(Xr, n, d, id, k, reg, mu, &cnt);
    to_col_major(mu, k, d, centers_out);
    if (counts_out) std::copy(cnt.begin(), cnt.end(), counts_out);
    return 0;
}

int oracle_initialize_mu(const double* X, int64_t n, int32_t d, uint32_t k, uint32_t fit_id,
                         double* mu_out, int64_t* rows_out) {
    if (k < 2) return -1;
    dvec Xr = to_row_major(X, n, d), mu;
    std::vector<int64_t> rows;
    initialize_mu(Xr, n, d, k, fit_id, mu, &rows);
    to_col_major(mu, k, d, mu_out);
    if (rows_out) std::copy(rows.begin(), rows.end(), rows_out);
    return 0;
}

int oracle_bootstrap_rows(int64_t n, uint32_t boot_n, uint32_t fit_id, int64_t* rows_out) {
    std::vector<int64_t> rows;
    bootstrap_rows((size_t)n, boot_n, fit_id, rows);
    std::copy(rows.begin(), rows.end(), rows_out);
    return 0;
}

int oracle_cluster_norm(const double* X, int64_t n, int32_t d, const double* mu, uint32_t k,
                        const int32_t* idx, double reg, double* out) {
    dvec Xr = to_row_major(X, n, d), M = to_row_major(mu, k, d);
    ivec id(idx, idx + n);
    *out = cluster_norm(Xr, n, d, M, k, id, reg);
    return 0;
}

int oracle_n_used_clusters(const int32_t* idx, int64_t n, uint32_t k, uint32_t* out) {
    *out = n_used_clusters(k, idx, (size_t)n);
    return 0;
}

int oracle_time_lloyd(const double* X, int64_t n, int32_t d, const double* mu, uint32_t k, double reg,
                      int32_t no_zero, int32_t iters, double* seconds_out) {
    dvec Xr = to_row_major(X, n, d), M = to_row_major(mu, k, d);
    ivec idx;
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) {
        allocate_clusters(Xr, n, d, M, k, no_zero != 0, idx);
        reevaluate_centers(Xr, n, d, idx, k, reg, M);
    }
    auto t1 = std::chrono::steady_clock::now();
    *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    return 0;
}

}  // extern "C"
