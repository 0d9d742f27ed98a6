// InterfaceUtils.cpp — Rcpp host layer of the B200 build: drop-in replacement for the reference's
// src/InterfaceUtils.cpp.  The four exported functions keep their names, argument lists and the keys of the
// lists they return (/root/reference/src/InterfaceUtils.cpp:66-161), so every R caller of them stays what it is.
// The bodies no longer build Eigen matrices and a Clusterer: they hand R's own column-major buffers to the C ABI
// of libdepecher_b200 (include/depecher_b200.h).
//
// New exports (the one-process driver in rpkg/R/ uses them; after adding this file run Rcpp::compileAttributes()
// to regenerate src/RcppExports.cpp — rpkg/R/RcppExports.R shows the R side it produces):
//   dataset_create / dataset_destroy / dataset_gather / dataset_allocate_points   the matrix stays in HBM across calls
//                                                         (external pointer; the finalizer frees the device memory)
//   dataset_create_scaled            depecheLogCenterSd on the device while uploading (the scaled matrix never exists in R)
//   dataset_allocate_points_ranked   depeche()'s final dAllocate + renumbering by cluster size in one call
//   grid_search_each     one set of the penalty search = all former SOCK workers' grid_search calls as one batch
//   penalty_opt          depechePenaltyOpt's whole adaptive loop in one call (dpr_dataset_penalty_opt)
//   select_solution      depecheClusterCenterSelection's allocation + all-pairs ARI (dpr_dataset_select_solution)
//   set_engine_seed      reproducible runs (see next_seed)
//
// Differences a user can observe:
//   * rand_index returns the contingency-table value the reference's 10 000-pair estimate scatters around.
//   * errors (no B200, bad arguments) surface as R errors through Rcpp::stop.
#include <Rcpp.h>

#include <atomic>
#include <ctime>
#include <cstdlib>
#include <string>
#include <vector>
#include <unistd.h>

#include "depecher_b200.h"

using namespace Rcpp;

static void check(int rc) {
    if (rc != DPR_OK) stop(std::string("depecher_b200: ") + dpr_last_error());
}

// ---- randomness ------------------------------------------------------------------------------------------------
// The reference re-seeds libc rand() with time(0) + seed_off_set in every call (src/Clusterer.cpp:25,
// src/Clusterer.h:51-53): two calls with the SAME offset still draw differently, and depechePenaltyOpt relies on that —
// it passes i = 1..nCores in every set of its while loop (R/depechePenaltyOpt.R:50-54).  The library's entry points take an
// explicit seed instead, so this layer supplies the per-call variation: a process-wide call counter mixed with a base that is
// time- and process-dependent by default, or fixed by set_engine_seed(seed) / the environment variable DEPECHER_B200_SEED for
// reproducible runs (the same sequence of calls then draws the same numbers; set_engine_seed restarts the sequence).
static std::atomic<uint64_t> g_calls(0);
static bool g_seed_fixed = false, g_seed_init = false;
static uint64_t g_seed_base = 0;

static uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static uint64_t next_seed(unsigned long seed_off_set) {
    if (!g_seed_init) {
        g_seed_init = true;
        if (const char* e = std::getenv("DEPECHER_B200_SEED")) {
            g_seed_fixed = true;
            g_seed_base = std::strtoull(e, NULL, 10);
        }
    }
    const uint64_t n = g_calls.fetch_add(1);
    const uint64_t base = g_seed_fixed ? g_seed_base : ((uint64_t)std::time(NULL) ^ ((uint64_t)getpid() << 32));
    return splitmix64(base ^ splitmix64(n)) ^ splitmix64(0xD1B54A32D192ED03ull * ((uint64_t)seed_off_set + 1));
}

// [[Rcpp::export]]
void set_engine_seed(const double seed) {
    g_seed_init = true;
    g_seed_fixed = true;
    g_seed_base = (uint64_t)seed;
    g_calls.store(0);
}

// ---- the four exports of the reference -----------------------------------------------------------------------------

// [[Rcpp::export]]
List sparse_k_means(NumericMatrix X, const unsigned int k, const double reg, const bool no_zero,
                    const unsigned long seed_off_set = 0) {
    const int rows = X.nrow(), cols = X.ncol();
    IntegerVector ints(rows);
    NumericMatrix centers(k, cols);
    double norm = 0.0;
    check(dpr_sparse_k_means(REAL(X), rows, cols, k, reg, no_zero ? 1 : 0, next_seed(seed_off_set), INTEGER(ints), REAL(centers), &norm,
                             NULL));
    List ret;
    ret["i"] = ints;
    ret["o"] = clone(ints);        // the reference returns copies under o / v / m (InterfaceUtils.cpp:86-91)
    ret["c"] = centers;
    ret["v"] = clone(centers);
    ret["n"] = norm;
    ret["m"] = norm;
    return ret;
}

// list of 4 matrices per (iteration, k, penalty); [2],[3] repeat [0],[1] (Clusterer.cpp:363-368)
static List centre_list(const std::vector<double>& flat, IntegerVector k, int nreg, unsigned int iterations, int cols) {
    const int nk = k.size();
    List numCenters((size_t)iterations * nk * nreg);
    size_t off = 0, cell = 0;
    for (unsigned int it = 0; it < iterations; ++it)
        for (int j = 0; j < nk; ++j)
            for (int l = 0; l < nreg; ++l) {
                List four(4);
                for (int w = 0; w < 2; ++w) {
                    NumericMatrix m(k[j], cols);
                    std::copy(flat.begin() + off, flat.begin() + off + (size_t)k[j] * cols, m.begin());
                    off += (size_t)k[j] * cols;
                    four[w] = m;
                    four[w + 2] = clone(m);
                }
                numCenters[cell++] = four;
            }
    return numCenters;
}

// [[Rcpp::export]]
List grid_search(NumericMatrix X, IntegerVector k, NumericVector reg, const unsigned int iterations,
                 const unsigned int bootstrapSamples, const unsigned long seed_off_set = 0) {
    const int rows = X.nrow(), cols = X.ncol(), nk = k.size(), nreg = reg.size();
    NumericMatrix distances(nk, nreg), found_cluster(nk, nreg);
    size_t total = 0;
    for (int j = 0; j < nk; ++j) total += (size_t)k[j];
    std::vector<double> flat((size_t)iterations * nreg * 2 * cols * total);
    check(dpr_grid_search(REAL(X), rows, cols, INTEGER(k), nk, REAL(reg), nreg, iterations, bootstrapSamples, next_seed(seed_off_set),
                          REAL(distances), REAL(found_cluster), flat.data()));
    List ret;
    ret["d"] = distances;
    ret["z"] = clone(distances);
    ret["n"] = found_cluster;
    ret["m"] = clone(found_cluster);
    ret["c"] = centre_list(flat, k, nreg, iterations, cols);
    return ret;
}

// [[Rcpp::export]]
List allocate_points(NumericMatrix X, NumericMatrix mu, const bool no_zero) {
    const int rows = X.nrow();
    if (mu.ncol() != X.ncol()) stop("allocate_points: X and mu differ in the number of columns");
    IntegerVector inds(rows);
    check(dpr_allocate_points(REAL(X), rows, X.ncol(), REAL(mu), mu.nrow(), no_zero ? 1 : 0, INTEGER(inds)));
    List ret;
    ret["i"] = inds;
    return ret;
}

// [[Rcpp::export]]
const double rand_index(IntegerVector inds1, IntegerVector inds2, unsigned int k) {
    if (inds1.size() != inds2.size()) stop("rand_index: the label vectors differ in length");
    double out = 0.0;
    check(dpr_rand_index(INTEGER(inds1), INTEGER(inds2), inds1.size(), k, &out));
    return out;
}

// ---- resident dataset: one upload per depeche() instead of one per call and worker ------------------------------------
// (the reference copies the matrix element by element on every call: numeric_to_eigen, src/InterfaceUtils.cpp:23-33,70,102,137)

static void dataset_finalizer(SEXP ptr) {
    dpr_dataset_t ds = (dpr_dataset_t)R_ExternalPtrAddr(ptr);
    if (ds) {
        dpr_dataset_destroy(ds);
        R_ClearExternalPtr(ptr);
    }
}
static SEXP wrap_dataset(dpr_dataset_t ds) {
    SEXP ptr = PROTECT(R_MakeExternalPtr(ds, R_NilValue, R_NilValue));
    R_RegisterCFinalizerEx(ptr, dataset_finalizer, TRUE);   // TRUE: also at the end of the R session
    UNPROTECT(1);
    return ptr;
}
static dpr_dataset_t dataset_of(SEXP handle) {
    dpr_dataset_t ds = (dpr_dataset_t)R_ExternalPtrAddr(handle);
    if (!ds) stop("depecher_b200: the dataset handle has been destroyed (or comes from another R session)");
    return ds;
}

// [[Rcpp::export]]
SEXP dataset_create(NumericMatrix X) {
    dpr_dataset_t ds = NULL;
    check(dpr_dataset_create(REAL(X), X.nrow(), X.ncol(), &ds));
    return wrap_dataset(ds);
}

// depecheLogCenterSd on the device (R/depecheLogCenterSd.R:6-115) + the upload in one call: the scaled matrix never exists in R.
// centerMode: 0 none, 1 histogram peak (R/dScaleCoFunction.R:74-88), 2 mean, 3 the values in `centers`; scaleMode: 0 none, 1 the sd of all
// centred values, 2 `scaleValue` (the R wrapper maps depeche()'s center = "default" / "peak" / "mean" / FALSE / numeric and scale = TRUE /
// FALSE / numeric onto these, R/depeche.R:131-141).  Returns list(handle = <external pointer>, log2Applied, center, scale) — the
// logCenterSd record depeche() stores for dAllocate (R/dAllocate.R:81-90).
// [[Rcpp::export]]
List dataset_create_scaled(NumericMatrix X, const bool log2Off, const int centerMode, NumericVector centers, const int scaleMode,
                           const double scaleValue) {
    const int d = X.ncol();
    if (centerMode == 3 && centers.size() != d) stop("dataset_create_scaled: one centre per column is needed");
    NumericVector used((long)d);
    for (int j = 0; j < d; ++j) used[j] = centerMode == 3 ? centers[j] : 0.0;
    double sd = scaleMode == 2 ? scaleValue : 1.0;
    int32_t log2_applied = 0;
    dpr_dataset_t ds = NULL;
    check(dpr_dataset_create_scaled(REAL(X), X.nrow(), d, log2Off ? 1 : 0, centerMode, REAL(used), scaleMode, &sd, &log2_applied, &ds));
    List ret;
    ret["handle"] = wrap_dataset(ds);
    ret["log2Applied"] = log2_applied != 0;
    ret["center"] = used;
    ret["scale"] = sd;
    return ret;
}

// [[Rcpp::export]]
void dataset_destroy(SEXP handle) { dataset_finalizer(handle); }

// rows: 1-based row numbers, with replacement (the selection set of R/depeche.R:191-199)
// [[Rcpp::export]]
SEXP dataset_gather(SEXP handle, IntegerVector rows) {
    std::vector<int64_t> r((size_t)rows.size());
    for (long i = 0; i < rows.size(); ++i) r[(size_t)i] = (int64_t)rows[i] - 1;
    dpr_dataset_t out = NULL;
    check(dpr_dataset_gather(dataset_of(handle), r.data(), (int64_t)r.size(), &out));
    return wrap_dataset(out);
}

// [[Rcpp::export]]
List dataset_allocate_points(SEXP handle, NumericMatrix mu, const bool no_zero) {
    dpr_dataset_t ds = dataset_of(handle);
    int64_t n = 0;
    int32_t d = 0;
    check(dpr_dataset_shape(ds, &n, &d));
    if (mu.ncol() != d) stop("dataset_allocate_points: the dataset and mu differ in the number of columns");
    IntegerVector inds((long)n);
    check(dpr_dataset_allocate_points(ds, REAL(mu), mu.nrow(), no_zero ? 1 : 0, INTEGER(inds)));
    List ret;
    ret["i"] = inds;
    return ret;
}

// depeche()'s final step (R/depeche.R:228-252) in one call: dAllocate of the resident matrix and the renumbering of the clusters by
// decreasing size (sizes counted on the device, the new numbers applied while the labels are copied out).  Returns list(i = the new
// cluster numbers 1, 2, ..., order = the rows of mu (1-based) in rank order, 0 beyond the clusters that own cells).
// [[Rcpp::export]]
List dataset_allocate_points_ranked(SEXP handle, NumericMatrix mu, const bool no_zero) {
    dpr_dataset_t ds = dataset_of(handle);
    int64_t n = 0;
    int32_t d = 0;
    check(dpr_dataset_shape(ds, &n, &d));
    if (mu.ncol() != d) stop("dataset_allocate_points_ranked: the dataset and mu differ in the number of columns");
    IntegerVector inds((long)n);
    IntegerVector order((long)mu.nrow());   // the rows of mu (1-based) in rank order; 0 beyond the clusters that own cells
    check(dpr_dataset_allocate_points_ranked(ds, REAL(mu), mu.nrow(), no_zero ? 1 : 0, INTEGER(inds), INTEGER(order)));
    for (long r = 0; r < order.size(); ++r) order[r] += 1;
    List ret;
    ret["i"] = inds;
    ret["order"] = order;
    return ret;
}

// The former foreach over SOCK workers (R/depechePenaltyOpt.R:50-55) as one call: a list with one entry per worker, each what
// grid_search(dataMat, k, usedPenalties, 1, sampleSize, i) returns (keys d, z, n, m, c).
// [[Rcpp::export]]
List grid_search_each(SEXP handle, IntegerVector k, NumericVector reg, const unsigned int iterations,
                      const unsigned int bootstrapSamples, const unsigned long seed_off_set = 0) {
    dpr_dataset_t ds = dataset_of(handle);
    int64_t n = 0;
    int32_t cols = 0;
    check(dpr_dataset_shape(ds, &n, &cols));
    const int nk = k.size(), nreg = reg.size();
    size_t total = 0;
    for (int j = 0; j < nk; ++j) total += (size_t)k[j];
    const size_t m = (size_t)nk * nreg, per_worker = (size_t)nreg * 2 * cols * total;
    std::vector<double> ari(m * iterations), ncl(m * iterations), flat(per_worker * iterations);
    check(dpr_dataset_grid_search_each(ds, INTEGER(k), nk, REAL(reg), nreg, iterations, bootstrapSamples, next_seed(seed_off_set), ari.data(),
                                       ncl.data(), flat.data()));
    List workers((size_t)iterations);
    for (unsigned int w = 0; w < iterations; ++w) {
        NumericMatrix distances(nk, nreg), found_cluster(nk, nreg);
        std::copy(ari.begin() + w * m, ari.begin() + (w + 1) * m, distances.begin());
        std::copy(ncl.begin() + w * m, ncl.begin() + (w + 1) * m, found_cluster.begin());
        std::vector<double> mine(flat.begin() + w * per_worker, flat.begin() + (w + 1) * per_worker);
        List ret;
        ret["d"] = distances;
        ret["z"] = clone(distances);
        ret["n"] = found_cluster;
        ret["m"] = clone(found_cluster);
        ret["c"] = centre_list(mine, k, nreg, 1, cols);
        workers[(size_t)w] = ret;
    }
    return workers;
}

// depechePenaltyOpt's adaptive loop (R/depechePenaltyOpt.R:47-192, 242-258) in one call.  Returns bestPenalty (rounded to one
// digit like the reference's names), bestPosition (1-based), ARI and nClust per penalty, workerIterations, warn (bit 0: iteration
// cap reached before stability, bit 1 / 2: best penalty is the lowest / highest) and bestClusterCenters (list of k x d matrices).
// [[Rcpp::export]]
List penalty_opt(SEXP handle, NumericVector penalties, const unsigned int k, const unsigned int sampleSize, const unsigned int nCores,
                 const unsigned int maxIter, const double minARIImprovement, const double optimARI, const unsigned long seed_off_set = 0) {
    dpr_dataset_t ds = dataset_of(handle);
    int64_t n = 0;
    int32_t cols = 0;
    check(dpr_dataset_shape(ds, &n, &cols));
    const int n_pen = penalties.size();
    dpr_penalty_opt_t opt = NULL;
    check(dpr_penalty_opt_create(REAL(penalties), n_pen, (int32_t)k, cols, sampleSize, (int32_t)nCores, (int32_t)maxIter, minARIImprovement, optimARI, &opt));
    int rc = dpr_dataset_penalty_opt(ds, opt, next_seed(seed_off_set), 0);
    double best = 0.0;
    int32_t pos = 0, iterations = 0, n_sol = 0, warn = 0;
    NumericVector ari(n_pen), nclust(n_pen);
    if (rc == DPR_OK) rc = dpr_penalty_opt_result(opt, &best, &pos, REAL(ari), REAL(nclust), &iterations, &n_sol, &warn);
    std::vector<double> flat((size_t)n_sol * k * cols);
    if (rc == DPR_OK && n_sol > 0) rc = dpr_penalty_opt_solutions(opt, flat.data());
    dpr_penalty_opt_destroy(opt);
    check(rc);
    List solutions((size_t)n_sol);
    for (int s = 0; s < n_sol; ++s) {
        NumericMatrix m(k, cols);
        std::copy(flat.begin() + (size_t)s * k * cols, flat.begin() + (size_t)(s + 1) * k * cols, m.begin());
        solutions[(size_t)s] = m;
    }
    List ret;
    ret["bestPenalty"] = best;
    ret["bestPosition"] = pos + 1;
    ret["ARI"] = ari;
    ret["nClust"] = nclust;
    ret["workerIterations"] = iterations;
    ret["warn"] = warn;
    ret["bestClusterCenters"] = solutions;
    return ret;
}

// depecheClusterCenterSelection.R:11-37: the selection set allocated to every stored solution and all pairs scored in one pass.
// Returns the 1-based index of the most central solution and the mean ARI of each.
// [[Rcpp::export]]
List select_solution(SEXP handle, List allSolutions, const unsigned int k) {
    dpr_dataset_t ds = dataset_of(handle);
    int64_t n = 0;
    int32_t cols = 0;
    check(dpr_dataset_shape(ds, &n, &cols));
    const int m = (int)allSolutions.size();
    std::vector<double> flat((size_t)m * k * cols);
    for (int s = 0; s < m; ++s) {
        NumericMatrix mat = as<NumericMatrix>(allSolutions[(size_t)s]);
        if (mat.nrow() != (int)k || mat.ncol() != cols) stop("select_solution: every solution must be a k x ncol(data) matrix");
        std::copy(mat.begin(), mat.begin() + (size_t)k * cols, flat.begin() + (size_t)s * k * cols);
    }
    int32_t best = 0;
    NumericVector meanARI(m);
    check(dpr_dataset_select_solution(ds, flat.data(), m, (int32_t)k, &best, REAL(meanARI)));
    List ret;
    ret["best"] = best + 1;
    ret["meanARI"] = meanARI;
    return ret;
}
