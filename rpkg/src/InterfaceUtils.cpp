// InterfaceUtils.cpp — Rcpp host layer of the B200 build: drop-in replacement for the reference's
// src/InterfaceUtils.cpp.  The four exported functions keep their names, argument lists and the keys of the
// lists they return (/root/reference/src/InterfaceUtils.cpp:66-161), so R/RcppExports.R, src/RcppExports.cpp
// and every R caller stay byte-for-byte what they are.  The bodies no longer build Eigen matrices and a
// Clusterer: they hand R's own column-major buffers to the C ABI of libdepecher_b200 (include/depecher_b200.h).
//
// Differences a user can observe:
//   * seed_off_set is used as the seed itself (the reference adds it to time(0), src/Clusterer.h:51-53), so a
//     call is reproducible; pass a different offset for a different draw, as depechePenaltyOpt() already does.
//   * rand_index returns the contingency-table value the reference's 10 000-pair estimate scatters around.
//   * errors (no B200, bad arguments) surface as R errors through Rcpp::stop.
#include <Rcpp.h>

#include "depecher_b200.h"

using namespace Rcpp;

static void check(int rc) {
    if (rc != DPR_OK) stop(std::string("depecher_b200: ") + dpr_last_error());
}

// [[Rcpp::export]]
List sparse_k_means(NumericMatrix X, const unsigned int k, const double reg, const bool no_zero,
                    const unsigned long seed_off_set = 0) {
    const int rows = X.nrow(), cols = X.ncol();
    IntegerVector ints(rows);
    NumericMatrix centers(k, cols);
    double norm = 0.0;
    check(dpr_sparse_k_means(REAL(X), rows, cols, k, reg, no_zero ? 1 : 0, (uint64_t)seed_off_set, INTEGER(ints), REAL(centers), &norm,
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

// [[Rcpp::export]]
List grid_search(NumericMatrix X, IntegerVector k, NumericVector reg, const unsigned int iterations,
                 const unsigned int bootstrapSamples, const unsigned long seed_off_set = 0) {
    const int rows = X.nrow(), cols = X.ncol(), nk = k.size(), nreg = reg.size();
    NumericMatrix distances(nk, nreg), found_cluster(nk, nreg);
    size_t total = 0;
    for (int j = 0; j < nk; ++j) total += (size_t)k[j];
    std::vector<double> flat((size_t)iterations * nreg * 2 * cols * total);
    check(dpr_grid_search(REAL(X), rows, cols, INTEGER(k), nk, REAL(reg), nreg, iterations, bootstrapSamples, (uint64_t)seed_off_set,
                          REAL(distances), REAL(found_cluster), flat.data()));
    // list of 4 matrices per (iteration, k, penalty); [2],[3] repeat [0],[1] (Clusterer.cpp:363-368)
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
    List ret;
    ret["d"] = distances;
    ret["z"] = clone(distances);
    ret["n"] = found_cluster;
    ret["m"] = clone(found_cluster);
    ret["c"] = numCenters;
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
