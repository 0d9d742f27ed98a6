/*
 * dpr_oracle.h — C interface of the CPU parity oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain-C++ restatement of the reference's penalized
 * k-means engine (/root/reference/src/Clusterer.cpp:24-397 and the four exports in
 * src/InterfaceUtils.cpp:66-161).  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline may load it; the shipped library (depecher_b200/) never does.
 *
 * Pinning: the reference ships no golden vectors for these functions (its tests are stochastic
 * end-to-end properties, tests/testthat/test_{depeche,dAllocate,dOptPenalty}.R).  The restatement
 * is pinned instead against the reference's OWN source: oracle/_ref/libdepecher_ref.so is
 * src/Clusterer.cpp compiled where it lies against a small scalar stand-in for the Eigen/Rcpp
 * headers (oracle/eigen_shim/), and tests/test_oracle_vs_ref.py demands bit-identical labels,
 * centres, ARI and cluster counts from both under the same libc rand() stream.  The transcribed
 * testthat properties and the stored depeche() result in data/testDataDepeche.RData are the
 * second pin (tests/golden/).  What stays unpinned is Eigen's own arithmetic (summation order
 * inside norm()/sum(), NaN handling of cwiseMax): Eigen is not in this image.
 *
 * Matrices cross this interface COLUMN-MAJOR (R layout) like the product's C ABI; inside, the
 * oracle works on row-major doubles like the reference (typedef RowMatrixXd, src/Clusterer.h:19).
 */
#ifndef DPR_ORACLE_H
#define DPR_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* randomness source: 0 = libc rand() in the reference's call order (pins against _ref);
 *                    1 = Philox streams of include/dpr_philox.h (pins the CUDA library). */
#define ORACLE_RNG_LIBC 0
#define ORACLE_RNG_PHILOX 1
void oracle_set_rng(int mode, uint64_t seed);
/* libc mode: srand(seed32) — what Clusterer::reseed does with time(0)+offset (src/Clusterer.h:51-53) */
void oracle_srand(unsigned int seed32);
/* reproduce (1) or skip (0) the two discarded no_zero allocations of optimize_param (:373-374) */
void oracle_set_dead_work(int on);

int oracle_allocate_points(const double* X, int64_t n, int32_t d, const double* mu, int32_t k,
                           int32_t no_zero, int32_t* idx_out);
int oracle_sparse_k_means(const double* X, int64_t n, int32_t d, uint32_t k, double reg, int32_t no_zero,
                          uint32_t fit_id, int32_t* idx_out, double* centers_out, double* norm_out,
                          int32_t* n_iter_out);
int oracle_grid_search(const double* X, int64_t n, int32_t d, const int32_t* k, int32_t nk,
                       const double* reg, int32_t nreg, uint32_t iterations, uint32_t boot_n,
                       double* ari_out, double* nclust_out, double* centers_out);
/* Monte-Carlo estimate exactly as the reference (10 000 pairs); pair_id selects the Philox stream */
int oracle_rand_index(const int32_t* a, const int32_t* b, int64_t n, uint32_t k, uint32_t pair_id, double* out);
/* the value the Monte-Carlo estimate converges to (contingency table); not in the reference, used
 * to bound the sampling error in tests */
int oracle_rand_index_exact(const int32_t* a, const int32_t* b, int64_t n, uint32_t k, double* out);

/* stage-level */
int oracle_find_centers_from(const double* X, int64_t n, int32_t d, const double* init_mu, uint32_t k,
                             double reg, int32_t no_zero, int32_t max_iter, int32_t* idx_out,
                             double* centers_out, double* norm_out, int32_t* n_iter_out);
int oracle_reevaluate_centers(const double* X, int64_t n, int32_t d, const int32_t* idx, uint32_t k,
                              double reg, double* centers_out, int64_t* counts_out);
int oracle_initialize_mu(const double* X, int64_t n, int32_t d, uint32_t k, uint32_t fit_id,
                         double* mu_out, int64_t* rows_out);
int oracle_bootstrap_rows(int64_t n, uint32_t boot_n, uint32_t fit_id, int64_t* rows_out);
int oracle_cluster_norm(const double* X, int64_t n, int32_t d, const double* mu, uint32_t k,
                        const int32_t* idx, double reg, double* out);
int oracle_n_used_clusters(const int32_t* idx, int64_t n, uint32_t k, uint32_t* out);
/* relative top-2 distance margin per row (documents near-ties): (d2-d1)/d1 on the sqrt distances */
int oracle_assignment_margin(const double* X, int64_t n, int32_t d, const double* mu, int32_t k,
                             int32_t no_zero, double* margin_out);

/* CPU baseline: `iters` Lloyd iterations (allocate_clusters + reevaluate_centers, reference cost
 * structure) on a row-major copy made once; returns seconds spent in the iterations only. */
int oracle_time_lloyd(const double* X, int64_t n, int32_t d, const double* mu, uint32_t k, double reg,
                      int32_t no_zero, int32_t iters, double* seconds_out);

#ifdef __cplusplus
}
#endif
#endif
