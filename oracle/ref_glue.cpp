/*
 * ref_glue.cpp — C entry points around the REFERENCE's own engine (TEST INFRASTRUCTURE).
 *
 * Built only by oracle/Makefile into oracle/_ref/libdepecher_ref.so together with
 * /root/reference/src/Clusterer.cpp, which is compiled where it lies (never copied) against the
 * header stand-ins in oracle/eigen_shim/.  The wrappers restate the marshalling of the four Rcpp
 * exports (src/InterfaceUtils.cpp:66-161: column-major R matrix -> RowMatrixXd element by element,
 * Clusterer c; c.reseed(seed_off_set); ...) with plain pointers instead of Rcpp types.
 *
 * `class Tester` is the friend the reference declares (src/Clusterer.h:56) but never defines; it
 * gives the stage-level tests access to the private reevaluate_centers / initialize_mu / ...
 */
#include "Clusterer.h"

#include <chrono>
#include <cstring>

static long g_fixed_time = 1234567;
extern "C" long dpr_ref_fixed_time(void) { return g_fixed_time; }

static RowMatrixXd from_colmajor(const double* cm, long rows, long cols) {
    RowMatrixXd m = RowMatrixXd::Zero(rows, cols);
    for (long i = 0; i < rows; ++i)
        for (long j = 0; j < cols; ++j) m(i, j) = cm[j * rows + i];
    return m;
}
static void to_colmajor(const RowMatrixXd& m, double* cm) {
    for (long i = 0; i < m.rows(); ++i)
        for (long j = 0; j < m.cols(); ++j) cm[j * m.rows() + i] = m(i, j);
}
static Eigen::VectorXi to_vec(const int32_t* p, long n) {
    Eigen::VectorXi v = Eigen::VectorXi::Zero(n);
    for (long i = 0; i < n; ++i) v(i) = p[i];
    return v;
}

class Tester {
public:
    static RowMatrixXd reevaluate(Clusterer& c, const RowMatrixXd& X, const Eigen::VectorXi& idx, unsigned k, double reg) {
        return c.reevaluate_centers(X, idx, k, reg);
    }
    static RowMatrixXd init_mu(Clusterer& c, const RowMatrixXd& X, unsigned k) { return c.initialize_mu(X, k); }
    static RowMatrixXd bootstrap(Clusterer& c, const RowMatrixXd& X, unsigned s) { return c.bootstrap_data(X, s); }
    static double norm(Clusterer& c, const RowMatrixXd& X, const RowMatrixXd& mu, const Eigen::VectorXi& idx, double reg) {
        return c.cluster_norm(X, mu, idx, reg);
    }
    static unsigned n_used(Clusterer& c, unsigned k, const Eigen::VectorXi& idx) { return c.n_used_clusters(k, idx); }
};

extern "C" {

/* the value time(0) returns inside the reference; srand(time(0)+offset) then fixes the stream */
void ref_set_time(long t) { g_fixed_time = t; }

int ref_allocate_points(const double* X, int64_t n, int32_t d, const double* mu, int32_t k, int32_t no_zero,
                        int32_t* idx_out) {
    Clusterer c = Clusterer();
    RowMatrixXd data = from_colmajor(X, n, d), m = from_colmajor(mu, k, d);
    Eigen::VectorXi idx = c.allocate_clusters(data, m, no_zero != 0);
    for (int64_t i = 0; i < n; ++i) idx_out[i] = idx(i);
    return 0;
}

int ref_sparse_k_means(const double* X, int64_t n, int32_t d, uint32_t k, double reg, int32_t no_zero,
                       unsigned long seed_off_set, int32_t* idx_out, double* centers_out, double* norm_out) {
    RowMatrixXd data = from_colmajor(X, n, d);
    Clusterer c = Clusterer();
    c.reseed(seed_off_set);
    const Return_values part = c.find_centers(data, k, reg, no_zero != 0);
    for (int64_t i = 0; i < n; ++i) idx_out[i] = part.indexes(i);
    to_colmajor(part.centers, centers_out);
    *norm_out = part.norm;
    return 0;
}

int ref_grid_search(const double* X, int64_t n, int32_t d, const int32_t* k, int32_t nk, const double* reg,
                    int32_t nreg, uint32_t iterations, uint32_t boot_n, unsigned long seed_off_set,
                    double* ari_out, double* nclust_out, double* centers_out) {
    RowMatrixXd data = from_colmajor(X, n, d);
    Eigen::Matrix<unsigned int, -1, 1> kv = Eigen::Matrix<unsigned int, -1, 1>::Zero(nk);
    for (int32_t j = 0; j < nk; ++j) kv(j) = (unsigned)k[j];
    Eigen::VectorXd rv = Eigen::VectorXd::Zero(nreg);
    for (int32_t l = 0; l < nreg; ++l) rv(l) = reg[l];
    Clusterer c = Clusterer();
    c.reseed(seed_off_set);
    Optimization_values vals = c.optimize_param(data, kv, rv, iterations, boot_n);
    to_colmajor(vals.distances, ari_out);
    to_colmajor(vals.found_cluster, nclust_out);
    if (centers_out) {
        size_t off = 0;
        for (size_t p = 0; p < vals.centers.size(); ++p)
            for (int w = 0; w < 2; ++w) {
                to_colmajor(vals.centers[p][w], centers_out + off);
                off += (size_t)vals.centers[p][w].size();
            }
    }
    return 0;
}

int ref_rand_index(const int32_t* a, const int32_t* b, int64_t n, uint32_t k, unsigned long seed_off_set, double* out) {
    Clusterer c = Clusterer();
    c.reseed(seed_off_set);
    *out = c.cluster_distance(to_vec(a, n), to_vec(b, n), k);
    return 0;
}

int ref_reevaluate_centers(const double* X, int64_t n, int32_t d, const int32_t* idx, uint32_t k, double reg,
                           double* centers_out) {
    Clusterer c = Clusterer();
    RowMatrixXd data = from_colmajor(X, n, d);
    to_colmajor(Tester::reevaluate(c, data, to_vec(idx, n), k, reg), centers_out);
    return 0;
}

int ref_initialize_mu(const double* X, int64_t n, int32_t d, uint32_t k, unsigned long seed_off_set, double* mu_out) {
    Clusterer c = Clusterer();
    c.reseed(seed_off_set);
    RowMatrixXd data = from_colmajor(X, n, d);
    to_colmajor(Tester::init_mu(c, data, k), mu_out);
    return 0;
}

int ref_bootstrap_data(const double* X, int64_t n, int32_t d, uint32_t boot_n, unsigned long seed_off_set, double* out) {
    Clusterer c = Clusterer();
    c.reseed(seed_off_set);
    RowMatrixXd data = from_colmajor(X, n, d);
    to_colmajor(Tester::bootstrap(c, data, boot_n), out);
    return 0;
}

int ref_cluster_norm(const double* X, int64_t n, int32_t d, const double* mu, uint32_t k, const int32_t* idx,
                     double reg, double* out) {
    Clusterer c = Clusterer();
    *out = Tester::norm(c, from_colmajor(X, n, d), from_colmajor(mu, k, d), to_vec(idx, n), reg);
    return 0;
}

int ref_n_used_clusters(const int32_t* idx, int64_t n, uint32_t k, uint32_t* out) {
    Clusterer c = Clusterer();
    *out = Tester::n_used(c, k, to_vec(idx, n));
    return 0;
}

/* CPU baseline: `iters` x (allocate_clusters + reevaluate_centers) of the reference itself */
int ref_time_lloyd(const double* X, int64_t n, int32_t d, const double* mu, uint32_t k, double reg, int32_t no_zero,
                   int32_t iters, double* seconds_out) {
    Clusterer c = Clusterer();
    RowMatrixXd data = from_colmajor(X, n, d), m = from_colmajor(mu, k, d);
    auto t0 = std::chrono::steady_clock::now();
    for (int it = 0; it < iters; ++it) {
        Eigen::VectorXi idx = c.allocate_clusters(data, m, no_zero != 0);
        m = Tester::reevaluate(c, data, idx, k, reg);
    }
    auto t1 = std::chrono::steady_clock::now();
    *seconds_out = std::chrono::duration<double>(t1 - t0).count();
    return 0;
}

}  // extern "C"
