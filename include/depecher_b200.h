/*
 * depecher_b200.h — C ABI of the B200-native penalized k-means engine.
 *
 * This is the drop-in boundary for DepecheR's hot path: the four entry points below replace,
 * one for one, the bodies of the four Rcpp exports of the reference
 *     sparse_k_means   src/InterfaceUtils.cpp:66-94
 *     grid_search      src/InterfaceUtils.cpp:99-131
 *     allocate_points  src/InterfaceUtils.cpp:134-148
 *     rand_index       src/InterfaceUtils.cpp:151-161
 * which are what `.Call('_DepecheR_*')` (R/RcppExports.R:4-18, src/RcppExports.cpp:14-83)
 * reaches.  Plain pointers and sizes only; matrices are COLUMN-MAJOR doubles exactly as R
 * hands them over (no transpose at the boundary, unlike numeric_to_eigen,
 * src/InterfaceUtils.cpp:23-33); labels are 0-based int32 like the reference returns them.
 *
 * All functions return DPR_OK (0) or a negative error code; dpr_last_error() gives the text
 * (thread-local).  There is NO CPU fallback: without a usable sm_100 device every compute
 * entry point fails with DPR_ERR_NO_DEVICE.
 *
 * Randomness: the reference seeds libc rand() from time(0) (src/Clusterer.h:51-53).  Here
 * every entry point that draws takes an explicit 64-bit seed; the stream layout is fixed in
 * include/dpr_philox.h so results are reproducible and checkable against the oracle.
 */
#ifndef DEPECHER_B200_H
#define DEPECHER_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DPR_OK 0
#define DPR_ERR_INVALID (-1)   /* bad argument (k < 2, n < 1, null pointer, label out of range ...) */
#define DPR_ERR_NO_DEVICE (-2) /* no sm_100 CUDA device / driver */
#define DPR_ERR_CUDA (-3)      /* CUDA runtime error; text in dpr_last_error() */
#define DPR_ERR_NCCL (-4)      /* NCCL error or NCCL not loadable */
#define DPR_ERR_STATE (-5)     /* call order / handle misuse */

/* rand_index estimator: the contingency-table value is the expectation of the reference's
 * 10 000-pair Monte-Carlo estimate (src/Clusterer.cpp:286-304). */
#define DPR_ARI_EXACT 0
#define DPR_ARI_MONTE_CARLO 1

typedef struct dpr_dataset* dpr_dataset_t;

/* ---- housekeeping ------------------------------------------------------------------- */
const char* dpr_last_error(void);
int dpr_version(void);                   /* 10000*major + 100*minor + patch */
int dpr_device_count(int32_t* out);
int dpr_set_device(int32_t ordinal);     /* default: device 0 (or LOCAL_RANK via the host layer) */
int dpr_set_stream(void* cuda_stream);   /* run everything on the caller's cudaStream_t (NULL = library stream) */
int dpr_synchronize(void);
int dpr_set_ari_mode(int32_t mode);      /* DPR_ARI_EXACT (default) | DPR_ARI_MONTE_CARLO */
/* which kernel runs the pass over the cells: DPR_PASS_AUTO (default: the tcgen05 tensor-core pass when the shape fits
 * tensor memory, else the FFMA pass) | DPR_PASS_FFMA (always the FFMA pass; A/B measurements and tests).  Both give
 * the same labels. */
/* diagnostics: cells whose label was decided by the exact fp64 path (two best fp32 scores closer than the error bound)
 * since the last reset */
int dpr_recheck_count(int64_t* out, int32_t reset);
#define DPR_PASS_AUTO 0
#define DPR_PASS_FFMA 1
int dpr_set_pass_kernel(int32_t which);
/* counters for bench/test evidence: kernels launched by this library since the last reset */
int dpr_kernel_launches(int64_t* out, int32_t reset);
/* bench evidence: with enable != 0 every launch of the fused pass is bracketed by CUDA events on the
 * launching stream; the call returns (and resets) the summed duration and the number of launches. */
int dpr_pass_timing(int32_t enable, double* ms_sum_out, int64_t* launches_out);

/* ---- the four drop-in entry points (host buffers in, host buffers out) ------------------ */

/* allocate_points(X, mu, no_zero) -> i       src/InterfaceUtils.cpp:134-148 -> Clusterer::allocate_clusters
 * (src/Clusterer.cpp:34-87).  X: n x d col-major, mu: k x d col-major, idx_out: n labels (row of mu). */
int dpr_allocate_points(const double* X, int64_t n, int32_t d, const double* mu, int32_t k,
                        int32_t no_zero, int32_t* idx_out);

/* sparse_k_means(X, k, reg, no_zero, seed_off_set) -> {i,c,n}   src/InterfaceUtils.cpp:66-94 ->
 * Clusterer::find_centers (src/Clusterer.cpp:213-265).  centers_out: k x d col-major; norm_out: the
 * penalised SSQ of cluster_norm (:198-211); n_iter_out (nullable): assignments performed. */
int dpr_sparse_k_means(const double* X, int64_t n, int32_t d, uint32_t k, double reg, int32_t no_zero,
                       uint64_t seed, int32_t* idx_out, double* centers_out, double* norm_out,
                       int32_t* n_iter_out);

/* grid_search(X, k, reg, iterations, bootstrapSamples, seed_off_set) -> {d,n,c}
 * src/InterfaceUtils.cpp:99-131 -> Clusterer::optimize_param (src/Clusterer.cpp:334-394).
 * ari_out, nclust_out: nk x nreg col-major (the reference's d/z and n/m, which are copies of each other,
 * :387-390).  centers_out (nullable): for every problem (it, j, l) in the reference's loop order
 * (:349-352) the two fitted centre matrices ret1, ret2, each k[j] x d col-major, packed back to back
 * (entries [3],[4] of the reference's list are duplicates of [1],[2], :363-368). */
int dpr_grid_search(const double* X, int64_t n, int32_t d, const int32_t* k, int32_t nk,
                    const double* reg, int32_t nreg, uint32_t iterations, uint32_t boot_n,
                    uint64_t seed, double* ari_out, double* nclust_out, double* centers_out);

/* rand_index(inds1, inds2, k) -> double      src/InterfaceUtils.cpp:151-161 -> Clusterer::cluster_distance
 * (src/Clusterer.cpp:267-305).  Labels must lie in [0, k] (R passes 1-based labels with k = max, SURVEY A-9). */
int dpr_rand_index(const int32_t* a, const int32_t* b, int64_t n, uint32_t k, double* out);

/* ---- resident dataset: X stays in HBM (fp32, column-major) across calls ---------------------- */
/* The reference re-marshals X on every call (src/InterfaceUtils.cpp:70,102,137); depeche() makes
 * dozens of grid_search calls on the same matrix (R/depechePenaltyOpt.R:47-55). */
int dpr_dataset_create(const double* X, int64_t n, int32_t d, dpr_dataset_t* out);
int dpr_dataset_create_f32(const float* X, int64_t n, int32_t d, dpr_dataset_t* out);
/* synthetic generator on device (bench only): G blobs, centre coords in {-a,0,a}, unit noise, / sd. */
/* row0: global index of the first cell (row shards of one virtual matrix draw disjoint cells from the
 * same mixture); scale: divisor applied to every value, 0 = this shard's own standard deviation; the
 * divisor actually used is returned in *scale_out (nullable). */
int dpr_dataset_create_synthetic(int64_t n, int32_t d, int32_t groups, double amplitude, uint64_t seed,
                                 int64_t row0, double scale, double* scale_out, dpr_dataset_t* out);
/* depecheLogCenterSd on the device (R/depecheLogCenterSd.R:6-115; "peak" centre: R/dScaleCoFunction.R:74-88): optional log2
 * transform when the pooled kurtosis exceeds 100, per-column centring, division by one global standard deviation, then the
 * resident dataset of the scaled cells.  center_mode: 0 none, 1 histogram peak, 2 mean, 3 given in centers_inout.
 * scale_mode: 0 none, 1 sd of all centred values, 2 given in *scale_inout.  On return centers_inout[d] (nullable for mode 0),
 * *scale_inout and *log2_applied_out hold what was applied — the logCenterSd record dAllocate needs (R/dAllocate.R:81-90). */
int dpr_dataset_create_scaled(const double* X, int64_t n, int32_t d, int32_t log2_off, int32_t center_mode,
                              double* centers_inout, int32_t scale_mode, double* scale_inout, int32_t* log2_applied_out,
                              dpr_dataset_t* out);
int dpr_dataset_destroy(dpr_dataset_t ds);
int dpr_dataset_shape(dpr_dataset_t ds, int64_t* n, int32_t* d);
/* copy rows [row0, row0+rows) back as column-major doubles (tests / CPU baseline sample) */
int dpr_dataset_read(dpr_dataset_t ds, int64_t row0, int64_t rows, double* out);

int dpr_dataset_allocate_points(dpr_dataset_t ds, const double* mu, int32_t k, int32_t no_zero,
                                int32_t* idx_out /* host, nullable */);
/* the same with `base` added to every label on the way out: base = 1 gives R's 1-based cluster numbers
 * (R/dAllocate.R:127-134) without another pass over a vector with one entry per cell */
int dpr_dataset_allocate_points_base(dpr_dataset_t ds, const double* mu, int32_t k, int32_t no_zero, int32_t base,
                                     int32_t* idx_out /* host, nullable */);
/* depeche()'s final step (R/depeche.R:228-252): allocate and renumber the clusters 1, 2, ... by decreasing size (ties in the order
 * of the old numbers).  order_out[k]: the old 0-based label that became 1, 2, ...; -1 beyond the clusters that own cells. */
int dpr_dataset_allocate_points_ranked(dpr_dataset_t ds, const double* mu, int32_t k, int32_t no_zero,
                                       int32_t* idx_out /* host */, int32_t* order_out /* host, k */);
int dpr_dataset_sparse_k_means(dpr_dataset_t ds, uint32_t k, double reg, int32_t no_zero, uint64_t seed,
                               int32_t* idx_out /* nullable */, double* centers_out, double* norm_out,
                               int32_t* n_iter_out);
int dpr_dataset_grid_search(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg,
                            uint32_t iterations, uint32_t boot_n, uint64_t seed, double* ari_out,
                            double* nclust_out, double* centers_out);
/* The same work with one result per iteration instead of their mean: ari_each/nclust_each hold `iterations`
 * matrices (nk x nreg col-major, iteration-major).  This is what the reference gets from `iterations` separate
 * SOCK worker processes each calling grid_search(..., iterations = 1, ...) (R/depechePenaltyOpt.R:50-55);
 * here all of them are one batch of problems on the device. */
int dpr_dataset_grid_search_each(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg,
                                 uint32_t iterations, uint32_t boot_n, uint64_t seed, double* ari_each,
                                 double* nclust_each, double* centers_out);

/* A contiguous range of the grid's problems (grid split over processes or GPUs): problem c = (iteration * nk + j) * nreg + l in
 * the reference's loop order (src/Clusterer.cpp:349-352).  Random streams are addressed by the problem's number in the whole
 * grid, so a range computed on its own gives the labels, ARI and cluster counts of that range of the whole grid exactly, and its
 * centres to the rounding of their fp64 sums (the order of those additions follows the launch's partition of the tiles).  ari_out / nclust_out: n_problems values;
 * centers_out (nullable): ret1, ret2 (k[j] x d col-major) per problem. */
int dpr_dataset_grid_search_range(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg,
                                  uint32_t iterations, uint32_t boot_n, uint64_t seed, int32_t first_problem,
                                  int32_t n_problems, double* ari_out, double* nclust_out, double* centers_out);

/* ---- stage-level entry points (parity hooks; each is one deterministic stage of the reference) ----- */

/* Lloyd loop from injected initial centres: find_centers (src/Clusterer.cpp:244-252) without
 * initialize_mu.  init_mu/centers_out k x d col-major. */
int dpr_dataset_find_centers_from(dpr_dataset_t ds, const double* init_mu, uint32_t k, double reg,
                                  int32_t no_zero, int32_t max_iter, int32_t* idx_out, double* centers_out,
                                  double* norm_out, int32_t* n_iter_out);
/* one centre update for given labels: reevaluate_centers (src/Clusterer.cpp:89-108). */
int dpr_dataset_reevaluate_centers(dpr_dataset_t ds, const int32_t* idx, uint32_t k, double reg,
                                   double* centers_out, int64_t* counts_out /* nullable */);
/* k-means++ seeding: initialize_mu + element_from_vector (src/Clusterer.cpp:123-176); stream = fit id.
 * rows_out (nullable): the k chosen row numbers. */
int dpr_dataset_initialize_mu(dpr_dataset_t ds, uint32_t k, uint64_t seed, uint32_t fit_id,
                              double* mu_out, int64_t* rows_out);
/* bootstrap row numbers of bootstrap_data (src/Clusterer.cpp:308-321) for one fit id */
int dpr_bootstrap_rows(int64_t n, uint32_t boot_n, uint64_t seed, uint32_t fit_id, int64_t* rows_out);
/* resample a dataset by explicit rows (with replacement) into a new resident dataset */
int dpr_dataset_gather(dpr_dataset_t ds, const int64_t* rows, int64_t m, dpr_dataset_t* out);
/* cluster_norm (src/Clusterer.cpp:198-211) and n_used_clusters (:323-332) */
int dpr_dataset_cluster_norm(dpr_dataset_t ds, const double* mu, uint32_t k, const int32_t* idx, double reg,
                             double* out);
int dpr_n_used_clusters(const int32_t* idx, int64_t n, uint32_t k, uint32_t* out);
/* batched scoring: allocate the dataset to m centre sets (no_zero per call) in ONE pass over X and
 * return the all-pairs ARI matrix (m x m, col-major) of depecheClusterCenterSelection
 * (R/depecheClusterCenterSelection.R:11-29).  mus: m matrices k x d col-major back to back.
 * labels_out (nullable): m x n int32 (set-major). */
int dpr_dataset_allocate_many(dpr_dataset_t ds, const double* mus, int32_t m, int32_t k, int32_t no_zero,
                              int32_t* labels_out, double* ari_out);

/* ---- depechePenaltyOpt and depecheClusterCenterSelection behind the ABI ------------------------------------------------
 * The reference keeps the adaptive penalty search in R (R/depechePenaltyOpt.R:14-265) and returns to R between its sets, each
 * set a foreach over SOCK workers that receive the whole matrix (:37-38,50-55).  With these entry points a one-process driver
 * runs the whole search, and the selection of the most central solution (R/depecheClusterCenterSelection.R:11-37), without
 * leaving the library between sets.  The state machine itself (create / used / feed / result / solutions) is plain host code
 * and needs no device: a caller that computes the sets elsewhere can drive it step by step. */
typedef struct dpr_penalty_opt* dpr_penalty_opt_t;
/* arguments as depechePenaltyOpt's: penalties (unscaled, e.g. 2^seq(0, 5, by = 0.5)), k, d = ncol(inDataFrameUsed), sampleSize,
 * nCores (worker results per set), maxIter, minARIImprovement, optimARI */
int dpr_penalty_opt_create(const double* penalties, int32_t n_pen, int32_t k, int32_t d, uint32_t sample_size, int32_t n_cores,
                           int32_t max_iter, double min_ari_improvement, double optim_ari, dpr_penalty_opt_t* out);
int dpr_penalty_opt_destroy(dpr_penalty_opt_t opt);
/* the next set: how many penalties are still in use, their positions (0-based) and scaled values penalty * sampleSize *
 * sqrt(d) / 1450 (:22-23), and whether the loop rule (:47-48) asks for another set.  pos_out / reg_out hold n_pen entries. */
int dpr_penalty_opt_used(dpr_penalty_opt_t opt, int32_t* n_used_out, int32_t* pos_out, double* reg_out, int32_t* continue_out);
/* results of one set: n_cores rows of n_used values each (worker-major: the layout of dpr_dataset_grid_search_each for one
 * k), centres (nullable) [worker][penalty][2] matrices k x d col-major.  Applies :75-165 (trivial solutions zeroed, mean /
 * sd per penalty, pruning with the k1 / k2 schedule, stdChange). */
int dpr_penalty_opt_feed(dpr_penalty_opt_t opt, const double* ari_each, const double* nclust_each, const double* centres,
                         int32_t* continue_out);
/* :173-264: best penalty (rounded to one digit like the reference's names) and its position, mean ARI / mean cluster count
 * per penalty (n_pen values each, nullable), worker-iterations run, number of centre matrices stored at the best penalty,
 * warn_flags bit 0: iteration cap reached before stability, bit 1 / 2: the best penalty is the lowest / highest given */
int dpr_penalty_opt_result(dpr_penalty_opt_t opt, double* best_penalty_out, int32_t* best_pos_out, double* mean_ari_out,
                           double* mean_nclust_out, int32_t* worker_iterations_out, int32_t* n_solutions_out,
                           int32_t* warn_flags_out);
/* bestClusterCenters (:255-258): n_solutions matrices k x d col-major, in worker-iteration order, ret1 then ret2 */
int dpr_penalty_opt_solutions(dpr_penalty_opt_t opt, double* centres_out);
/* the whole loop on a resident dataset: while (continue) { used -> one batched set (dpr_dataset_grid_search_each) -> feed }.
 * split_over_ranks != 0 (after dpr_comm_init; every rank holds the same matrix and makes this call): the (worker, penalty)
 * problems of a set are dealt over the ranks and their small results summed into place with one allreduce per set; ARI and
 * cluster counts do not depend on the number of ranks (the stored centres agree to the rounding of their fp64 sums). */
int dpr_dataset_penalty_opt(dpr_dataset_t ds, dpr_penalty_opt_t opt, uint64_t seed, int32_t split_over_ranks);
/* depecheClusterCenterSelection: allocate the selection set to each of the m solutions (zero-sum rows dropped as dAllocate
 * does, R/dAllocate.R:99-103), all-pairs rand_index, first solution with the highest mean ARI.  mean_ari_out: m, nullable. */
int dpr_dataset_select_solution(dpr_dataset_t selection_set, const double* solutions, int32_t m, int32_t k, int32_t* best_out,
                                double* mean_ari_out);

/* ---- device-resident Lloyd iterations (bench: the timed hot loop, nothing crosses PCIe) --------- */
/* Runs `iters` fused iterations (assignment + sums/counts + soft-threshold update) from mu_inout without the
 * convergence exit.  ramp_i >= 0: continue a fit with the penalty ramp frozen at iteration ramp_i (steady state).
 * ramp_i < 0: a fresh fit — previous labels zeroed, iterations numbered 0, 1, 2, ... like find_centers.
 * changed_out (nullable) = labels changed in the last iteration. */
int dpr_dataset_lloyd_iterations(dpr_dataset_t ds, double* mu_inout, uint32_t k, double reg, int32_t no_zero,
                                 int32_t ramp_i, int32_t iters, int64_t* changed_out);
/* same for the plain assignment pass (dAllocate's kernel), labels stay on device */
int dpr_dataset_assign_iterations(dpr_dataset_t ds, const double* mu, uint32_t k, int32_t no_zero,
                                  int32_t iters);

/* ---- multi-GPU (one process per GPU; rows of X sharded; NCCL allreduce of K x (D+1) partials) ------
 * The exchange is opt-in per call: only the row-sharded entry points — dpr_dataset_find_centers_from (every rank: its shard
 * and the same initial centres), dpr_dataset_lloyd_iterations, dpr_dataset_reevaluate_centers — sum their tables over the
 * ranks; everything else (sparse_k_means, grid_search, allocate_*, rand_index) works on the calling rank's data alone whatever
 * the communicator, and dpr_dataset_penalty_opt(split_over_ranks) deals independent problems over the ranks. */
/* unique_id: the 128-byte ncclUniqueId made by rank 0 (dpr_comm_unique_id) and broadcast by the host
 * layer (torch.distributed / MPI / R sockets).  nccl_path: library to dlopen (NULL = "libnccl.so.2"). */
int dpr_comm_unique_id(void* id128, const char* nccl_path);
int dpr_comm_init(int32_t rank, int32_t world, const void* id128, const char* nccl_path);
int dpr_comm_destroy(void);
int dpr_comm_info(int32_t* rank, int32_t* world);
/* Optional, ranks on one node (NVLink): the per-iteration exchange of a row-sharded fit through peer memory instead of
   ncclAllReduce.  Every rank calls dpr_comm_peer_handle (allocates its mailbox, returns its 64-byte CUDA IPC handle), the host
   gathers the handles in rank order, every rank calls dpr_comm_peer_attach.  Results are identical on every rank (each adds the
   ranks' tables in rank order).  Without these calls the NCCL path is used. */
int dpr_comm_peer_handle(void* handle64);
int dpr_comm_peer_attach(const void* handles, int32_t world);
int dpr_comm_peer_detach(void);

#ifdef __cplusplus
}
#endif
#endif /* DEPECHER_B200_H */
