// centres.cuh — launchers of centres.cu
#pragma once
#include "common.cuh"

namespace dpr {
int launch_prepare_centres(Problem* d_problems, int n_problems, int table_cap, cudaStream_t st);
int launch_reduce_partials(const Problem* d_problems, int n_problems, const int32_t* cta_tile_start, int grid_ctas, int k_max, int d,
                           const double* cta_tables, double* seg_out, int n_seg, int32_t* n_active_out, cudaStream_t st);
int launch_finalize(Problem* d_problems, int n_problems, const double* seg_in, int n_seg, int k_max, int d, int iter, int max_iter,
                    int table_cap, int allow_delta, int32_t* n_active_out, cudaStream_t st);
// both of the above in one launch (a cluster of 8 CTAs per problem); with several ranks and attached peer mailboxes (comm.cuh) the tables
// of all ranks are exchanged inside it over NVLink.  n_active2: two counters, iteration i uses [i & 1]
int launch_reduce_finalize(Problem* d_problems, int n_problems, const int32_t* cta_tile_start, int grid_ctas, const double* cta_tables, int k_max, int d,
                           int iter, int max_iter, int table_cap, int allow_delta, int32_t* n_active2, bool exchange, cudaStream_t st);
// the same for batches of many problems: one CTA per problem sums the tables of the few pass CTAs that touch it, no cluster (single GPU)
int launch_finalize_direct(Problem* d_problems, int n_problems, const int32_t* cta_tile_start, int grid_ctas, const double* cta_tables, int k_max, int d,
                           int iter, int max_iter, int table_cap, int allow_delta, int32_t* n_active2, cudaStream_t st);
bool reduce_finalize_fits(int k_max, int d);
bool reduce_finalize_peer_fits(int n_problems, int k_max, int d);
int launch_partition(const Problem* d_problems, int n_problems, int grid_ctas, int total_tiles, int32_t* cta_tile_start, cudaStream_t st);
int launch_soft_threshold(const double* sums, const long long* counts, int k, int d, double reg, double* mu, cudaStream_t st);
}  // namespace dpr
