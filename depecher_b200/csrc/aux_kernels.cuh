// aux_kernels.cuh — launchers of aux_kernels.cu
#pragma once
#include <algorithm>

#include "common.cuh"

namespace dpr {

constexpr int kSeedBlock = 256;

struct SeedFit {            // one k-means++ seeding problem (device-resident array)
    const float* x;         // tile-major (x_index)
    int64_t n;
    uint32_t fit_id;        // Philox stream (include/dpr_philox.h)
    uint32_t pad_;
    double* mu;             // k x d row-major, rows filled one per round
    double* w;              // n running min squared distances
    double* block_sums;     // ceil(n / kSeedBlock)
    long long* rows;        // nullable: chosen row numbers (k)
    int* done;              // ticket counter of the round's CTAs (zero between rounds)
};

struct AriPair {
    const int32_t* a;
    const int32_t* b;
    int64_t n;
    uint32_t pair_id;
    uint32_t pad_;
};

int launch_f64_to_cols(const double* src, int64_t e0, int64_t count, int64_t n, int d, float* dst, cudaStream_t st);
int launch_f32_to_cols(const float* src, int64_t e0, int64_t count, int64_t n, int d, float* dst, cudaStream_t st);
int launch_cols_to_f64(const float* x, int64_t row0, int64_t rows, int d, double* out, cudaStream_t st);
int launch_max_finite(const float* v, int64_t n, float* out, cudaStream_t st);
// cell-major copy of a tile-major matrix (row c at c * 4 * tile_groups(d) floats)
int launch_rows_from_tiles(const float* x, int64_t n_tiles, int d, float* rows, cudaStream_t st);
int launch_tile_norms(const float* x, int64_t n_tiles, int d, float* out, cudaStream_t st);
int launch_gather(const float* x, int64_t n, int d, const int64_t* rows, uint64_t seed, uint32_t fit0, int64_t m, float* out, int n_fits,
                  int64_t fit_stride, cudaStream_t st);
int launch_bootstrap_rows(int64_t n, uint32_t boot_n, uint64_t seed, uint32_t fit, int64_t* rows, cudaStream_t st);
int launch_seeding(SeedFit* d_fits, int n_fits, int64_t n_max, int d, int k, uint64_t seed, cudaStream_t st);
int launch_ssq(const float* x, int64_t n, int d, const double* mu, int k, const int32_t* labels, double* block_out, int blocks,
               cudaStream_t st);
int launch_ari(const AriPair* d_pairs, int n_pairs, int64_t n_max, int kk, int mode, uint64_t seed, unsigned long long* d_tables,
               unsigned int* d_stab, int32_t* d_bad, double* d_out, cudaStream_t st);
// exact ARI of n_pairs contingency tables (kk x kk, u64) that were each filled from the same n cells
int launch_ari_from_tables(int n_pairs, int64_t n, int kk, const unsigned long long* d_tables, double* d_out, cudaStream_t st);
int launch_label_counts(const int32_t* labels, int64_t n, int kk, unsigned long long* counts, cudaStream_t st);   // counts[l] = cells with label l
int launch_label_presence(const int32_t* labels, int64_t n, int kk, int32_t* present, int32_t* bad, cudaStream_t st);
int launch_synth(float* x, int64_t n, int d, const float* centres, int groups, uint64_t seed, int64_t row0, double* sum, double* sumsq,
                 cudaStream_t st);
int launch_scale(float* x, int64_t n_pad, int d, float inv_sd, cudaStream_t st);

}  // namespace dpr
