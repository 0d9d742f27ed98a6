// scale.cuh — launchers of scale.cu (device-side depecheLogCenterSd)
#pragma once
#include <algorithm>

#include "common.cuh"

namespace dpr {
int stat_blocks();
int launch_col_sum_minmax(const double* raw, int64_t n, int d, double* psum, double* pmin, double* pmax, cudaStream_t st);
int launch_col_moments(const double* raw, int64_t n, int d, const double* shift, double* p2, double* p4, cudaStream_t st);
int launch_log2(double* raw, int64_t n, int d, const double* colmin, int per_column, cudaStream_t st);
int launch_hist(const double* col, int64_t n, const double* fuzzy_breaks, int nb, double lo, double inv_by, unsigned int* counts,
                int* argmax_out, cudaStream_t st);
int launch_scale_to_tiles(const double* raw, int64_t n, int d, const double* centre, double sd, float* out, cudaStream_t st);
}  // namespace dpr
