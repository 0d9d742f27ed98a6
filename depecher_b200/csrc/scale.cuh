// scale.cuh — launchers of scale.cu (device-side depecheLogCenterSd)
#pragma once
#include <algorithm>

#include "common.cuh"

namespace dpr {
int stat_blocks();
int launch_widen(const float* src, double* dst, int64_t count, cudaStream_t st);   // dst[i] = (double)src[i]
int launch_col_sum_minmax(const double* raw, int64_t n, int d, double* psum, double* pmin, double* pmax, cudaStream_t st);
int launch_col_moments(const double* raw, int64_t n, int d, const double* shift, double* p2, double* p4, cudaStream_t st);
int launch_log2(double* raw, int64_t n, int d, const double* colmin, int per_column, cudaStream_t st);
// hist() of every marker in one launch: the breaks are R's pretty() sequence l + i * by (ends exact, values within 1e-14 * by of zero
// snapped to zero, api.cu:r_pretty_breaks) with hist.default's fuzz — recomputed per value on the device with the host's operations
// instead of being looked up in a table (200 000 breaks per marker at 10M cells)
struct HistCol {
    const double* col;      // n values
    unsigned int* counts;   // nb bins (zeroed by the launcher)
    double l, by, u, diddle, inv_by;
    int nb, pad_;
};
int launch_hist_cols(const HistCol* d_cols, int d, int64_t n, unsigned int* counts_all, size_t counts_total, int* argmax_out, cudaStream_t st);
int launch_scale_to_tiles(const double* raw, int64_t n, int d, const double* centre, double sd, float* out, cudaStream_t st);
}  // namespace dpr
