// api_internal.cuh — pieces of api.cu that other translation units of the library use (not part of the C ABI)
#pragma once
#include <initializer_list>
#include <vector>

#include "common.cuh"

namespace dpr {
// The problems [c_lo, c_hi) of optimize_param's (iteration, penalty) grid for ONE k (src/Clusterer.cpp:349-384), numbered
// c = iteration * nreg + penalty like the reference's loops: bootstrap pairs, fits, full-data scoring, ARI.  Random streams are
// addressed by the global problem number, so a range computed on its own equals the same range of the whole grid (centres: to the rounding of their fp64 sums).
// ari_out / nclust_out: c_hi - c_lo values; centers_out (nullable): per problem ret1, ret2 (k x d col-major each).
int grid_search_cells(dpr_dataset* ds, int32_t k, const double* reg, int32_t nreg, uint32_t iterations, uint32_t boot_n, uint64_t seed, int c_lo,
                      int c_hi, double* ari_out, double* nclust_out, double* centers_out);
// element-wise sum over the ranks of host vectors (small results of a grid split; every element is non-zero on one rank only)
int comm_allreduce_host_f64(std::initializer_list<std::vector<double>*> bufs);
}  // namespace dpr
