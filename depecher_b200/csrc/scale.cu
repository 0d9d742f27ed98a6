// scale.cu — depecheLogCenterSd on the device: the one other O(N*D) step of depeche().
//
// Restates /root/reference/R/depecheLogCenterSd.R:6-115 and the "peak" centre of R/dScaleCoFunction.R:74-88:
//   1. kurtosis of all values (moments::kurtosis = n*sum((x-m)^4)/sum((x-m)^2)^2) > 100 and !log2Off -> log2 transform
//      (log2(x+1) when min >= 0, else log2(x - colmin + 1));
//   2. per-column centre: mid of the fullest bin of hist(x, breaks = n/50 (10 below 500 rows)) | column mean | given | none;
//   3. division by ONE standard deviation taken over all centred values (sd(as.matrix(.)), n-1 denominator);
// and writes the result as the resident fp32 tile-major dataset.  All statistics are fp64 with fixed-order reductions;
// (x - c)/sd is evaluated in fp64 and rounded to fp32 once, exactly like R's doubles handed to the fp32 upload.
// The histogram breakpoints are R's pretty() sequence, found on the host (api side); the device recomputes a break from the
// sequence's parameters where it needs one and applies the fuzz R applies (hist.default: breaks + 1e-7 * median(diff(breaks)),
// right-closed bins, lowest value included).
#include "scale.cuh"

namespace dpr {

constexpr int kStatBlocks = 256;   // partial results per column

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x == 0)
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += red[q];
    __syncthreads();
    return t;   // valid in thread 0
}

// per column: sum, min, max (NaN ignored by min/max like R's range() would fail on; depeche's input has no NaN)
__global__ void col_sum_minmax_kernel(const double* __restrict__ raw, int64_t n, double* psum, double* pmin, double* pmax) {
    __shared__ double red[32];
    const int j = blockIdx.y;
    const double* col = raw + (int64_t)j * n;
    double s = 0.0, lo = INFINITY, hi = -INFINITY;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const double v = col[i];
        s += v;
        lo = fmin(lo, v);
        hi = fmax(hi, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    __shared__ double rlo[32], rhi[32];
    if ((threadIdx.x & 31) == 0) {
        rlo[threadIdx.x >> 5] = lo;
        rhi[threadIdx.x >> 5] = hi;
    }
    const double t = block_sum(s, red);
    if (threadIdx.x == 0) {
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) {
            lo = fmin(lo, rlo[q]);
            hi = fmax(hi, rhi[q]);
        }
        psum[j * gridDim.x + blockIdx.x] = t;
        pmin[j * gridDim.x + blockIdx.x] = lo;
        pmax[j * gridDim.x + blockIdx.x] = hi;
    }
}

// per column: sum (x - shift_j)^2 and sum (x - shift_j)^4
__global__ void col_moments_kernel(const double* __restrict__ raw, int64_t n, const double* __restrict__ shift, double* p2, double* p4) {
    __shared__ double red[32];
    const int j = blockIdx.y;
    const double* col = raw + (int64_t)j * n;
    const double m = shift[j];
    double s2 = 0.0, s4 = 0.0;
    const int64_t per = (n + gridDim.x - 1) / gridDim.x, i0 = blockIdx.x * per, i1 = min(n, i0 + per);
    for (int64_t i = i0 + threadIdx.x; i < i1; i += blockDim.x) {
        const double v = col[i] - m, q = v * v;
        s2 += q;
        s4 += q * q;
    }
    const double t2 = block_sum(s2, red);
    const double t4 = block_sum(s4, red);
    if (threadIdx.x == 0) {
        p2[j * gridDim.x + blockIdx.x] = t2;
        p4[j * gridDim.x + blockIdx.x] = t4;
    }
}

// R/depecheLogCenterSd.R:21-41
__global__ void log2_kernel(double* raw, int64_t n, int d, const double* __restrict__ colmin, int per_column) {
    const int64_t total = n * d;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int j = (int)(t / n);
        double v = per_column ? log2(raw[t] - colmin[j] + 1.0) : log2(raw[t] + 1.0);
        if (v != v) v = 0.0;
        raw[t] = v;
    }
}

// (x - c_j) / sd in fp64 -> fp32, into the tile-major resident layout
__global__ void scale_to_tiles_kernel(const double* __restrict__ raw, int64_t n, int d, const double* __restrict__ centre, double sd,
                                      float* __restrict__ out) {
    // a thread takes one cell's four markers of a group: four coalesced column reads, one 16-byte store into the tile (a warp writes 512
    // consecutive bytes; one marker per thread wrote 4 bytes of every 16: each sector of the tile in four to eight pieces)
    const int groups = tile_groups(d);
    const int64_t total = n * groups;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t g = t / n, i = t - g * n;
        float v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int j = 4 * (int)g + q;
            v[q] = j < d ? (float)((raw[(int64_t)j * n + i] - centre[j]) / sd) : 0.f;
        }
        *reinterpret_cast<float4*>(out + x_index(i, 4 * (int)g, d)) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

static inline int grid_for(int64_t n) { return (int)std::max<int64_t>(1, std::min<int64_t>(148 * 8, (n + 255) / 256)); }

// fp32 values that are the exact images of the caller's doubles (checked on the host) back to doubles
__global__ void widen_kernel(const float* __restrict__ src, double* __restrict__ dst, int64_t count) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) dst[i] = (double)src[i];
}
int launch_widen(const float* src, double* dst, int64_t count, cudaStream_t st) {
    widen_kernel<<<(int)std::min<int64_t>((count + 255) / 256, 148 * 16), 256, 0, st>>>(src, dst, count);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_col_sum_minmax(const double* raw, int64_t n, int d, double* psum, double* pmin, double* pmax, cudaStream_t st) {
    col_sum_minmax_kernel<<<dim3(kStatBlocks, d), 256, 0, st>>>(raw, n, psum, pmin, pmax);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_col_moments(const double* raw, int64_t n, int d, const double* shift, double* p2, double* p4, cudaStream_t st) {
    col_moments_kernel<<<dim3(kStatBlocks, d), 256, 0, st>>>(raw, n, shift, p2, p4);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_log2(double* raw, int64_t n, int d, const double* colmin, int per_column, cudaStream_t st) {
    log2_kernel<<<grid_for(n * d), 256, 0, st>>>(raw, n, d, colmin, per_column);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
// hist.default + C_BinCount (right-closed bins, lowest value included, hist.default's fuzz) for every marker
__device__ __forceinline__ double hist_fuzzy_break(const HistCol& c, int i) {
    double v = i == 0 ? c.l : (i == c.nb ? c.u : __dadd_rn(c.l, __dmul_rn((double)i, c.by)));   // (separate multiply and add: what the host computes)
    if (fabs(v) < 1e-14 * c.by) v = 0.0;
    return i == 0 ? v - c.diddle : v + c.diddle;
}
__global__ void hist_cols_kernel(const HistCol* __restrict__ cols, int64_t n) {
    const HistCol c = cols[blockIdx.y];
    const double f0 = hist_fuzzy_break(c, 0), f1 = hist_fuzzy_break(c, c.nb);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double x = c.col[i];
        if (!(f0 <= x && x <= f1)) continue;                 // outside the (fuzzy) range: not counted, like BinCount
        int b = (int)((x - c.l) * c.inv_by);
        b = b < 0 ? 0 : (b > c.nb - 1 ? c.nb - 1 : b);
        while (b < c.nb - 1 && x > hist_fuzzy_break(c, b + 1)) ++b;   // right-closed bins: fb[b] < x <= fb[b+1]
        while (b > 0 && !(x > hist_fuzzy_break(c, b))) --b;
        atomicAdd(&c.counts[b], 1u);                           // integer counts: order-independent
    }
}
__global__ void argmax_cols_kernel(const HistCol* __restrict__ cols, int* out) {   // match(max(counts), counts): first maximum; one CTA per marker
    const HistCol c = cols[blockIdx.x];
    __shared__ unsigned int sv[1024];
    __shared__ int si[1024];
    unsigned int bv = 0;
    int bi = 0x7fffffff;
    for (int i = threadIdx.x; i < c.nb; i += blockDim.x) {
        const unsigned int v = c.counts[i];
        if (v > bv || (v == bv && i < bi)) {
            bv = v;
            bi = i;
        }
    }
    sv[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) {
            const unsigned int ov = sv[threadIdx.x + o];
            const int oi = si[threadIdx.x + o];
            if (ov > sv[threadIdx.x] || (ov == sv[threadIdx.x] && oi < si[threadIdx.x])) {
                sv[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) out[blockIdx.x] = si[0];
}
int launch_hist_cols(const HistCol* d_cols, int d, int64_t n, unsigned int* counts_all, size_t counts_total, int* argmax_out, cudaStream_t st) {
    DPR_CUDA(cudaMemsetAsync(counts_all, 0, sizeof(unsigned int) * counts_total, st));
    const int gx = (int)std::max<int64_t>(1, std::min<int64_t>((148 * 16 + d - 1) / d, (n + 255) / 256));
    hist_cols_kernel<<<dim3(gx, d), 256, 0, st>>>(d_cols, n);
    argmax_cols_kernel<<<d, 1024, 0, st>>>(d_cols, argmax_out);
    count_launch(2);
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_scale_to_tiles(const double* raw, int64_t n, int d, const double* centre, double sd, float* out, cudaStream_t st) {
    scale_to_tiles_kernel<<<grid_for(n * tile_groups(d)), 256, 0, st>>>(raw, n, d, centre, sd, out);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int stat_blocks() { return kStatBlocks; }

}  // namespace dpr
