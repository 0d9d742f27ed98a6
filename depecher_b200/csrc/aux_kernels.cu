// aux_kernels.cu — the O(N) helpers around the fused pass: layout conversion, bootstrap gather,
// k-means++ seeding, cluster_norm, contingency tables / adjusted Rand index, synthetic data.
// Each kernel names the reference lines it replaces.
#include "aux_kernels.cuh"

#include "../../include/dpr_philox.h"

namespace dpr {

// ---------------------------------------------------------------------------------------------------
// R doubles (column-major n x d, flat range [e0, e0+count)) -> resident fp32 columns with leading
// dimension ld.  Replaces numeric_to_eigen (src/InterfaceUtils.cpp:23-33), which transposes; here R's
// layout is already the one the kernels want.
// ---------------------------------------------------------------------------------------------------
__global__ void f64_to_f32_cols_kernel(const double* __restrict__ src, int64_t e0, int64_t count, int64_t n, int d,
                                       float* __restrict__ dst) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = e0 + t;
        const int64_t j = e / n, i = e - j * n;
        dst[x_index(i, (int)j, d)] = (float)src[t];
    }
}
__global__ void f32_to_f32_cols_kernel(const float* __restrict__ src, int64_t e0, int64_t count, int64_t n, int d,
                                       float* __restrict__ dst) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < count; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = e0 + t;
        const int64_t j = e / n, i = e - j * n;
        dst[x_index(i, (int)j, d)] = src[t];
    }
}
__global__ void cols_to_f64_kernel(const float* __restrict__ x, int64_t row0, int64_t rows, int d, double* __restrict__ out) {
    const int64_t total = rows * d;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t j = t / rows, i = t - j * rows;
        out[t] = (double)x[x_index(row0 + i, (int)j, d)];
    }
}

// ---------------------------------------------------------------------------------------------------
// per-tile upper bound of ||x||^2 (one warp per tile).  The fused pass needs ||x||^2 only inside the proven error
// bound that decides which cells are re-examined in fp64; a bound shared by the 128 cells of a tile is as safe and
// saves two FMAs per marker and lane in the hot loop.
// ---------------------------------------------------------------------------------------------------
__global__ void tile_norms_kernel(const float* __restrict__ x, int64_t n_tiles, int d, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int groups = tile_groups(d);
    for (int64_t t = warp; t < n_tiles; t += nwarps) {
        const float4* tile = reinterpret_cast<const float4*>(x + t * (int64_t)tile_floats(d));
        float a = 0.f, b = 0.f, c = 0.f, e = 0.f;   // cells lane, lane+32, lane+64, lane+96
        for (int g = 0; g < groups; ++g) {
            const float4* grp = tile + g * (kGroupFloats / 4);
            const float4 v0 = grp[lane], v1 = grp[lane + 32], v2 = grp[lane + 64], v3 = grp[lane + 96];
            a = fmaf(v0.x, v0.x, fmaf(v0.y, v0.y, fmaf(v0.z, v0.z, fmaf(v0.w, v0.w, a))));
            b = fmaf(v1.x, v1.x, fmaf(v1.y, v1.y, fmaf(v1.z, v1.z, fmaf(v1.w, v1.w, b))));
            c = fmaf(v2.x, v2.x, fmaf(v2.y, v2.y, fmaf(v2.z, v2.z, fmaf(v2.w, v2.w, c))));
            e = fmaf(v3.x, v3.x, fmaf(v3.y, v3.y, fmaf(v3.z, v3.z, fmaf(v3.w, v3.w, e))));
        }
        const bool bad = (a != a) || (b != b) || (c != c) || (e != e);
        float m = fmaxf(fmaxf(a, b), fmaxf(c, e));
        if (bad) m = INFINITY;                                   // a NaN cell: every cell of the tile is decided exactly
        for (int o = 16; o > 0; o >>= 1) {
            const float w = __shfl_xor_sync(0xffffffffu, m, o);
            m = w > m ? w : m;
        }
        if (lane == 0) out[t] = m * 1.00001f;                    // covers the fp32 rounding of the squared norm itself
    }
}

// largest finite value of v[0..n) (0 when there is none): the scale the tensor-core pass sizes its fp16 split by
__global__ void max_finite_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
    __shared__ float red[32];
    float m = 0.f;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
        const float t = v[i];
        if (t < INFINITY && t > m) m = t;
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) m = fmaxf(m, red[q]);
        *out = m;
    }
}

// ---------------------------------------------------------------------------------------------------
// bootstrap_data (src/Clusterer.cpp:308-321): rows drawn with replacement, gathered into a new matrix.
// rows_explicit == nullptr -> row = Philox(seed, BOOT, fit, i) % n   (include/dpr_philox.h)
// ---------------------------------------------------------------------------------------------------
__global__ void gather_rows_kernel(const float* __restrict__ x, int64_t n, int d, const int64_t* __restrict__ rows_explicit,
                                   uint64_t seed, uint32_t fit0, int64_t m, float* __restrict__ out, int n_fits,
                                   int64_t fit_stride) {
    const int64_t total = m * n_fits;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int f = (int)(t / m);
        const int64_t i = t - (int64_t)f * m;
        int64_t row;
        if (rows_explicit) row = rows_explicit[t];
        else row = (int64_t)(dpr_rng_draw(seed, DPR_RNG_BOOT, fit0 + (uint32_t)f, (uint32_t)i, 0).w[0] % (uint64_t)n);
        float* o = out + (int64_t)f * fit_stride;
        for (int j = 0; j < d; ++j) o[x_index(i, j, d)] = x[x_index(row, j, d)];
    }
}
__global__ void bootstrap_rows_kernel(int64_t n, uint32_t boot_n, uint64_t seed, uint32_t fit, int64_t* rows) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < boot_n; i += gridDim.x * blockDim.x)
        rows[i] = (int64_t)(dpr_rng_draw(seed, DPR_RNG_BOOT, fit, i, 0).w[0] % (uint64_t)n);
}

// ---------------------------------------------------------------------------------------------------
// k-means++ seeding: initialize_mu + element_from_vector (src/Clusterer.cpp:123-176), batched over fits.
// Round c (c = 1..k-1):  seed_dist   w_i = min(w_i, ||(-x_i) + mu_{c-1}||^2) in fp64, same operation
//                                    order as the reference (weights are bit-identical), + block sums
//                        seed_pick   x = u * total, first i whose running sum exceeds x (upper_bound of
//                                    the cumulative map, :139-140; zero weights are skipped, :130)
// ---------------------------------------------------------------------------------------------------
__global__ void seed_first_kernel(SeedFit* fits, int d, uint64_t seed) {
    SeedFit f = fits[blockIdx.x];
    const int64_t row = (int64_t)(dpr_rng_draw(seed, DPR_RNG_SEED, f.fit_id, 0, 0).w[0] % (uint64_t)f.n);
    for (int j = threadIdx.x; j < d; j += blockDim.x) f.mu[j] = (double)f.x[x_index(row, j, d)];
    if (threadIdx.x == 0 && f.rows) f.rows[0] = row;
}

__device__ void seed_pick_body(const SeedFit& f, int d, int c, uint64_t seed);
// Round c: distances to centre c - 1 (the running minimum, block sums), then — by the CTA of the fit that finishes last (a ticket
// counter per fit) — the draw of centre c.  One launch per round instead of two: the rounds of a batch of small fits are
// launch-latency-bound (28 launches fewer per seeding).
template <int U>   // blocks of 256 cells per trip of a CTA
__global__ void seed_dist_kernel(const SeedFit* fits, int d, int c /* centre just chosen: c-1 >= 0 */, int first, uint64_t seed) {
    extern __shared__ double s_mu[];
    DPR_GRID_DEPENDENCY();   // (a round reads the centre the previous round's launch drew: the rounds are launched as dependent kernels)
    const SeedFit f = fits[blockIdx.y];
    const int nb = (int)((f.n + kSeedBlock - 1) / kSeedBlock);
    for (int j = threadIdx.x; j < d; j += blockDim.x) s_mu[j] = f.mu[(size_t)(c - 1) * d + j];
    __syncthreads();
    // U blocks of 256 cells per trip: a thread holds one cell of each, the marker groups of the U cells are requested together and their
    // sums advance side by side (U independent fp64 chains), and the CTA meets once per U blocks instead of once per block.  Per cell the
    // operations and their order are unchanged, and a block's sum is still its eight warp sums added in warp order.  (U = 4 for large
    // matrices: a round over 10M x 40 cells 0.38 -> 0.35 ms; batches of small fits keep U = 1: they need the CTAs more than the overlap.)
    __shared__ double s_red[U][kSeedBlock / 32];
    for (int b0 = blockIdx.x * U; b0 < nb; b0 += gridDim.x * U) {
        int64_t i[U];
        bool on[U];
        double s[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            i[u] = (int64_t)(b0 + u) * kSeedBlock + threadIdx.x;
            on[u] = b0 + u < nb && i[u] < f.n;
            if (!on[u]) i[u] = 0;   // (loads stay unconditional: cell 0 exists)
            s[u] = 0.0;
        }
        double old[U];
        if (!first) {
#pragma unroll
            for (int u = 0; u < U; ++u) old[u] = on[u] ? f.w[i[u]] : 0.0;
        }
        // a cell's markers come four at a time (one 16-byte load per marker group); differences and squares of a group are independent,
        // only the additions into s form the left-to-right chain of the reference
        int j = 0;
        for (; j + 4 <= d; j += 4) {
            float4 xv[U];
#pragma unroll
            for (int u = 0; u < U; ++u) xv[u] = __ldg(reinterpret_cast<const float4*>(f.x + x_index(i[u], 0, d)) + (size_t)(j >> 2) * (kGroupFloats / 4));
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double d0 = __dadd_rn(-(double)xv[u].x, s_mu[j]), d1 = __dadd_rn(-(double)xv[u].y, s_mu[j + 1]);
                const double d2 = __dadd_rn(-(double)xv[u].z, s_mu[j + 2]), d3 = __dadd_rn(-(double)xv[u].w, s_mu[j + 3]);
                const double q0 = __dmul_rn(d0, d0), q1 = __dmul_rn(d1, d1), q2 = __dmul_rn(d2, d2), q3 = __dmul_rn(d3, d3);
                s[u] = __dadd_rn(s[u], q0);
                s[u] = __dadd_rn(s[u], q1);
                s[u] = __dadd_rn(s[u], q2);
                s[u] = __dadd_rn(s[u], q3);
            }
        }
        for (; j < d; ++j) {
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const double diff = __dadd_rn(-(double)f.x[x_index(i[u], j, d)], s_mu[j]);
                s[u] = __dadd_rn(s[u], __dmul_rn(diff, diff));
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            double w = 0.0;
            if (on[u]) {
                if (!first && !(s[u] < old[u])) s[u] = old[u];   // if (temp_dists(j) < dists(j)) dists(j) = temp_dists(j)   (:167-169)
                f.w[i[u]] = s[u];
                w = s[u] > 0 ? s[u] : 0.0;
            }
            // deterministic block sum
            for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
            if ((threadIdx.x & 31) == 0) s_red[u][threadIdx.x >> 5] = w;
        }
        __syncthreads();
        if (threadIdx.x < U && b0 + (int)threadIdx.x < nb) {
            double t = 0.0;
            for (int q = 0; q < kSeedBlock / 32; ++q) t += s_red[threadIdx.x][q];
            f.block_sums[b0 + threadIdx.x] = t;
        }
        __syncthreads();
    }
    __shared__ int s_last;
    __threadfence();   // this CTA's weights and block sums are visible before its ticket is
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(f.done, 1) == (int)gridDim.x - 1;
    __syncthreads();
    if (!s_last) return;
    if (threadIdx.x == 0) *f.done = 0;   // (for the next round)
    __threadfence();
    seed_pick_body(f, d, c, seed);
}

// One CTA per fit.  The cumulative sum the reference keeps in a std::map is evaluated hierarchically:
// thread segments of block sums -> the segment holding x -> its blocks -> the cells of one block.  Every level
// is summed in ascending order, so the pick only differs from a purely sequential cumsum when x falls within
// rounding distance (~1e-16 relative) of a boundary.
__device__ void seed_pick_body(const SeedFit& f, int d, int c, uint64_t seed) {
    // The scans are sequential by definition (a running sum with an early exit), and one thread does them — but never on global
    // memory: every stretch it walks (the block sums of a segment, the weights of a block) is first brought into shared memory by the
    // whole CTA, and the thread reads it four values at a time.  (Walking global memory with a data-dependent exit costs a DRAM/L2 round
    // trip per element: 40-80 us per round, which was most of a seeding round for batches of small fits and a third of it for 10M cells.)
    const int nb = (int)((f.n + kSeedBlock - 1) / kSeedBlock);
    const int T = blockDim.x;   // == kSeedBlock == 256
    const int seg = (nb + T - 1) / T;
    __shared__ __align__(16) double s_seg[256];
    __shared__ __align__(16) double s_buf[256];
    __shared__ long long s_row;
    __shared__ double s_x, s_cum;
    __shared__ int s_sg, s_blk, s_last_blk, s_found;
    {
        double t = 0.0;
        const int b0 = threadIdx.x * seg, b1 = min(nb, b0 + seg);
        for (int b = b0; b < b1; b += 16) {   // sixteen block sums requested at once, added in ascending order (a missing one adds +0)
            double v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = b + q < b1 ? __ldcg(f.block_sums + b + q) : 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q) t += v[q];
        }
        s_seg[threadIdx.x] = t;
    }
    __syncthreads();
    // first element of v[0..n) at which the running sum (started at cum) exceeds x, skipping zero weights; cum is left at the sum before
    // that element (or after all n); last_pos: the last positive element seen.  Same additions in the same order as a plain loop.
    auto scan = [](const double* v, int n, double x, double& cum, int& last_pos) {
        for (int t0 = 0; t0 < n; t0 += 4) {
            const double2 a = *reinterpret_cast<const double2*>(v + t0), b = *reinterpret_cast<const double2*>(v + t0 + 2);
            const double w[4] = {a.x, a.y, b.x, b.y};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if (t0 + q >= n) return -1;
                if (w[q] > 0) {
                    last_pos = t0 + q;
                    const double nxt = cum + w[q];
                    if (nxt > x) return t0 + q;
                    cum = nxt;
                }
            }
        }
        return -1;
    };
    if (threadIdx.x == 0) {
        double total = 0.0;
        for (int t = 0; t < T; ++t) total += s_seg[t];
        const double u = dpr_rng_unit(dpr_rng_draw(seed, DPR_RNG_SEED, f.fit_id, (uint32_t)c, 0).w[0]);
        const double x = u * total;
        double cum = 0.0;
        int lp = -1;
        s_sg = scan(s_seg, T, x, cum, lp);
        s_x = x;
        s_cum = cum;
        s_blk = -1;
        s_last_blk = -1;
        s_found = 0;
        s_row = 0;
    }
    __syncthreads();
    const int sg = s_sg;
    if (sg >= 0) {   // (uniform)
        // the blocks of the segment, 256 at a time
        const int b0 = sg * seg, b1 = min(nb, b0 + seg);
        for (int base = b0; base < b1; base += T) {
            s_buf[threadIdx.x] = base + (int)threadIdx.x < b1 ? f.block_sums[base + threadIdx.x] : 0.0;
            __syncthreads();
            if (threadIdx.x == 0) {
                double cum = s_cum;
                int lp = -1;
                const int at = scan(s_buf, min(T, b1 - base), s_x, cum, lp);
                s_cum = cum;
                if (lp >= 0) s_last_blk = base + lp;
                if (at >= 0) {
                    s_blk = base + at;
                    s_found = 1;
                }
            }
            __syncthreads();
            if (s_found) break;
        }
        const int blk = s_blk >= 0 ? s_blk : s_last_blk;   // rounding of the segment sums: take the segment's last non-empty block
        // the cells of the block
        const int64_t i0 = (int64_t)blk * kSeedBlock, i1 = min(f.n, i0 + (int64_t)kSeedBlock);
        s_buf[threadIdx.x] = i0 + threadIdx.x < i1 ? f.w[i0 + threadIdx.x] : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            double run = s_cum;
            int lp = -1;
            const int at = scan(s_buf, (int)(i1 - i0), s_x, run, lp);
            s_row = at >= 0 ? i0 + at : (lp < 0 ? i0 : i0 + lp);
        }
        __syncthreads();
    }
    const long long row = s_row;
    if (threadIdx.x == 0 && f.rows) f.rows[c] = row;
    for (int j = threadIdx.x; j < d; j += blockDim.x) f.mu[(size_t)c * d + j] = (double)f.x[x_index(row, j, d)];
}

// ---------------------------------------------------------------------------------------------------
// cluster_norm (src/Clusterer.cpp:198-211): sum_i ||x_i - mu_{l_i}||^2  (+ reg * sum(mu), added by the host)
// ---------------------------------------------------------------------------------------------------
__global__ void ssq_kernel(const float* __restrict__ x, int64_t n, int d, const double* __restrict__ mu, int k,
                           const int32_t* __restrict__ labels, double* __restrict__ block_out) {
    extern __shared__ double s_mu[];
    for (int e = threadIdx.x; e < k * d; e += blockDim.x) s_mu[e] = mu[e];
    __syncthreads();
    double acc = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const double* m = s_mu + (size_t)labels[i] * d;
        double s = 0.0;
        for (int j = 0; j < d; ++j) {
            const double diff = (double)x[x_index(i, j, d)] - m[j];
            s += diff * diff;
        }
        acc += s;
    }
    __shared__ double s_red[32];
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < (int)(blockDim.x >> 5); ++q) t += s_red[q];
        block_out[blockIdx.x] = t;
    }
}

// ---------------------------------------------------------------------------------------------------
// adjusted Rand index.  Exact: contingency table (warp-aggregated shared-memory histogram) -> the value the
// reference's 10 000-pair Monte-Carlo estimate converges to (src/Clusterer.cpp:267-305).  MC: the estimate
// itself with Philox pairs.  Labels in [0, k]  (tables are (k+1) wide, SURVEY A-9).
// ---------------------------------------------------------------------------------------------------
__global__ void contingency_kernel(const AriPair* pairs, int kk /* k+1 */, unsigned long long* tables, int32_t* bad) {
    extern __shared__ unsigned int s_tab[];
    const AriPair pr = pairs[blockIdx.y];
    const int cells = kk * kk;
    for (int e = threadIdx.x; e < cells; e += blockDim.x) s_tab[e] = 0u;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t per_block = ((pr.n + gridDim.x - 1) / gridDim.x + 31) & ~(int64_t)31;
    const int64_t i0 = blockIdx.x * per_block, i1 = min(pr.n, i0 + per_block);
    for (int64_t base = i0 + (threadIdx.x & ~31); base < i1; base += blockDim.x) {
        const int64_t i = base + lane;
        int key = -1;
        if (i < i1) {
            const int a = pr.a[i], b = pr.b[i];
            if (a < 0 || b < 0 || a >= kk || b >= kk) *bad = 1;
            else key = a * kk + b;
        }
        const unsigned m = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && lane == __ffs(m) - 1) atomicAdd(&s_tab[key], (unsigned)__popc(m));
    }
    __syncthreads();
    unsigned long long* out = tables + (size_t)blockIdx.y * cells;
    for (int e = threadIdx.x; e < cells; e += blockDim.x)
        if (s_tab[e]) atomicAdd(&out[e], (unsigned long long)s_tab[e]);
}

// same value from a table that is too wide for shared memory (rare: k > ~230)
__global__ void contingency_global_kernel(const AriPair* pairs, int kk, unsigned long long* tables, int32_t* bad) {
    const AriPair pr = pairs[blockIdx.y];
    unsigned long long* out = tables + (size_t)blockIdx.y * kk * kk;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < pr.n; i += (int64_t)gridDim.x * blockDim.x) {
        const int a = pr.a[i], b = pr.b[i];
        if (a < 0 || b < 0 || a >= kk || b >= kk) *bad = 1;
        else atomicAdd(&out[(size_t)a * kk + b], 1ULL);
    }
}

__global__ void ari_from_tables_kernel(const AriPair* pairs, int kk, const unsigned long long* tables, double* out) {
    const int p = blockIdx.x;
    if (threadIdx.x != 0) return;
    const unsigned long long* t = tables + (size_t)p * kk * kk;
    const double N = (double)pairs[p].n, denom = N * (N - 1.0);
    double same12 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int a = 0; a < kk; ++a) {
        unsigned long long r = 0;
        for (int b = 0; b < kk; ++b) {
            const unsigned long long v = t[(size_t)a * kk + b];
            r += v;
            same12 += (double)v * (double)(v ? v - 1 : 0);
        }
        s1 += (double)r * (double)(r ? r - 1 : 0);
    }
    for (int b = 0; b < kk; ++b) {
        unsigned long long c = 0;
        for (int a = 0; a < kk; ++a) c += t[(size_t)a * kk + b];
        s2 += (double)c * (double)(c ? c - 1 : 0);
    }
    const double A = s1 / denom, B = s2 / denom, both = same12 / denom;
    const double ri = both + (1.0 - A - B + both);
    const double e = A * B + (1.0 - A) * (1.0 - B);
    out[p] = (ri - e) / (1.0 - e);
}

// the same for tables that every pair fills from the same n cells (the scoring pass counted them: score_pass.cu)
__global__ void ari_from_tables_n_kernel(long long n, int kk, const unsigned long long* tables, double* out) {
    const int p = blockIdx.x;
    const unsigned long long* t = tables + (size_t)p * kk * kk;
    double same12 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int a = threadIdx.x; a < kk; a += blockDim.x) {   // thread a: row a and column a (integer sums: any order gives the same value)
        unsigned long long r = 0, c = 0;
        for (int b = 0; b < kk; ++b) {
            const unsigned long long v = t[(size_t)a * kk + b];
            r += v;
            c += t[(size_t)b * kk + a];
            same12 += (double)v * (double)(v ? v - 1 : 0);
        }
        s1 += (double)r * (double)(r ? r - 1 : 0);
        s2 += (double)c * (double)(c ? c - 1 : 0);
    }
    // products of integers below 2^53 are exact in fp64 only up to 2^53: add them in a fixed order (lane order, then warp order)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        same12 += __shfl_down_sync(0xffffffffu, same12, o);
        s1 += __shfl_down_sync(0xffffffffu, s1, o);
        s2 += __shfl_down_sync(0xffffffffu, s2, o);
    }
    __shared__ double s_w[3][8];
    if ((threadIdx.x & 31) == 0) {
        s_w[0][threadIdx.x >> 5] = same12;
        s_w[1][threadIdx.x >> 5] = s1;
        s_w[2][threadIdx.x >> 5] = s2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double q12 = 0.0, q1 = 0.0, q2 = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
            q12 += s_w[0][w];
            q1 += s_w[1][w];
            q2 += s_w[2][w];
        }
        const double N = (double)n, denom = N * (N - 1.0);
        const double A = q1 / denom, B = q2 / denom, both = q12 / denom;
        const double ri = both + (1.0 - A - B + both);
        const double e = A * B + (1.0 - A) * (1.0 - B);
        out[p] = (ri - e) / (1.0 - e);
    }
}

// Monte-Carlo estimate: pair `num` is (w0 % n, w1 % n) of Philox(seed, PAIR, pair_id, num, attempt), redrawn
// while i == j (:291-295); `stab` counts agreeing pairs (:296-300).
__global__ void ari_mc_count_kernel(const AriPair* pairs, uint64_t seed, unsigned int* stab, int32_t* bad, int kk) {
    const AriPair pr = pairs[blockIdx.y];
    const uint32_t num = blockIdx.x * blockDim.x + threadIdx.x;
    int agree = 0;
    if (num < 10000u) {
        uint64_t i = 0, j = 0;
        uint32_t attempt = 0;
        while (i == j) {
            const dpr_u32x4 w = dpr_rng_draw(seed, DPR_RNG_PAIR, pr.pair_id, num, attempt++);
            i = w.w[0] % (uint64_t)pr.n;
            j = w.w[1] % (uint64_t)pr.n;
        }
        const int a1 = pr.a[i], a2 = pr.a[j], b1 = pr.b[i], b2 = pr.b[j];
        if (a1 < 0 || a2 < 0 || b1 < 0 || b2 < 0 || a1 >= kk || a2 >= kk || b1 >= kk || b2 >= kk) *bad = 1;
        agree = ((a1 == a2) && (b1 == b2)) || ((a1 != a2) && (b1 != b2));
    }
    const unsigned m = __ballot_sync(0xffffffffu, agree);
    if ((threadIdx.x & 31) == 0 && m) atomicAdd(&stab[blockIdx.y], (unsigned)__popc(m));
}
__global__ void ari_mc_final_kernel(const AriPair* pairs, int kk, const unsigned long long* tables, const unsigned int* stab, double* out) {
    const int p = blockIdx.x;
    if (threadIdx.x != 0) return;
    const unsigned long long* t = tables + (size_t)p * kk * kk;
    const double size = (double)pairs[p].n;
    double a1 = 0, a2 = 0, b1 = 0, b2 = 0;
    for (int c = 0; c < kk; ++c) {
        unsigned long long r = 0, q = 0;
        for (int b = 0; b < kk; ++b) { r += t[(size_t)c * kk + b]; q += t[(size_t)b * kk + c]; }
        const double p1 = (double)r / size, p2 = (double)q / size;
        a1 += p1 * (p1 - 1.0 / size) * (size / (size - 1));
        a2 += p2 * (p2 - 1.0 / size) * (size / (size - 1));
        b1 += p1 * (1 - (p1 - 1.0 / size) * (size / (size - 1)));
        b2 += p2 * (1 - (p2 - 1.0 / size) * (size / (size - 1)));
    }
    const double f = a1 * a2 + b1 * b2;
    out[p] = (((double)stab[p] / 10000.0) - f) / (1 - f);
}

// n_used_clusters (src/Clusterer.cpp:323-332): distinct labels
__global__ void label_presence_kernel(const int32_t* labels, int64_t n, int kk, int32_t* present, int32_t* bad) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = labels[i];
        if (l < 0 || l >= kk) *bad = 1;
        else if (!present[l]) present[l] = 1;
    }
}

// cluster sizes: counts[l] = number of cells with label l (0 <= l < kk; kk small: per-CTA counts in shared memory first)
__global__ void label_counts_kernel(const int32_t* __restrict__ labels, int64_t n, int kk, unsigned long long* counts) {
    extern __shared__ unsigned int s_cnt[];
    for (int i = threadIdx.x; i < kk; i += blockDim.x) s_cnt[i] = 0u;
    __syncthreads();
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int l = labels[i];
        if (l >= 0 && l < kk) atomicAdd(&s_cnt[l], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kk; i += blockDim.x)
        if (s_cnt[i]) atomicAdd(&counts[i], (unsigned long long)s_cnt[i]);
}

// ---------------------------------------------------------------------------------------------------
// synthetic cells (bench): group g = Philox % G, x = centre[g][j] + N(0,1)  (generateSparseData shape,
// R/simulateData.R:47-80, generalised), then / global sd (depecheLogCenterSd.R:102-106) by scale_kernel.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float2 box_muller(uint32_t a, uint32_t b) {
    const float u1 = ((float)a + 1.0f) * 2.3283064e-10f, u2 = (float)b * 2.3283064e-10f;
    const float r = sqrtf(-2.0f * logf(u1));
    float s, c;
    sincosf(6.2831853f * u2, &s, &c);
    return make_float2(r * c, r * s);
}
__global__ void synth_kernel(float* x, int64_t n, int d, const float* __restrict__ centres, int groups, uint64_t seed,
                             int64_t row0, double* sum_out, double* sumsq_out) {
    double s = 0.0, ss = 0.0;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t gi = row0 + i;  // global cell number: shards of one virtual matrix never repeat a cell
        const int g = (int)(dpr_rng_draw(seed, 11, (uint32_t)gi, (uint32_t)(gi >> 32), 0).w[0] % (uint32_t)groups);
        for (int j0 = 0; j0 < d; j0 += 4) {
            const dpr_u32x4 w = dpr_rng_draw(seed, 12, (uint32_t)gi, (uint32_t)(gi >> 32), (uint32_t)j0);
            const float2 n0 = box_muller(w.w[0], w.w[1]), n1 = box_muller(w.w[2], w.w[3]);
            const float z[4] = {n0.x, n0.y, n1.x, n1.y};
            for (int q = 0; q < 4 && j0 + q < d; ++q) {
                const float v = centres[g * d + j0 + q] + z[q];
                x[x_index(i, j0 + q, d)] = v;
                s += v;
                ss += (double)v * v;
            }
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        ss += __shfl_xor_sync(0xffffffffu, ss, o);
    }
    if ((threadIdx.x & 31) == 0) {
        atomicAdd(sum_out, s);
        atomicAdd(sumsq_out, ss);
    }
}
__global__ void scale_kernel(float* x, int64_t n_pad, int d, float inv_sd) {
    const int64_t total = (n_pad / kSlotCells) * (int64_t)tile_floats(d);   // padding cells / markers are zero and stay zero
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) x[t] *= inv_sd;
}

// ---------------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------------
static inline int blocks_for(int64_t n, int per = 256, int cap = 148 * 16) {
    int64_t b = (n + per - 1) / per;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

int launch_f64_to_cols(const double* src, int64_t e0, int64_t count, int64_t n, int d, float* dst, cudaStream_t st) {
    f64_to_f32_cols_kernel<<<blocks_for(count), 256, 0, st>>>(src, e0, count, n, d, dst);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_f32_to_cols(const float* src, int64_t e0, int64_t count, int64_t n, int d, float* dst, cudaStream_t st) {
    f32_to_f32_cols_kernel<<<blocks_for(count), 256, 0, st>>>(src, e0, count, n, d, dst);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_cols_to_f64(const float* x, int64_t row0, int64_t rows, int d, double* out, cudaStream_t st) {
    cols_to_f64_kernel<<<blocks_for(rows * d), 256, 0, st>>>(x, row0, rows, d, out);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_max_finite(const float* v, int64_t n, float* out, cudaStream_t st) {
    max_finite_kernel<<<1, 1024, 0, st>>>(v, n, out);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
// One CTA per tile: the tile's float4s (a cell's 4 markers of one group) in through shared memory, out with consecutive
// threads writing consecutive float4s of the rows (both sides coalesced).
__global__ void __launch_bounds__(256) rows_from_tiles_kernel(const float* __restrict__ x, int64_t n_tiles, int d, float* __restrict__ rows) {
    extern __shared__ float4 s_tile[];   // [groups][128]
    const int groups = tile_groups(d);
    const size_t tile_fl = tile_floats(d);
    for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        const float* src = x + (size_t)t * tile_fl;
        for (int i = threadIdx.x; i < groups * kSlotCells; i += blockDim.x) {
            const int g = i >> 7, c = i & 127;
            s_tile[i] = *reinterpret_cast<const float4*>(src + (size_t)g * kGroupFloats + 4 * c);
        }
        __syncthreads();
        float4* dst = reinterpret_cast<float4*>(rows) + (size_t)t * kSlotCells * groups;
        for (int i = threadIdx.x; i < groups * kSlotCells; i += blockDim.x) {
            const int c = i / groups, g = i - c * groups;
            dst[i] = s_tile[g * kSlotCells + c];
        }
        __syncthreads();
    }
}
int launch_rows_from_tiles(const float* x, int64_t n_tiles, int d, float* rows, cudaStream_t st) {
    const size_t shm = (size_t)tile_groups(d) * kSlotCells * sizeof(float4);
    if (shm > 200 * 1024) return fail(DPR_ERR_INVALID, "rows_from_tiles: tile too large for shared memory");
    static size_t configured = 48 * 1024;
    if (shm > configured) {
        DPR_CUDA(cudaFuncSetAttribute(rows_from_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    const int grid = (int)std::min<int64_t>(n_tiles, 148 * 8);
    rows_from_tiles_kernel<<<grid, 256, shm, st>>>(x, n_tiles, d, rows);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_tile_norms(const float* x, int64_t n_tiles, int d, float* out, cudaStream_t st) {
    tile_norms_kernel<<<blocks_for(n_tiles * 32), 256, 0, st>>>(x, n_tiles, d, out);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_gather(const float* x, int64_t n, int d, const int64_t* rows, uint64_t seed, uint32_t fit0, int64_t m, float* out, int n_fits,
                  int64_t fit_stride, cudaStream_t st) {
    gather_rows_kernel<<<blocks_for(m * n_fits), 256, 0, st>>>(x, n, d, rows, seed, fit0, m, out, n_fits, fit_stride);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_bootstrap_rows(int64_t n, uint32_t boot_n, uint64_t seed, uint32_t fit, int64_t* rows, cudaStream_t st) {
    bootstrap_rows_kernel<<<blocks_for(boot_n), 256, 0, st>>>(n, boot_n, seed, fit, rows);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_seeding(SeedFit* d_fits, int n_fits, int64_t n_max, int d, int k, uint64_t seed, cudaStream_t st) {
    seed_first_kernel<<<n_fits, 64, 0, st>>>(d_fits, d, seed);
    count_launch();
    const int nb = (int)((n_max + kSeedBlock - 1) / kSeedBlock);
    const bool wide = n_max >= (1 << 20);   // four blocks per trip of a CTA
    const int gx = std::max(1, std::min(wide ? (nb + 3) / 4 : nb, (148 * 8 + n_fits - 1) / n_fits));
    for (int c = 1; c < k; ++c) {
        const SeedFit* cf = d_fits;
        if (wide) DPR_CUDA(launch_dependent(seed_dist_kernel<4>, dim3(gx, n_fits), dim3(kSeedBlock), sizeof(double) * d, st, cf, d, c, (int)(c == 1), seed));
        else DPR_CUDA(launch_dependent(seed_dist_kernel<1>, dim3(gx, n_fits), dim3(kSeedBlock), sizeof(double) * d, st, cf, d, c, (int)(c == 1), seed));
        count_launch();
    }
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_ssq(const float* x, int64_t n, int d, const double* mu, int k, const int32_t* labels, double* block_out, int blocks,
               cudaStream_t st) {
    const size_t shm = sizeof(double) * k * d;
    if (shm > 200 * 1024) return fail(DPR_ERR_INVALID, "cluster_norm: k * d too large for the shared-memory centre table");
    static size_t configured = 48 * 1024;
    if (shm > configured) {
        DPR_CUDA(cudaFuncSetAttribute(ssq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured = 200 * 1024;
    }
    ssq_kernel<<<blocks, 256, shm, st>>>(x, n, d, mu, k, labels, block_out);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_ari(const AriPair* d_pairs, int n_pairs, int64_t n_max, int kk, int mode, uint64_t seed, unsigned long long* d_tables,
               unsigned int* d_stab, int32_t* d_bad, double* d_out, cudaStream_t st) {
    const size_t shm = sizeof(unsigned int) * kk * kk;
    const int gx = std::max(1, std::min((int)((n_max + 8191) / 8192), std::max(1, 148 * 4 / n_pairs)));
    if (shm <= 160 * 1024) {
        static size_t configured = 0;
        if (shm > 48 * 1024 && shm > configured) {
            DPR_CUDA(cudaFuncSetAttribute(contingency_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            configured = 160 * 1024;
        }
        contingency_kernel<<<dim3(gx, n_pairs), 256, shm, st>>>(d_pairs, kk, d_tables, d_bad);
    } else {
        contingency_global_kernel<<<dim3(gx, n_pairs), 256, 0, st>>>(d_pairs, kk, d_tables, d_bad);
    }
    count_launch();
    if (mode == DPR_ARI_MONTE_CARLO) {
        ari_mc_count_kernel<<<dim3((10000 + 255) / 256, n_pairs), 256, 0, st>>>(d_pairs, seed, d_stab, d_bad, kk);
        ari_mc_final_kernel<<<n_pairs, 32, 0, st>>>(d_pairs, kk, d_tables, d_stab, d_out);
        count_launch(2);
    } else {
        ari_from_tables_kernel<<<n_pairs, 32, 0, st>>>(d_pairs, kk, d_tables, d_out);
        count_launch();
    }
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_ari_from_tables(int n_pairs, int64_t n, int kk, const unsigned long long* d_tables, double* d_out, cudaStream_t st) {
    ari_from_tables_n_kernel<<<n_pairs, 64, 0, st>>>((long long)n, kk, d_tables, d_out);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_label_presence(const int32_t* labels, int64_t n, int kk, int32_t* present, int32_t* bad, cudaStream_t st) {
    label_presence_kernel<<<blocks_for(n), 256, 0, st>>>(labels, n, kk, present, bad);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_label_counts(const int32_t* labels, int64_t n, int kk, unsigned long long* counts, cudaStream_t st) {
    DPR_CUDA(cudaMemsetAsync(counts, 0, sizeof(unsigned long long) * kk, st));
    label_counts_kernel<<<std::min(blocks_for(n), 148 * 8), 256, sizeof(unsigned int) * kk, st>>>(labels, n, kk, counts);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_synth(float* x, int64_t n, int d, const float* centres, int groups, uint64_t seed, int64_t row0, double* sum, double* sumsq,
                 cudaStream_t st) {
    synth_kernel<<<blocks_for(n), 256, 0, st>>>(x, n, d, centres, groups, seed, row0, sum, sumsq);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_scale(float* x, int64_t n_pad, int d, float inv_sd, cudaStream_t st) {
    scale_kernel<<<blocks_for((n_pad / kSlotCells) * (int64_t)tile_floats(d)), 256, 0, st>>>(x, n_pad, d, inv_sd);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
