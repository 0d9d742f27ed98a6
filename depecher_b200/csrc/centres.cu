// centres.cu — everything that happens to the K x D centres between two passes over the cells:
//   reduce_partials_kernel   fixed-order sum of the per-warp partial tables of fused_pass_kernel
//   finalize_kernel          convergence rule + soft-threshold update + fp32 chunk table
//   prepare_centres_kernel   chunk table for centres supplied by the caller (allocate_points, scoring)
//
// finalize restates the tail of Clusterer::reevaluate_centers (/root/reference/src/Clusterer.cpp:103-107,
//   mu = min((S+reg/2)/n, max((S-reg/2)/n, 0)) with the x86-64 packet NaN semantics, so an empty cluster
//   becomes an exact zero row) and the loop control of Clusterer::find_centers (:244-252: stop when the
//   labels did not change AND i > 20; penalty ramp reg*((float)i/(float)20) capped at reg).
#include "centres.cuh"

#include "comm.cuh"

#include <cuda_fp16.h>

#include <cstdio>

namespace dpr {

// Build header / chunk table / slot map of one problem from its fp64 centres.  One CTA; k, d are small.
// Active rows: no_zero -> rows with any non-zero coefficient (isZero(0), Clusterer.cpp:41-45; NaN counts as
// non-zero); otherwise every row, except that all-zero rows after the first are dropped: they are identical,
// and the reference's first-minimum rule (minCoeff, :64) can only ever pick the first of them.
__device__ void build_centre_table(const Problem& pr, const double* __restrict__ mu, int no_zero, int table_cap) {   // mu: pr.mu or a shared-memory copy of it
    constexpr int kMaxRows = kMaxKC * kMaxChunks;
    __shared__ int s_act[kMaxRows];
    __shared__ double s_cc[kMaxRows];
    __shared__ float s_ccf[kMaxRows];
    __shared__ unsigned char s_zero[kMaxRows];
    __shared__ int s_nact, s_bad;
    // The header is assembled in shared memory and copied to the problem's header in global memory once, at the end: the phases below
    // read what thread 0 has just written (chunk sizes and offsets, cc_max), and through global memory every such read was an L2 round trip
    __shared__ CentreHeader s_h;
    const int k = pr.k, d = pr.d;
    CentreHeader* h = &s_h;
    for (int i = threadIdx.x; i < (int)(sizeof(CentreHeader) / 4); i += blockDim.x) reinterpret_cast<int*>(&s_h)[i] = 0;
    const float xx_max_early = (threadIdx.x == 0 && pr.tc_table) ? *pr.xx_max : 0.f;   // (requested now, needed by thread 0 four phases later)
    if (threadIdx.x == 0) s_bad = 0;
    __syncthreads();
    // one thread per row: all-zero flag, non-finite flag, ||c||^2 in fp64
    for (int c = threadIdx.x; c < k && c < kMaxRows; c += blockDim.x) {
        bool zero = true, bad = false;
        double s = 0.0;
        for (int t = 0; t < d; ++t) {
            const double v = mu[(size_t)c * d + t];
            if (!(fabs(v) <= 0.0)) zero = false;
            if (!isfinite(v)) bad = true;
            s += v * v;
        }
        s_zero[c] = zero;
        s_cc[c] = s;
        s_ccf[c] = (float)s;
        if (bad) s_bad = 1;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        int n = 0, ref_active = 0, seen_zero = 0;
        float ccmax = 0.f;
        for (int c = 0; c < k && c < kMaxRows; ++c) {
            const bool zero = s_zero[c];
            bool use;
            if (no_zero) {
                use = !zero;
                ref_active += use;
            } else {
                ref_active += 1;
                use = !(zero && seen_zero);
                if (zero) seen_zero = 1;
            }
            if (use) {
                s_act[n++] = c;
                ccmax = fmaxf(ccmax, s_ccf[c]);   // (the largest ||c||^2 of the active rows, taken in the same sweep)
            }
        }
        s_nact = n;
        const bool trivial = ref_active < 2 || n < 2;
        const int n_chunks = trivial ? 0 : (n + kMaxKC - 1) / kMaxKC;
        h->k = k;
        h->k_act = n;
        h->n_chunks = n_chunks;
        h->trivial = trivial;
        h->force_exact = s_bad;
        int off = 0, base = 0;
        h->cc_max = ccmax * 1.0000002f;
        for (int q = 0; q < n_chunks; ++q) {
            const int real = n / n_chunks + (q < n % n_chunks ? 1 : 0);
            const int kc = (real + 1) & ~1;
            const int kcp = (kc + 3) & ~3;
            h->chunk_kc[q] = kc;
            h->chunk_off[q] = off;
            h->chunk_base[q] = base;
            off += kcp * (4 * tile_groups(d) + 1);   // marker rows padded to whole groups of 4 (zero rows)
            base += kc;
        }
        h->n_slots = base;
        h->table_floats = off;
        if (off > table_cap) {  // cannot happen for shapes admitted by plan_pass; keep the kernel safe anyway
            h->trivial = 1;
            h->n_chunks = 0;
        }
    }
    __syncthreads();
    const int n = s_nact, n_chunks = h->n_chunks;
    int real_before = 0;
    for (int q = 0; q < n_chunks; ++q) {
        const int real = n / n_chunks + (q < n % n_chunks ? 1 : 0);
        const int kc = h->chunk_kc[q], kcp = (kc + 3) & ~3;
        float* cc = pr.table + h->chunk_off[q];
        float* w = cc + kcp;
        for (int t = threadIdx.x; t < kcp; t += blockDim.x) {
            const int row = t < real ? s_act[real_before + t] : -1;
            cc[t] = row >= 0 ? (float)s_cc[row] : 3.0e38f;
            if (t < kc) pr.orig[h->chunk_base[q] + t] = row;
        }
        // (a session of the tensor-core pass runs the FFMA kernel only to sum given labels, which reads no coefficients: skipped there)
        if (!pr.tc_table)
            for (int e = threadIdx.x; e < kcp * 4 * tile_groups(d); e += blockDim.x) {
                const int j = e / kcp, t = e % kcp;
                w[e] = (t < real && j < d) ? (float)(-2.0 * mu[(size_t)s_act[real_before + t] * d + j]) : 0.f;
            }
        real_before += real;
    }
    // Tensor-core path (tc_pass.cu): the B operand of D = X * (-2 C)^T for tcgen05.mma kind::f16, K-major no-swizzle
    // canonical layout (core matrix = 8 rows x 8 halves).  Cells and centres are first multiplied by powers of two that
    // bring their largest magnitude to <= 2^13 (exact), then split as v = high + low with both parts fp16, so that
    // xh*ch + xh*cl + xl*ch carries fp32-grade accuracy (error model in tc_pass.cu).  ONE operand of 2*npad rows: rows
    // [0, npad) hold the high parts, rows [npad, 2*npad) the low parts, so a single product against xh yields xh*ch and
    // xh*cl side by side: half (row, marker) at byte (marker/8) * 32*npad + 16*row + 2*(marker%8).  Active rows in
    // order, no dummies inside; padding rows are zero and their cc is +huge.
    if (pr.tc_table) {
        const int npad = pr.tc_npad, g16 = 2 * ((d + 15) >> 4);
        float* cc = pr.tc_table;
        __half* b = reinterpret_cast<__half*>(cc + npad);
        __shared__ float s_sc;
        if (threadIdx.x == 0) {
            h->tc_n = max(16, (n + 15) & ~15);
            auto pow2_to_2p13 = [](float vmax) {   // exponent e with vmax * 2^e <= 2^13, clamped
                if (!(vmax > 0.f) || !(vmax < INFINITY)) return 0;
                int q;
                frexpf(vmax, &q);   // vmax = m * 2^q, m in [0.5, 1)
                return max(-60, min(60, 13 - q));
            };
            // no scaling (exponent 0) when the magnitudes already sit in fp16's comfortable range [2^-6, 2^13]
            auto scale_exp = [&](float vmax) {
                if (vmax >= 0.015625f && vmax <= 8192.f) return 0;
                return pow2_to_2p13(vmax);
            };
            const int ex = scale_exp(sqrtf(xx_max_early) * 1.0001f);
            const int ec = scale_exp(2.f * sqrtf(h->cc_max) * 1.0001f);
            h->tc_scale_x = ldexpf(1.f, ex);
            h->tc_unscale = ldexpf(1.f, -(ex + ec));
            // |X - xh - xl| <= 2^-22 |X| + 2^-25 (fp16 subnormal spacing), same for the centres: the absolute parts add up to
            // 2^-25 (2^-ec ||x||_1 + 2^-ex ||2c||_1) per score, ||.||_1 <= sqrt(d) ||.||_2; doubled for a difference of two scores
            const float k = 2.02f * 2.9802322e-8f * sqrtf((float)d);
            h->tc_abs1 = k * ldexpf(1.f, -ec);
            h->tc_abs2 = k * ldexpf(1.f, -ex) * 2.f * sqrtf(h->cc_max);
            s_sc = ldexpf(1.f, ec);
        }
        __syncthreads();
        const float sc = s_sc;
        for (int t = threadIdx.x; t < npad; t += blockDim.x) {
            cc[t] = t < n ? (float)s_cc[s_act[t]] : 3.0e38f;
            pr.tc_orig[t] = t < n ? s_act[t] : -1;
        }
        for (int e = threadIdx.x; e < g16 * npad * 8; e += blockDim.x) {
            const int g = e / (npad * 8), row = (e % (npad * 8)) >> 3, j = 8 * g + (e & 7);
            const float v = (row < n && j < d) ? (float)(-2.0 * mu[(size_t)s_act[row] * d + j]) * sc : 0.f;
            const __half hi = __float2half_rn(v);
            const __half lo = __float2half_rn(v - __half2float(hi));
            const size_t at = (size_t)g * (16 * npad) + (size_t)row * 8 + (e & 7);   // in halves
            b[at] = hi;
            b[at + (size_t)npad * 8] = lo;
        }
    }
    __syncthreads();   // (the header is complete: tc_n and the scales were the last fields)
    for (int i = threadIdx.x; i < (int)(sizeof(CentreHeader) / 4); i += blockDim.x) reinterpret_cast<int*>(pr.hdr)[i] = reinterpret_cast<const int*>(&s_h)[i];
}

__global__ void prepare_centres_kernel(Problem* problems, int table_cap) {
    const Problem pr = problems[blockIdx.x];
    build_centre_table(pr, pr.mu, pr.no_zero, table_cap);
}

// ---------------------------------------------------------------------------------------------------
// reduce: out[p][seg][entry] = sum, in ascending CTA order, of the (CTA, problem) tables written by the
// fused pass for the CTAs of segment `seg` whose tile range touches problem p.
// entry layout: [k_max*d sums][k_max counts][1 changed]
// ---------------------------------------------------------------------------------------------------
__global__ void reduce_partials_kernel(const Problem* problems, const int32_t* cta_tile_start, int grid_ctas, int k_max, int d,
                                       const double* cta_tables, double* seg_out, int n_seg, int32_t* n_active_out) {
    const int p = blockIdx.x, seg = blockIdx.y;
    if (n_active_out && p == 0 && seg == 0 && threadIdx.x == 0) *n_active_out = 0;   // finalize counts the problems still running
    const Problem pr = problems[p];
    const int64_t entries = (int64_t)k_max * d + k_max + 1;
    double* out = seg_out + ((int64_t)p * n_seg + seg) * entries;
    const int c0 = (int)((int64_t)grid_ctas * seg / n_seg), c1 = (int)((int64_t)grid_ctas * (seg + 1) / n_seg);
    const int t0 = pr.first_tile, t1 = pr.first_tile + pr.n_tiles;
    // the CTAs of this segment whose tile range touches the problem are consecutive: [ca, cb)
    int ca = c1, cb = c0;
    if (!pr.done)
        for (int c = c0; c < c1; ++c) {
            const int a = cta_tile_start[c], b = cta_tile_start[c + 1];
            if (a >= t1 || b <= t0 || a >= b) continue;
            ca = min(ca, c);
            cb = max(cb, c + 1);
        }
    for (int64_t e = threadIdx.x; e < entries; e += blockDim.x) {
        double acc = 0.0;
        for (int c = ca; c < cb; c += 8) {   // 8 loads in flight, then the fixed-order sum
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int cc = c + q;
                bool use = cc < cb;
                if (use) {
                    const int a = cta_tile_start[cc], b = cta_tile_start[cc + 1];
                    use = !(a >= t1 || b <= t0 || a >= b);
                }
                v[q] = use ? cta_tables[(size_t)(cc + p) * entries + e] : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += v[q];
        }
        out[e] = acc;
    }
}

// ---------------------------------------------------------------------------------------------------
// finalize: one CTA per problem.  iter = index i of the assignment that just ran.
// ---------------------------------------------------------------------------------------------------
// `in`: n_seg tables of this problem ([k_max*d sums][k_max counts][1 changed] each), summed in segment order; s_mu: shared memory
// for the new centres when they fit (mu_in_smem): later phases read them there instead of through L2
#ifdef DPR_RF_PROFILE
#define RF_STAMP(i) do { if (threadIdx.x == 0) rf_t[i] = clock64(); } while (0)
__device__ long long rf_t[16];
#else
#define RF_STAMP(i)
#endif
__device__ void finalize_body(Problem* problems, int p, const Problem& pr, const double* in, int n_seg, int k_max, int iter, int max_iter, int table_cap,
                              int allow_delta, int32_t* n_active_out, int mu_in_smem, double* s_mu) {
    const int k = pr.k, d = pr.d;
    RF_STAMP(5);
    const int64_t stride = (int64_t)k_max * d;
    const int64_t entries = stride + k_max + 1;
    __shared__ long long s_changed;
    // totals in fixed segment order.  First the counts and the changed-label count (the update of every coefficient needs its row's count) ...
    auto total = [&](int64_t e) {
        double acc = 0.0;
        for (int s0 = 0; s0 < n_seg; s0 += 16) {   // all loads in flight, then the fixed-order sum
            double v[16];
#pragma unroll
            for (int q = 0; q < 16; ++q) v[q] = s0 + q < n_seg ? in[(int64_t)(s0 + q) * entries + e] : 0.0;
#pragma unroll
            for (int q = 0; q < 16; ++q) acc += v[q];
        }
        return acc;
    };
    for (int64_t e = stride + threadIdx.x; e < entries; e += blockDim.x) {
        const double acc = total(e);
        if (e < stride + k_max) {
            const int c = (int)(e - stride);
            if (c < k) pr.counts[c] = (pr.delta ? pr.counts[c] : 0LL) + (long long)acc;
        } else {
            s_changed = (long long)acc;
        }
    }
    __threadfence_block();
    __syncthreads();
    RF_STAMP(6);
    const long long changed = s_changed;
    const bool converged = (changed == 0 && iter > 20);
    if (threadIdx.x == 0) {
        problems[p].changed = changed;
        problems[p].n_assign = iter + 1;
        // every pass after the first one scatters only the cells whose label changes (sums += delta): the running sums are
        // valid from here on, and even a slot where every cell moves costs at most two contributions per cell
        problems[p].delta = allow_delta ? 1 : 0;
    }
    const bool last = iter + 1 >= max_iter;
    if (threadIdx.x == 0 && (converged || last)) problems[p].done = 1;
    // ... then the running sums and, unless the fit has just converged (Clusterer.cpp:247-249: break BEFORE the update, the centres
    // stay those of the previous iteration; at the iteration cap the reference has just run its last update, so it is applied), the
    // soft threshold in the same sweep.   reg_i = min(reg, reg * ((float)i / (float)20))   (Clusterer.cpp:251)
    const double ramp = pr.reg * (double)((float)iter / (float)20);
    const double reg_i = pr.reg < ramp ? pr.reg : ramp;
    for (int64_t e = threadIdx.x; e < stride; e += blockDim.x) {
        const double acc = total(e);
        const int c = (int)(e / d), j = (int)(e % d);
        if (c >= k) continue;
        const size_t at = (size_t)c * d + j;
        const double S = (pr.delta ? pr.sums[at] : 0.0) + acc;
        pr.sums[at] = S;
        if (converged) continue;
        const double nn = (double)pr.counts[c];
        const double lo = (S - reg_i / 2) / nn;
        const double hi = (S + reg_i / 2) / nn;
        const double m0 = lo > 0.0 ? lo : 0.0;  // _mm_max_pd(lo, 0): second operand when lo is NaN
        const double m = hi < m0 ? hi : m0;     // _mm_min_pd(hi, m0)
        pr.mu[at] = m;
        if (mu_in_smem) s_mu[at] = m;
    }
    if (converged) return;
    __threadfence_block();
    __syncthreads();
    RF_STAMP(7);
    build_centre_table(pr, mu_in_smem ? s_mu : pr.mu, pr.no_zero, table_cap);
    __syncthreads();
    RF_STAMP(8);
    if (threadIdx.x == 0 && n_active_out && !(iter + 1 >= max_iter)) atomicAdd(n_active_out, 1);
#ifdef DPR_RF_PROFILE
    if (threadIdx.x == 0 && iter == 5)
        printf("reduce_finalize (cycles): find ranges %lld, table sums %lld, cluster barrier %lld, dsmem sums %lld, barrier %lld | running sums %lld, soft threshold %lld, centre tables %lld\n",
               rf_t[1] - rf_t[0], rf_t[2] - rf_t[1], rf_t[3] - rf_t[2], rf_t[4] - rf_t[3], rf_t[5] - rf_t[4], rf_t[6] - rf_t[5], rf_t[7] - rf_t[6], rf_t[8] - rf_t[7]);
#endif
}

__global__ void finalize_kernel(Problem* problems, const double* seg_in, int n_seg, int k_max, int iter, int max_iter, int table_cap,
                                int allow_delta, int32_t* n_active_out, int mu_in_smem) {
    const int p = blockIdx.x;
    const Problem pr = problems[p];
    if (pr.done) return;
    extern __shared__ double s_dyn[];
    const int64_t entries = (int64_t)k_max * pr.d + k_max + 1;
    finalize_body(problems, p, pr, seg_in + (int64_t)p * n_seg * entries, n_seg, k_max, iter, max_iter, table_cap, allow_delta, n_active_out, mu_in_smem, s_dyn);
}

// ---------------------------------------------------------------------------------------------------
// reduce + finalize in one launch for batches of MANY problems (the bootstrap fits of a penalty search: 220 problems of 10 000 cells, each
// touched by one to three CTAs of the pass): one CTA per problem adds the (CTA, problem) tables of those CTAs in ascending order into its
// shared memory and carries on as finalize_kernel does.  (The two-kernel path spends a launch of 8-16 us per iteration on
// reduce_partials_kernel, whose 16 segment tables per problem are almost all empty for such batches.)
// n_active: two counters, as in reduce_finalize_kernel below.
// ---------------------------------------------------------------------------------------------------
__global__ void finalize_direct_kernel(Problem* problems, const int32_t* cta_tile_start, int grid_ctas, const double* cta_tables, int k_max, int iter,
                                       int max_iter, int table_cap, int allow_delta, int32_t* n_active2, int mu_in_smem) {
    const int p = blockIdx.x;
    DPR_GRID_DEPENDENCY();
    if (p == 0 && threadIdx.x == 0) n_active2[(iter + 1) & 1] = 0;
    const Problem pr = problems[p];
    if (pr.done) return;
    extern __shared__ double s_dyn[];
    const int64_t entries = (int64_t)k_max * pr.d + k_max + 1;
    double* s_mu = s_dyn;
    double* s_part = s_dyn + (mu_in_smem ? (size_t)k_max * pr.d : 0);
    const int t0 = pr.first_tile, t1 = pr.first_tile + pr.n_tiles;
    // the CTAs of the pass whose tile range touches the problem are consecutive: [ca, cb).  Ranges are sorted: the first CTA whose
    // range ends beyond t0, the first one whose range starts at or beyond t1 (binary searches by one thread: a few dependent loads)
    __shared__ int s_ca, s_cb;
    if (threadIdx.x == 0) {
        int lo = 0, hi = grid_ctas;
        while (lo < hi) {   // first c with cta_tile_start[c + 1] > t0
            const int mid = (lo + hi) >> 1;
            if (cta_tile_start[mid + 1] > t0) hi = mid; else lo = mid + 1;
        }
        s_ca = lo;
        hi = grid_ctas;
        while (lo < hi) {   // first c >= ca with cta_tile_start[c] >= t1
            const int mid = (lo + hi) >> 1;
            if (cta_tile_start[mid] >= t1) hi = mid; else lo = mid + 1;
        }
        s_cb = lo;
    }
    __syncthreads();
    const int ca = s_ca, cb = s_cb;
    for (int64_t e = threadIdx.x; e < entries; e += blockDim.x) {
        double acc = 0.0;
        for (int c = ca; c < cb; c += 8) {   // 8 loads in flight, then the fixed-order sum
            double v[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const int cc = c + q;
                bool use = cc < cb;
                if (use) use = cta_tile_start[cc] < cta_tile_start[cc + 1];   // (a CTA without tiles wrote nothing)
                v[q] = use ? __ldcg(cta_tables + (size_t)(cc + p) * entries + e) : 0.0;
            }
#pragma unroll
            for (int q = 0; q < 8; ++q) acc += v[q];
        }
        s_part[e] = acc;
    }
    __syncthreads();
    finalize_body(problems, p, pr, s_part, 1, k_max, iter, max_iter, table_cap, allow_delta, n_active2 + (iter & 1), mu_in_smem, s_mu);
}

// ---------------------------------------------------------------------------------------------------
// reduce + finalize in one launch (single GPU: nothing to exchange between the two).  A cluster of kFuseCtas CTAs per problem:
// CTA r adds up the r-th eighth of the (CTA, problem) tables the pass wrote, in ascending CTA order, into its shared memory;
// after the cluster barrier CTA 0 adds the eight partial tables in rank order through distributed shared memory and carries on
// as finalize_kernel does.  Every order is fixed: results are bit-reproducible.
// n_active: two counters; iteration i counts the problems still running in n_active[i & 1] and clears the other one.
// ---------------------------------------------------------------------------------------------------
constexpr int kFuseCtas = 8;
__global__ void __cluster_dims__(kFuseCtas, 1, 1) reduce_finalize_kernel(Problem* problems, const int32_t* cta_tile_start, int grid_ctas, const double* cta_tables,
                                                                        int k_max, int iter, int max_iter, int table_cap, int allow_delta,
                                                                        int32_t* n_active2, int mu_in_smem, const PeerExchange px) {
    const int p = blockIdx.y;
    const unsigned r = blockIdx.x;   // rank in the cluster
    DPR_GRID_DEPENDENCY();
    const Problem pr = problems[p];
    extern __shared__ double s_dyn[];
    const int64_t entries = (int64_t)k_max * pr.d + k_max + 1;
    double* s_mu = s_dyn;
    double* s_part = s_dyn + (mu_in_smem ? (size_t)k_max * pr.d : 0);   // this CTA's partial table; in CTA 0 later the total
    if (p == 0 && r == 0 && threadIdx.x == 0) n_active2[(iter + 1) & 1] = 0;
    if (r == 0) RF_STAMP(0);
    if (!pr.done) {
        const int t0 = pr.first_tile, t1 = pr.first_tile + pr.n_tiles;
        // the CTAs of the pass whose tile range touches the problem are consecutive: [ca, cb)
        __shared__ int s_ca, s_cb;
        if (threadIdx.x == 0) {
            s_ca = grid_ctas;
            s_cb = 0;
        }
        __syncthreads();
        for (int c = threadIdx.x; c < grid_ctas; c += blockDim.x) {
            const int a = cta_tile_start[c], b = cta_tile_start[c + 1];
            if (a >= t1 || b <= t0 || a >= b) continue;
            atomicMin(&s_ca, c);
            atomicMax(&s_cb, c + 1);
        }
        __syncthreads();
        const int ca = s_ca, cb = max(s_cb, s_ca);
        if (r == 0) RF_STAMP(1);
        const int c0 = ca + (int)((int64_t)(cb - ca) * r / kFuseCtas), c1 = ca + (int)((int64_t)(cb - ca) * (r + 1) / kFuseCtas);
        // which of this CTA's tables exist (the pass CTAs whose range touches the problem): one bit each, found once
        constexpr int kInFlight = 24;
        __shared__ unsigned s_use[8];   // up to 8 x 24 tables per CTA of the cluster
        if (threadIdx.x < 8 * 32) {
            const int chunk = threadIdx.x >> 5, cc = c0 + chunk * kInFlight + (threadIdx.x & 31);
            bool use = (threadIdx.x & 31) < kInFlight && cc < c1;
            if (use) {
                const int a = cta_tile_start[cc], b = cta_tile_start[cc + 1];
                use = !(a >= t1 || b <= t0 || a >= b);
            }
            const unsigned m = __ballot_sync(0xffffffffu, use);
            if ((threadIdx.x & 31) == 0) s_use[chunk] = m;
        }
        __syncthreads();
        for (int64_t e = threadIdx.x; e < entries; e += blockDim.x) {
            double acc = 0.0;
            for (int c = c0, chunk = 0; c < c1; c += kInFlight, ++chunk) {   // all loads of the chunk in flight, then the fixed-order sum
                const unsigned m = chunk < 8 ? s_use[chunk] : 0u;
                double v[kInFlight];
#pragma unroll
                for (int q = 0; q < kInFlight; ++q) {
                    bool use = (m >> q) & 1u;
                    if (chunk >= 8) {   // (more than 192 tables per CTA: no pass grid is that large; kept correct all the same)
                        const int cc = c + q;
                        use = cc < c1;
                        if (use) {
                            const int a = cta_tile_start[cc], b = cta_tile_start[cc + 1];
                            use = !(a >= t1 || b <= t0 || a >= b);
                        }
                    }
                    v[q] = use ? __ldcg(cta_tables + (size_t)(c + q + p) * entries + e) : 0.0;
                }
#pragma unroll
                for (int q = 0; q < kInFlight; ++q) acc += v[q];
            }
            s_part[e] = acc;
        }
    }
    if (r == 0) RF_STAMP(2);
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (r == 0) RF_STAMP(3);
    if (r == 0 && !pr.done) {
        const uint32_t mine = (uint32_t)__cvta_generic_to_shared(s_part);
        for (int64_t e = threadIdx.x; e < entries; e += blockDim.x) {
            double v[kFuseCtas];
#pragma unroll
            for (int q = 0; q < kFuseCtas; ++q) {   // partial table of CTA q of the cluster
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(mine + (uint32_t)(e * 8)), "r"(q));
                asm volatile("ld.shared::cluster.f64 %0, [%1];" : "=d"(v[q]) : "r"(remote));
            }
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < kFuseCtas; ++q) acc += v[q];
            s_part[e] = acc;   // (each thread overwrites only the entries it has just read from its own table)
        }
    }
    if (r == 0) RF_STAMP(4);
    // the other CTAs' tables must stay alive until CTA 0 has read them
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    if (r != 0 || pr.done) return;
    __syncthreads();
    if (px.world > 1) {
        // rows sharded over ranks: this rank's table goes into every rank's mailbox over NVLink (plain peer stores), then the
        // flag of this (source, problem) is raised there; the tables of all ranks are then read out of the own mailbox and added
        // in rank order — every rank forms the same sum.  Mailboxes alternate with the parity of the exchange number: a rank can
        // only be one exchange ahead of the slowest one, because it needs that one's flag to finish its own.
        const unsigned par = (unsigned)(px.seq & 1ull);
        const size_t slot = ((size_t)par * kMaxPeers + px.rank) * kPeerDoubles + (size_t)p * entries;
        for (int q = 0; q < px.world; ++q)
            for (int64_t e = threadIdx.x; e < entries; e += blockDim.x) px.tables[q][slot + e] = s_part[e];
        __threadfence_system();
        __syncthreads();
        if (threadIdx.x < px.world) {
            unsigned long long* f = px.flags[threadIdx.x] + ((size_t)par * kMaxPeers + px.rank) * kPeerFlags + p;
            asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(px.seq) : "memory");
        }
        if (threadIdx.x < px.world) {   // wait for source `threadIdx.x`
            const unsigned long long* f = px.flags[px.rank] + ((size_t)par * kMaxPeers + threadIdx.x) * kPeerFlags + p;
            unsigned long long v = 0;
            const long long t0 = clock64();
            do {
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
                if (v != px.seq && clock64() - t0 > (20ll << 30)) {   // ~10 s: a rank is missing — give up instead of hanging the GPU
                    *px.failed = 1;
                    break;
                }
            } while (v != px.seq);
        }
        __syncthreads();
        const double* mine = px.tables[px.rank] + (size_t)par * kMaxPeers * kPeerDoubles + (size_t)p * entries;
        for (int64_t e = threadIdx.x; e < entries; e += blockDim.x) {
            double v[kMaxPeers];
#pragma unroll
            for (int q = 0; q < kMaxPeers; ++q) v[q] = q < px.world ? __ldcg(mine + (size_t)q * kPeerDoubles + e) : 0.0;
            double acc = 0.0;
#pragma unroll
            for (int q = 0; q < kMaxPeers; ++q) acc += v[q];
            s_part[e] = acc;
        }
        __syncthreads();
    }
    finalize_body(problems, p, pr, s_part, 1, k_max, iter, max_iter, table_cap, allow_delta, n_active2 + (iter & 1), mu_in_smem, s_mu);
}

// ---------------------------------------------------------------------------------------------------
// re-partition: contiguous tile ranges per CTA holding equal numbers of tiles of problems that are still running
// (converged fits keep their place in the tile order but weigh nothing), so late Lloyd iterations of a batch stay balanced.
// One CTA.  Boundaries only move; the tile numbering, and with it every warp's accumulation order inside a range, is fixed
// by (range, problem), so results stay reproducible for a given sequence of convergences.
// ---------------------------------------------------------------------------------------------------
__global__ void partition_kernel(const Problem* problems, int n_problems, int grid_ctas, int total_tiles, int32_t* cta_tile_start) {
    extern __shared__ long long s_pref[];   // [n_problems + 1] running count of live tiles
    if (threadIdx.x == 0) {
        long long acc = 0;
        for (int p = 0; p < n_problems; ++p) {
            s_pref[p] = acc;
            acc += problems[p].done ? 0 : problems[p].n_tiles;
        }
        s_pref[n_problems] = acc;
    }
    __syncthreads();
    const long long live = s_pref[n_problems];
    for (int c = threadIdx.x; c <= grid_ctas; c += blockDim.x) {
        int start = total_tiles;
        if (live > 0 && c < grid_ctas) {
            const long long target = live * c / grid_ctas;   // first live tile of CTA c
            int lo = 0, hi = n_problems - 1;                 // last problem whose prefix <= target and that is live
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (s_pref[mid] <= target) lo = mid; else hi = mid - 1;
            }
            while (lo < n_problems - 1 && s_pref[lo + 1] == s_pref[lo]) ++lo;   // skip converged problems
            start = problems[lo].first_tile + (int)(target - s_pref[lo]);
        }
        if (c == 0) start = 0;
        cta_tile_start[c] = start;
    }
}

// one update from given sums/counts (reevaluate_centers for explicit labels): mu only
__global__ void soft_threshold_kernel(const double* sums, const long long* counts, int k, int d, double reg, double* mu) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < k * d; e += gridDim.x * blockDim.x) {
        const double S = sums[e], nn = (double)counts[e / d];
        const double lo = (S - reg / 2) / nn, hi = (S + reg / 2) / nn;
        const double m0 = lo > 0.0 ? lo : 0.0;
        mu[e] = hi < m0 ? hi : m0;
    }
}

int launch_prepare_centres(Problem* d_problems, int n_problems, int table_cap, cudaStream_t st) {
    prepare_centres_kernel<<<n_problems, 128, 0, st>>>(d_problems, table_cap);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

int launch_reduce_partials(const Problem* d_problems, int n_problems, const int32_t* cta_tile_start, int grid_ctas, int k_max, int d,
                           const double* cta_tables, double* seg_out, int n_seg, int32_t* n_active_out, cudaStream_t st) {
    dim3 g(n_problems, n_seg);
    reduce_partials_kernel<<<g, 256, 0, st>>>(d_problems, cta_tile_start, grid_ctas, k_max, d, cta_tables, seg_out, n_seg, n_active_out);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

int launch_finalize(Problem* d_problems, int n_problems, const double* seg_in, int n_seg, int k_max, int d, int iter, int max_iter,
                    int table_cap, int allow_delta, int32_t* n_active_out, cudaStream_t st) {
    const size_t mu_bytes = (size_t)k_max * d * sizeof(double);
    const int mu_in_smem = mu_bytes <= 64 * 1024;
    static bool configured = false;
    if (!configured) {
        DPR_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
        configured = true;
    }
    finalize_kernel<<<n_problems, 512, mu_in_smem ? mu_bytes : 0, st>>>(d_problems, seg_in, n_seg, k_max, iter, max_iter, table_cap, allow_delta, n_active_out,
                                                                        mu_in_smem);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

int launch_reduce_finalize(Problem* d_problems, int n_problems, const int32_t* cta_tile_start, int grid_ctas, const double* cta_tables, int k_max, int d,
                           int iter, int max_iter, int table_cap, int allow_delta, int32_t* n_active2, bool exchange, cudaStream_t st) {
    PeerExchange px{};
    px.world = 1;
    if (exchange && ctx().world > 1) DPR_TRY(peer_exchange_next(&px));   // (row-sharded sessions only: Session::exchanges)
    const size_t mu_bytes = (size_t)k_max * d * sizeof(double);
    const int mu_in_smem = mu_bytes <= 64 * 1024;
    const size_t part_bytes = ((size_t)k_max * d + k_max + 1) * sizeof(double);
    const size_t smem = (mu_in_smem ? mu_bytes : 0) + part_bytes;
    static bool configured = false;
    if (!configured) {
        DPR_CUDA(cudaFuncSetAttribute(reduce_finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured = true;
    }
    dim3 g(kFuseCtas, n_problems);
    DPR_CUDA(launch_dependent(reduce_finalize_kernel, g, dim3(512), smem, st, d_problems, cta_tile_start, grid_ctas, cta_tables, k_max, iter, max_iter, table_cap,
                              allow_delta, n_active2, mu_in_smem, px));
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
int launch_finalize_direct(Problem* d_problems, int n_problems, const int32_t* cta_tile_start, int grid_ctas, const double* cta_tables, int k_max, int d,
                           int iter, int max_iter, int table_cap, int allow_delta, int32_t* n_active2, cudaStream_t st) {
    const size_t mu_bytes = (size_t)k_max * d * sizeof(double);
    const int mu_in_smem = mu_bytes <= 64 * 1024;
    const size_t smem = (mu_in_smem ? mu_bytes : 0) + ((size_t)k_max * d + k_max + 1) * sizeof(double);
    static bool configured = false;
    if (!configured) {
        DPR_CUDA(cudaFuncSetAttribute(finalize_direct_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
        configured = true;
    }
    DPR_CUDA(launch_dependent(finalize_direct_kernel, dim3(n_problems), dim3(512), smem, st, d_problems, cta_tile_start, grid_ctas, cta_tables, k_max, iter, max_iter,
                              table_cap, allow_delta, n_active2, mu_in_smem));
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}
bool reduce_finalize_peer_fits(int n_problems, int k_max, int d) {   // the tables of one exchange fit a mailbox slot
    return n_problems <= kPeerFlags && (size_t)n_problems * ((size_t)k_max * d + k_max + 1) <= kPeerDoubles;
}
bool reduce_finalize_fits(int k_max, int d) {
    const size_t mu_bytes = (size_t)k_max * d * sizeof(double);
    return (mu_bytes <= 64 * 1024 ? mu_bytes : 0) + ((size_t)k_max * d + k_max + 1) * sizeof(double) <= 160 * 1024;
}

int launch_partition(const Problem* d_problems, int n_problems, int grid_ctas, int total_tiles, int32_t* cta_tile_start, cudaStream_t st) {
    partition_kernel<<<1, 256, sizeof(long long) * (n_problems + 1), st>>>(d_problems, n_problems, grid_ctas, total_tiles, cta_tile_start);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

int launch_soft_threshold(const double* sums, const long long* counts, int k, int d, double reg, double* mu, cudaStream_t st) {
    soft_threshold_kernel<<<(k * d + 255) / 256, 256, 0, st>>>(sums, counts, k, d, reg, mu);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
