// fused_pass.cu — the hot kernel: nearest-centre assignment fused with the per-cluster sums/counts.
//
// Replaces Clusterer::allocate_clusters (/root/reference/src/Clusterer.cpp:34-87) and the scatter-add
// half of Clusterer::reevaluate_centers (:99-102) with ONE pass over the cells.
//
// Shape of the work:
//   * X lives in HBM as column-major fp32 (R's own layout, so every marker column is a contiguous
//     stream).  A warp owns a "slot" of 128 consecutive cells x d markers in shared memory, filled by
//     d bulk-async copies (TMA, cp.async.bulk + mbarrier complete_tx) of 512 B each.
//   * Distance micro-kernel: each lane holds 4 cells x up to 32 centres = 128 fp32 accumulators as
//     packed pairs and issues FFMA2 (fma.rn.f32x2): s_k = ||c_k||^2 - 2 x.c_k.  One LDS.128 brings the
//     lane's 4 cells of a marker, one broadcast LDS.128 brings 4 centres.  Measured on B200: this tile
//     shape runs at 98 % of the FP32 pipe, scalar FFMA or narrower tiles lose 10-30 %
//     (profiles/microbench/r01_pipes_b200.log).
//   * Argmin in registers with the centre index packed into the 5 low mantissa bits; cells whose two
//     best fp32 scores are closer than a proven error bound are re-decided in fp64 exactly as the
//     reference does (sqrt of the left-to-right sum, strict <), so labels are bit-exact for
//     fp32-representable data.
//   * Lloyd mode: the warp counting-sorts its 128 labels (match.any, no atomics), then lanes switch to
//     marker ownership and add each label's run in fp64 registers before one read-modify-write of the
//     warp's PRIVATE partial table.  Fixed tile->warp map + private tables => bit-reproducible sums.
#include "common.cuh"
#include "fused_pass.cuh"

namespace dpr {

// ---------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + bulk async copy (TMA, SASS UBLKCP)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------------------------------------------
// distance micro-kernel for one chunk of KC centres
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pack_idx(float s, int idx) {
    return __uint_as_float((__float_as_uint(s) & 0xFFFFFFE0u) | (unsigned)idx);
}

template <int KC, bool FIRST>
__device__ __forceinline__ void dist_chunk(const float* __restrict__ slot, const float* __restrict__ tab, int d, int lane,
                                           float (&xx)[4], float (&best)[4], float (&second)[4]) {
    constexpr int NP = KC / 2;
    constexpr int KCP = (KC + 3) & ~3;
    float2 acc[4][NP];
    {
        const float2* cc = reinterpret_cast<const float2*>(tab);
#pragma unroll
        for (int t = 0; t < NP; ++t) {
            float2 c = cc[t];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r][t] = c;
        }
    }
    const float4* xs = reinterpret_cast<const float4*>(slot) + lane;
    const float4* cw = reinterpret_cast<const float4*>(tab + KCP);
#pragma unroll 2
    for (int j = 0; j < d; ++j) {
        const float4 xv = xs[j * (kRowStride / 4)];
        const float x[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int q = 0; q < KCP / 4; ++q) {
            const float4 c = cw[j * (KCP / 4) + q];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 xr = make_float2(x[r], x[r]);
                acc[r][2 * q] = __ffma2_rn(xr, make_float2(c.x, c.y), acc[r][2 * q]);
                if (2 * q + 1 < NP) acc[r][2 * q + 1] = __ffma2_rn(xr, make_float2(c.z, c.w), acc[r][2 * q + 1]);
            }
        }
        if (FIRST) {
#pragma unroll
            for (int r = 0; r < 4; ++r) xx[r] = fmaf(x[r], x[r], xx[r]);
        }
    }
#pragma unroll
    for (int t = 0; t < NP; ++t) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float a = pack_idx(acc[r][t].x, 2 * t);
            const float b = pack_idx(acc[r][t].y, 2 * t + 1);
            second[r] = fminf(second[r], fmaxf(a, best[r]));
            best[r] = fminf(best[r], a);
            second[r] = fminf(second[r], fmaxf(b, best[r]));
            best[r] = fminf(best[r], b);
        }
    }
}

template <bool FIRST>
__device__ __forceinline__ void dist_chunk_dispatch(int kc, const float* slot, const float* tab, int d, int lane, float (&xx)[4],
                                                    float (&best)[4], float (&second)[4]) {
    switch (kc) {
#define DPR_CASE(KC)                                                \
    case KC:                                                        \
        dist_chunk<KC, FIRST>(slot, tab, d, lane, xx, best, second); \
        break;
        DPR_CASE(2) DPR_CASE(4) DPR_CASE(6) DPR_CASE(8) DPR_CASE(10) DPR_CASE(12) DPR_CASE(14) DPR_CASE(16)
        DPR_CASE(18) DPR_CASE(20) DPR_CASE(22) DPR_CASE(24) DPR_CASE(26) DPR_CASE(28) DPR_CASE(30) DPR_CASE(32)
#undef DPR_CASE
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------------
// exact decision for one cell: the reference's arithmetic (Clusterer.cpp:56-66) in fp64 —
// d_k = sqrt(sum_t ((-x_t) + mu_kt)^2), separate multiply and add (no contraction: the reference is
// built for baseline x86-64, which has no FMA), strict-< first minimum over the active rows in order.
// ---------------------------------------------------------------------------------------------------
__device__ __noinline__ int exact_label(const float* __restrict__ slot, int cell, int d, const double* __restrict__ mu,
                                        const int32_t* __restrict__ orig, int n_slots) {
    int best = -1;
    double bv = 0.0;
    for (int a = 0; a < n_slots; ++a) {
        const int row = orig[a];
        if (row < 0) continue;
        const double* m = mu + (size_t)row * d;
        double s = 0.0;
        for (int t = 0; t < d; ++t) {
            const double diff = __dadd_rn(-(double)slot[t * kRowStride + cell], m[t]);
            s = __dadd_rn(s, __dmul_rn(diff, diff));
        }
        const double dist = sqrt(s);
        if (best < 0) {
            best = row;
            bv = dist;
        } else if (dist < bv) {
            bv = dist;
            best = row;
        }
    }
    return best < 0 ? 0 : best;
}

// ---------------------------------------------------------------------------------------------------
// Lloyd mode: sums/counts of one slot into the warp's private partial table
// ---------------------------------------------------------------------------------------------------
// off[0..k] (ints) and perm[0..127] (bytes) are the warp's scratch in shared memory.
__device__ __forceinline__ void slot_sums(const float* __restrict__ slot, const int (&lab)[4], int d, int k, int lane, int* off,
                                          unsigned char* perm, double* __restrict__ sums, long long* __restrict__ counts) {
    const unsigned full = 0xffffffffu;
    for (int i = lane; i <= k; i += 32) off[i] = 0;
    __syncwarp();
    int pos[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const int L = lab[r];
        const unsigned m = __match_any_sync(full, L);
        const int leader = __ffs(m) - 1;
        int base = 0;
        if (L >= 0 && lane == leader) {
            base = off[L + 1];
            off[L + 1] = base + __popc(m);
        }
        base = __shfl_sync(full, base, leader);
        pos[r] = base + __popc(m & ((1u << lane) - 1u));
        __syncwarp();
    }
    // counts -> exclusive offsets (off[L] = first position of label L, off[k] = valid cells)
    int carry = 0;
    for (int b = 0; b <= k; b += 32) {
        const int i = b + lane;
        int v = (i <= k) ? off[i] : 0;
        int s = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, s, o);
            if (lane >= o) s += t;
        }
        if (i <= k) off[i] = carry + s;  // inclusive over entries 0..i  (entry 0 is 0)
        carry += __shfl_sync(full, s, 31);
    }
    __syncwarp();
    // off[i] now = number of cells with label < i  (entry i held the count of label i-1)
#pragma unroll
    for (int r = 0; r < 4; ++r)
        if (lab[r] >= 0) perm[off[lab[r]] + pos[r]] = (unsigned char)(4 * lane + r);
    __syncwarp();
    for (int L = lane; L < k; L += 32) {
        const int c = off[L + 1] - off[L];
        if (c) counts[L] += c;
    }
    for (int jb = 0; jb < d; jb += 32) {
        const int j = jb + lane;
        const bool on = j < d;
        const float* row = slot + (on ? j : 0) * kRowStride;
        for (int Lb = 0; Lb < k; Lb += 32) {
            const int Li = Lb + lane;
            const bool present = (Li < k) && (off[Li + 1] > off[Li]);
            unsigned pm = __ballot_sync(full, present);
            while (pm) {
                const int L = Lb + __ffs(pm) - 1;
                pm &= pm - 1;
                const int s = off[L], e = off[L + 1];
                double* cell = sums + (size_t)L * d + j;
                const double old = on ? *cell : 0.0;
                double a0 = 0.0, a1 = 0.0;
                int p = s;
                for (; p + 1 < e; p += 2) {
                    a0 += (double)row[perm[p]];
                    a1 += (double)row[perm[p + 1]];
                }
                if (p < e) a0 += (double)row[perm[p]];
                if (on) *cell = old + (a0 + a1);
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads, 1) fused_pass_kernel(const PassParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = P.warps, nthreads = P.warps * 32;
    const int t_begin = P.cta_tile_start[blockIdx.x], t_end = P.cta_tile_start[blockIdx.x + 1];
    if (t_begin >= t_end) return;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);                                   // [W][S]
    CentreHeader* hdr = reinterpret_cast<CentreHeader*>(smem + P.off_hdr);
    float* table = reinterpret_cast<float*>(smem + P.off_table);
    int32_t* orig = reinterpret_cast<int32_t*>(smem + P.off_orig);
    int* wscr = reinterpret_cast<int*>(smem + P.off_scratch + (size_t)warp * P.scratch_bytes);
    unsigned char* perm = reinterpret_cast<unsigned char*>(wscr + P.k_max + 2);
    float* slots = reinterpret_cast<float*>(smem + P.off_slots) + (size_t)warp * P.slots_per_warp * P.slot_floats;
    const int S = P.slots_per_warp;

    if (lane == 0)
        for (int s = 0; s < S; ++s) mbar_init(&bars[warp * S + s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;  // bit s = parity the next wait on slot s expects

    // problem containing the first tile
    int p = 0;
    {
        int lo = 0, hi = P.n_problems - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.problems[mid].first_tile <= t_begin) lo = mid; else hi = mid - 1;
        }
        p = lo;
    }
    int tile = t_begin;
    while (tile < t_end) {
        const Problem pr = P.problems[p];
        const int p_end = min(t_end, pr.first_tile + pr.n_tiles);
        if (pr.done || p_end <= tile) {
            tile = max(tile, p_end);
            ++p;
            continue;
        }
        __syncthreads();  // nobody still reads the previous problem's table
        {
            const int* src = reinterpret_cast<const int*>(pr.hdr);
            int* dst = reinterpret_cast<int*>(hdr);
            for (int i = threadIdx.x; i < (int)(sizeof(CentreHeader) / 4); i += nthreads) dst[i] = src[i];
        }
        __syncthreads();
        const int n_slots = hdr->n_slots, n_chunks = hdr->n_chunks;
        {
            const float4* src = reinterpret_cast<const float4*>(pr.table);
            float4* dst = reinterpret_cast<float4*>(table);
            const int n4 = (hdr->table_floats + 3) >> 2;
            for (int i = threadIdx.x; i < n4; i += nthreads) dst[i] = src[i];
            for (int i = threadIdx.x; i < n_slots; i += nthreads) orig[i] = pr.orig[i];
        }
        __syncthreads();
        const int d = pr.d;
        const bool trivial = hdr->trivial != 0, force_exact = hdr->force_exact != 0;
        const float tau_c = (2.5f * (float)(d + 4)) * 5.9604645e-8f + 7.6293945e-6f;  // fp32 dot bound + 5 masked bits (2*2^-18)
        const float cc2 = 2.0f * hdr->cc_max;
        const uint32_t slot_bytes_tx = (uint32_t)d * (kSlotCells * 4);

        // partial tables of this (cta, problem, warp)
        const size_t tab = ((size_t)(blockIdx.x + p) * W + warp);
        double* my_sums = P.part_sums ? P.part_sums + tab * P.part_stride : nullptr;
        long long* my_counts = P.part_counts ? P.part_counts + tab * P.k_max : nullptr;
        long long changed = 0;

        const int first = tile + warp;
        const int count = first < p_end ? (p_end - first + W - 1) / W : 0;
        auto issue = [&](int i) {
            const int s = i % S;
            const int64_t cell0 = (int64_t)(first + i * W - pr.first_tile) * kSlotCells;
            float* dst = slots + (size_t)s * P.slot_floats;
            uint64_t* bar = &bars[warp * S + s];
            if (lane == 0) mbar_expect_tx(bar, slot_bytes_tx);
            __syncwarp();
            for (int j = lane; j < d; j += 32) bulk_g2s(dst + j * kRowStride, pr.x + (size_t)j * pr.ld + cell0, kSlotCells * 4, bar);
        };
        for (int i = 0; i < min(S, count); ++i) issue(i);

        for (int i = 0; i < count; ++i) {
            const int s = i % S;
            const int64_t cell0 = (int64_t)(first + i * W - pr.first_tile) * kSlotCells;
            const int64_t mycell = cell0 + 4 * lane;
            const float* slot = slots + (size_t)s * P.slot_floats;
            // previous labels (the reference compares against the previous assignment, Clusterer.cpp:247)
            int4 oldl = make_int4(0, 0, 0, 0);
            const bool full4 = mycell + 3 < pr.n;
            if ((P.compare_old || P.mode == kPassSums) && pr.labels) {
                if (full4) oldl = *reinterpret_cast<const int4*>(pr.labels + mycell);
                else {
                    if (mycell + 0 < pr.n) oldl.x = pr.labels[mycell + 0];
                    if (mycell + 1 < pr.n) oldl.y = pr.labels[mycell + 1];
                    if (mycell + 2 < pr.n) oldl.z = pr.labels[mycell + 2];
                }
            }
            mbar_wait(&bars[warp * S + s], (phase >> s) & 1u);
            phase ^= (1u << s);

            int lab[4] = {0, 0, 0, 0};
            if (P.mode == kPassSums) {
                lab[0] = oldl.x; lab[1] = oldl.y; lab[2] = oldl.z; lab[3] = oldl.w;
            } else if (!trivial) {
                float xx[4] = {0.f, 0.f, 0.f, 0.f};
                float best[4], second[4];
                int bchunk[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r = 0; r < 4; ++r) best[r] = second[r] = 3.3e38f;
                dist_chunk_dispatch<true>(hdr->chunk_kc[0], slot, table + hdr->chunk_off[0], d, lane, xx, best, second);
                for (int q = 1; q < n_chunks; ++q) {
                    float prev[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) prev[r] = best[r];
                    dist_chunk_dispatch<false>(hdr->chunk_kc[q], slot, table + hdr->chunk_off[q], d, lane, xx, best, second);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (best[r] < prev[r]) bchunk[r] = q;
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int a = hdr->chunk_base[bchunk[r]] + (int)(__float_as_uint(best[r]) & 31u);
                    const float tau = tau_c * (xx[r] + cc2);
                    const bool sure = (second[r] - best[r] > tau) && !force_exact;
                    if (sure) lab[r] = orig[a];
                    else if (mycell + r < pr.n) {
                        lab[r] = exact_label(slot, 4 * lane + r, d, pr.mu, orig, n_slots);
                        if (P.recheck_count) atomicAdd(P.recheck_count, 1ULL);
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r)
                if (mycell + r >= pr.n) lab[r] = -1;
            if (P.compare_old) {
                changed += (lab[0] >= 0 && lab[0] != oldl.x) + (lab[1] >= 0 && lab[1] != oldl.y) + (lab[2] >= 0 && lab[2] != oldl.z) +
                           (lab[3] >= 0 && lab[3] != oldl.w);
            }
            if (pr.labels && P.mode != kPassSums) {
                if (full4) *reinterpret_cast<int4*>(pr.labels + mycell) = make_int4(lab[0], lab[1], lab[2], lab[3]);
                else {
                    if (lab[0] >= 0) pr.labels[mycell + 0] = lab[0];
                    if (lab[1] >= 0) pr.labels[mycell + 1] = lab[1];
                    if (lab[2] >= 0) pr.labels[mycell + 2] = lab[2];
                }
            }
            if (P.mode != kPassAssign) slot_sums(slot, lab, d, pr.k, lane, wscr, perm, my_sums, my_counts);
            __syncwarp();
            if (i + S < count) issue(i + S);
        }
        if (P.part_changed) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
            if (lane == 0) P.part_changed[tab] = changed;
        }
        tile = p_end;
        ++p;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side: shared-memory plan + launch
// ---------------------------------------------------------------------------------------------------
constexpr size_t kBarBytes = 256;  // kWarpsPerCta * 3 slots * 8 B mbarriers, rounded up

int plan_pass(int d, int k_max, PassPlan* plan) {
    if (d < 1 || k_max < 1 || k_max > kMaxKC * kMaxChunks - 2 * kMaxChunks) return 0;
    const int slot_floats = d * kRowStride;
    const size_t slot_bytes = (size_t)slot_floats * 4;
    const int n_chunks = (k_max + kMaxKC - 1) / kMaxKC;
    const size_t table_bytes = ((size_t)(d + 1) * (k_max + 3 * n_chunks) * 4 + 127) & ~(size_t)127;
    const size_t orig_bytes = ((size_t)(k_max + 2 * n_chunks) * 4 + 127) & ~(size_t)127;
    const size_t scratch = (((size_t)(k_max + 2) * 4 + kSlotCells) + 127) & ~(size_t)127;
    const size_t hdr_bytes = (sizeof(CentreHeader) + 127) & ~(size_t)127;
    int W = kWarpsPerCta;
    for (; W >= 1; --W)
        if (kBarBytes + hdr_bytes + table_bytes + orig_bytes + (scratch + slot_bytes) * W <= (size_t)kSmemBudget) break;
    if (W < 1) return 0;  // caller reports the shape as unsupported
    const size_t fixed = kBarBytes + hdr_bytes + table_bytes + orig_bytes + scratch * W;
    int S = (int)(((size_t)kSmemBudget - fixed) / (slot_bytes * W));
    if (S > 3) S = 3;
    plan->warps = W;
    plan->slots_per_warp = S;
    plan->slot_floats = slot_floats;
    plan->off_hdr = (int)kBarBytes;
    plan->off_table = plan->off_hdr + (int)hdr_bytes;
    plan->off_orig = plan->off_table + (int)table_bytes;
    plan->off_scratch = plan->off_orig + (int)orig_bytes;
    plan->scratch_bytes = (int)scratch;
    plan->off_slots = plan->off_scratch + (int)(scratch * W);
    plan->smem_bytes = plan->off_slots + (size_t)S * W * slot_bytes;
    plan->table_floats_cap = (int)(table_bytes / 4);
    return 1;
}

int launch_fused_pass(const PassPlan& plan, PassParams P, int grid, cudaStream_t stream) {
    P.warps = plan.warps;
    P.slots_per_warp = plan.slots_per_warp;
    P.slot_floats = plan.slot_floats;
    P.off_hdr = plan.off_hdr;
    P.off_table = plan.off_table;
    P.off_orig = plan.off_orig;
    P.off_scratch = plan.off_scratch;
    P.scratch_bytes = plan.scratch_bytes;
    P.off_slots = plan.off_slots;
    static size_t configured = 0;
    if (plan.smem_bytes > configured) {
        DPR_CUDA(cudaFuncSetAttribute(fused_pass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        configured = kSmemBudget;
    }
    fused_pass_kernel<<<grid, plan.warps * 32, plan.smem_bytes, stream>>>(P);
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
