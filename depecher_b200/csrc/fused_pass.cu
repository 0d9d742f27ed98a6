// fused_pass.cu — the pass over the cells on the FP32 pipe: nearest-centre assignment fused with the per-cluster
// sums/counts.  The tensor-core pass (tc_pass.cu) takes the shapes that fit tensor memory; this kernel serves the
// others (more than 128 active centres, very wide tiles), the sums of given labels (kPassSums), and A/B measurements.
//
// Replaces Clusterer::allocate_clusters (/root/reference/src/Clusterer.cpp:34-87) and the scatter-add
// half of Clusterer::reevaluate_centers (:99-102) with ONE pass over the cells.
//
//   * X lives in HBM as fp32 tiles of 128 cells (common.cuh: x_index): a warp fills its "slot" (one tile in shared
//     memory) with ONE bulk-async copy (TMA, cp.async.bulk + mbarrier complete_tx); 8 warps per SM.
//   * Distance micro-kernel: each lane holds 4 cells (lane, lane+32, lane+64, lane+96) x up to 32 centres = 128
//     fp32 accumulators as packed pairs and issues FFMA2 (fma.rn.f32x2): s_k = ||c_k||^2 - 2 x.c_k.  One LDS.64
//     brings two markers of a cell, one broadcast LDS.128 brings 4 centres.  (Measured on B200: packed FFMA2 with
//     >= 32-wide centre tiles reaches 88-98 % of the FP32 pipe, scalar FFMA or narrower tiles lose 10-30 %:
//     profiles/microbench/r01_pipes_b200.log.)
//   * Argmin in registers with the centre index packed into the 5 low mantissa bits (FMNMX/FMNMX3);
//     cells whose two best fp32 scores are closer than a proven error bound are re-decided in fp64 exactly
//     as the reference does (sqrt of the left-to-right sum, strict <), so labels are bit-exact for
//     fp32-representable data.
//   * Lloyd mode: the warp counting-sorts its contributions by label (match.any), then lanes switch to marker
//     ownership, add each label's run in fp64 registers and issue one fp64 reduction per (label, marker) into the
//     warp's PRIVATE table.  Full mode contributes every cell (+x to its label); delta mode only the cells whose
//     label changed (+x to the new, -x to the old label).  Fixed tile->warp map + private tables (one lane per
//     address, program order) + fixed combine order => bit-reproducible sums.
#include "common.cuh"
#include "fused_pass.cuh"
#include "pass_device.cuh"

namespace dpr {


template <int KC>
__device__ __forceinline__ void dist_chunk(const float* __restrict__ slot, const float* __restrict__ tab, int groups, int lane,
                                           float (&best)[4], float (&second)[4]) {
    constexpr int NP = KC / 2;
    constexpr int KCP = (KC + 3) & ~3;
    float2 acc[4][NP];
    {
        const float2* cc = reinterpret_cast<const float2*>(tab);
#pragma unroll
        for (int t = 0; t < NP; ++t) {
            const float2 c = cc[t];
#pragma unroll
            for (int r = 0; r < 4; ++r) acc[r][t] = c;
        }
    }
    const float2* xs = reinterpret_cast<const float2*>(slot) + 2 * lane;   // this lane's cells: lane, lane+32, lane+64, lane+96
    const float4* cw = reinterpret_cast<const float4*>(tab + KCP);
    for (int h = 0; h < 2 * groups; ++h, cw += KCP / 2) {   // half a marker group per step keeps the x registers at 8
        float x[2][4];   // [marker of the half group][cell]
        const float2* xg = xs + (h >> 1) * (kGroupFloats / 2) + (h & 1);
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float2 v = xg[64 * r];
            x[0][r] = v.x; x[1][r] = v.y;
        }
#pragma unroll
        for (int jj = 0; jj < 2; ++jj) {
#pragma unroll
            for (int q = 0; q < KCP / 4; ++q) {
                const float4 c = cw[jj * (KCP / 4) + q];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float2 xr = make_float2(x[jj][r], x[jj][r]);
                    acc[r][2 * q] = __ffma2_rn(xr, make_float2(c.x, c.y), acc[r][2 * q]);
                    if (2 * q + 1 < NP) acc[r][2 * q + 1] = __ffma2_rn(xr, make_float2(c.z, c.w), acc[r][2 * q + 1]);
                }
            }
        }
    }
    // two smallest packed scores per cell: merge the sorted pair (lo, hi) of each accumulator pair
#pragma unroll
    for (int t = 0; t < NP; ++t) {
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const float a = pack_idx(acc[r][t].x, 2 * t);
            const float b = pack_idx(acc[r][t].y, 2 * t + 1);
            const float lo = fminf(a, b), hi = fmaxf(a, b);
            second[r] = fmin3(fmaxf(best[r], lo), second[r], hi);
            best[r] = fminf(best[r], lo);
        }
    }
}

__device__ __forceinline__ void dist_chunk_dispatch(int kc, const float* slot, const float* tab, int groups, int lane, float (&best)[4],
                                                    float (&second)[4]) {
    switch (kc) {
#define DPR_CASE(KC)                                                \
    case KC:                                                        \
        dist_chunk<KC>(slot, tab, groups, lane, best, second);       \
        break;
        DPR_CASE(2) DPR_CASE(4) DPR_CASE(6) DPR_CASE(8) DPR_CASE(10) DPR_CASE(12) DPR_CASE(14) DPR_CASE(16)
        DPR_CASE(18) DPR_CASE(20) DPR_CASE(22) DPR_CASE(24) DPR_CASE(26) DPR_CASE(28) DPR_CASE(30) DPR_CASE(32)
#undef DPR_CASE
        default: break;
    }
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kThreads, 1) fused_pass_kernel(const PassParams P) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int W = P.warps, nthreads = P.warps * 32;
    const int t_begin = P.cta_tile_start[blockIdx.x], t_end = P.cta_tile_start[blockIdx.x + 1];
    if (t_begin >= t_end) return;

    uint64_t* bars = reinterpret_cast<uint64_t*>(smem);  // [W][S]
    CentreHeader* hdr = reinterpret_cast<CentreHeader*>(smem + P.off_hdr);
    float* table = reinterpret_cast<float*>(smem + P.off_table);
    int32_t* orig = reinterpret_cast<int32_t*>(smem + P.off_orig);
    int* wscr = reinterpret_cast<int*>(smem + P.off_scratch + (size_t)warp * P.scratch_bytes);
    unsigned char* perm = reinterpret_cast<unsigned char*>(wscr + 2 * P.k_max + 2);
    float* slots = reinterpret_cast<float*>(smem + P.off_slots) + (size_t)warp * P.slots_per_warp * P.slot_floats;
    const int S = P.slots_per_warp;

    if (lane == 0)
        for (int s = 0; s < S; ++s) mbar_init(&bars[warp * S + s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    uint32_t phase = 0;  // bit s = parity the next wait on slot s expects

    // this warp's private tables (global, L2-resident)
    const size_t wtab = (size_t)blockIdx.x * W + warp;
    double* my_sums = P.warp_sums ? P.warp_sums + wtab * P.part_stride : nullptr;
    long long* my_counts = P.warp_counts ? P.warp_counts + wtab * P.k_max : nullptr;

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
        const int chunk0_kc = hdr->chunk_kc[0], chunk0_off = hdr->chunk_off[0];
        const int groups = tile_groups(d);
        const uint32_t slot_bytes_tx = (uint32_t)(tile_floats(d) * 4);
        const bool delta = MODE == kPassLloyd && pr.delta;
        const bool need_old = (P.compare_old || MODE == kPassSums || delta) && pr.labels;
        long long changed = 0;

        const int first = tile + warp;
        const int count = first < p_end ? (p_end - first + W - 1) / W : 0;
        // ONE bulk-async copy per slot: a tile is contiguous (and already swizzled) in HBM.  Tiles of this warp are W apart.
        const size_t tile_fl = tile_floats(d);
        const float* next_src = pr.x + (size_t)(first - pr.first_tile) * tile_fl;   // tile the next issue() loads
        int issued = 0, s_issue = 0;
        auto issue = [&]() {
            if (lane == 0) {
                mbar_expect_tx(&bars[warp * S + s_issue], slot_bytes_tx);
                bulk_g2s(smem_u32(slots + (size_t)s_issue * P.slot_floats), next_src, slot_bytes_tx, smem_u32(&bars[warp * S + s_issue]));
                // the tile this warp will want after that one: pull it into L2 now, so its bulk copy (issued only when the
                // slot is free again) pays L2 latency instead of HBM latency
                if (issued + 1 < count) bulk_prefetch_l2(next_src + (size_t)W * tile_fl, slot_bytes_tx);
            }
            next_src += (size_t)W * tile_fl;
            ++issued;
            s_issue = (s_issue + 1 == S) ? 0 : s_issue + 1;
        };
        for (int i = 0; i < min(S, count); ++i) issue();

        int s = 0;   // slot of tile i: advances round-robin like s_issue (no runtime modulo in the loop)
        int64_t cell0 = (int64_t)(first - pr.first_tile) * kSlotCells;
        const float* xx_ptr = pr.tile_xx + (first - pr.first_tile);
        for (int i = 0; i < count; ++i, cell0 += (int64_t)W * kSlotCells, xx_ptr += W) {
            const int64_t mycell = cell0 + lane;   // this lane's cells: mycell + 32 * r
            const float* slot = slots + (size_t)s * P.slot_floats;
            // previous labels (the reference compares against the previous assignment, Clusterer.cpp:247)
            int oldl[4] = {0, 0, 0, 0};
            if (need_old) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (mycell + 32 * r < pr.n) oldl[r] = pr.labels[mycell + 32 * r];
            }
            const float xx_tile = *xx_ptr;   // bound of ||x||^2 over this slot's cells
            mbar_wait(&bars[warp * S + s], (phase >> s) & 1u);
            phase ^= (1u << s);

            int lab[4] = {0, 0, 0, 0};
            if (MODE == kPassSums) {
#pragma unroll
                for (int r = 0; r < 4; ++r) lab[r] = oldl[r];
            } else if (!trivial) {
                float best[4], second[4];
                int bchunk[4] = {0, 0, 0, 0};
#pragma unroll
                for (int r = 0; r < 4; ++r) best[r] = second[r] = 3.3e38f;
                dist_chunk_dispatch(chunk0_kc, slot, table + chunk0_off, groups, lane, best, second);
                for (int q = 1; q < n_chunks; ++q) {
                    float prev[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r) prev[r] = best[r];
                    dist_chunk_dispatch(hdr->chunk_kc[q], slot, table + hdr->chunk_off[q], groups, lane, best, second);
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        if (best[r] < prev[r]) bchunk[r] = q;
                }
                const float tau = tau_c * (xx_tile + cc2);
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int a = hdr->chunk_base[bchunk[r]] + (int)(__float_as_uint(best[r]) & 31u);
                    const bool sure = (second[r] - best[r] > tau) && !force_exact;
                    lab[r] = orig[a < n_slots ? a : 0];
                    unsigned redo = __ballot_sync(0xffffffffu, !sure && (mycell + 32 * r < pr.n));
                    if (redo && P.recheck_count && lane == 0) atomicAdd(P.recheck_count, (unsigned long long)__popc(redo));
                    while (redo) {   // the whole warp re-decides each doubtful cell in fp64
                        const int src = __ffs(redo) - 1;
                        redo &= redo - 1;
                        const int v = exact_label_warp(slot, src + 32 * r, d, pr.mu, d, orig, n_slots, lane);
                        if (lane == src) lab[r] = v;
                    }
                }
            }
            bool ch[4];
            int nch = 0;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (mycell + 32 * r >= pr.n) lab[r] = -1;
                ch[r] = lab[r] >= 0 && lab[r] != oldl[r];
                nch += (int)ch[r];
            }
            if (P.compare_old) changed += nch;
            if (pr.labels && MODE != kPassSums) {
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    if (lab[r] >= 0) pr.labels[mycell + 32 * r] = lab[r];
            }
            if (MODE != kPassAssign) {
                unsigned mch[4];
                int moved = 0;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    mch[r] = __ballot_sync(0xffffffffu, ch[r]);
                    moved += __popc(mch[r]);
                }
                if (delta && moved <= 24) {
                    // few cells moved: no sort, each moved cell adds its row to the new label's sums and takes it off the old one's
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        unsigned m = mch[r];
                        while (m) {
                            const int src = __ffs(m) - 1;
                            m &= m - 1;
                            const int newL = __shfl_sync(0xffffffffu, lab[r], src);
                            const int oldL = __shfl_sync(0xffffffffu, oldl[r], src);
                            move_one(slot, d, lane, my_sums, my_counts, src + 32 * r, newL, oldL);
                        }
                    }
                } else {
                    int key[8], cell[8];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        cell[r] = cell[r + 4] = lane + 32 * r;
                        if (delta) {
                            key[r] = ch[r] ? 2 * lab[r] : -1;
                            key[r + 4] = ch[r] ? 2 * oldl[r] + 1 : -1;
                        } else {
                            key[r] = lab[r] >= 0 ? 2 * lab[r] : -1;
                            key[r + 4] = -1;
                        }
                    }
                    slot_contributions(slot, key, cell, d, pr.k, lane, wscr, perm, my_sums, my_counts);
                }
            }
            __syncwarp();
            if (issued < count) issue();
            s = (s + 1 == S) ? 0 : s + 1;
        }
        // ---- combine the warps' tables of this (CTA, problem) in fixed order; reset them for the next range
        if (P.cta_tables) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
            long long* s_changed = reinterpret_cast<long long*>(smem + P.off_hdr);  // header no longer needed
            __threadfence();   // the warps' table updates are visible to the whole CTA
            __syncthreads();
            if (lane == 0) s_changed[warp] = changed;
            __syncthreads();
            const int64_t stride = P.part_stride;
            const int64_t entries = stride + P.k_max + 1;
            double* out = P.cta_tables + (size_t)(blockIdx.x + p) * entries;
            const size_t wbase = (size_t)blockIdx.x * W;
            for (int64_t e = threadIdx.x; e < entries; e += nthreads) {
                double acc = 0.0;
                if (e < stride) {
                    if (P.warp_sums) {
                        double v[kWarpsPerCta];   // all loads in flight first, then the fixed-order sum
#pragma unroll
                        for (int w = 0; w < kWarpsPerCta; ++w) v[w] = w < W ? __ldcg(P.warp_sums + (wbase + w) * stride + e) : 0.0;
#pragma unroll
                        for (int w = 0; w < kWarpsPerCta; ++w) {
                            acc += v[w];
                            if (w < W && v[w] != 0.0) __stcg(P.warp_sums + (wbase + w) * stride + e, 0.0);
                        }
                    }
                } else if (e < stride + P.k_max) {
                    if (P.warp_counts) {
                        long long v[kWarpsPerCta];
#pragma unroll
                        for (int w = 0; w < kWarpsPerCta; ++w) v[w] = w < W ? __ldcg(P.warp_counts + (wbase + w) * P.k_max + (e - stride)) : 0LL;
                        long long t = 0;
#pragma unroll
                        for (int w = 0; w < kWarpsPerCta; ++w) {
                            t += v[w];
                            if (w < W && v[w] != 0) __stcg(P.warp_counts + (wbase + w) * P.k_max + (e - stride), 0LL);
                        }
                        acc = (double)t;
                    }
                } else {
                    long long t = 0;
                    for (int w = 0; w < W; ++w) t += s_changed[w];
                    acc = (double)t;
                }
                out[e] = acc;
            }
        }
        tile = p_end;
        ++p;
    }
}

// ---------------------------------------------------------------------------------------------------
// host side: shared-memory plan + launch
// ---------------------------------------------------------------------------------------------------
constexpr size_t kBarBytes = 512;  // kWarpsPerCta * 3 slots * 8 B mbarriers, rounded up

int plan_pass(int d, int k_max, PassPlan* plan) {
    if (d < 1 || k_max < 1 || k_max > kMaxKC * kMaxChunks - 2 * kMaxChunks) return 0;
    const int slot_floats = (int)((tile_floats(d) + 31) & ~(size_t)31);   // slots stay 128 B aligned
    const size_t slot_bytes = (size_t)slot_floats * 4;
    const int n_chunks = (k_max + kMaxKC - 1) / kMaxKC;
    const size_t table_bytes = ((size_t)(4 * tile_groups(d) + 1) * (k_max + 3 * n_chunks) * 4 + 127) & ~(size_t)127;
    const size_t orig_bytes = ((size_t)(k_max + 2 * n_chunks) * 4 + 127) & ~(size_t)127;
    const size_t scratch = (((size_t)(2 * k_max + 2) * 4 + 2 * kSlotCells) + 127) & ~(size_t)127;
    const size_t hdr_bytes = std::max<size_t>((sizeof(CentreHeader) + 127) & ~(size_t)127, 8 * kWarpsPerCta);
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
    static bool configured = false;
    if (!configured) {
        DPR_CUDA(cudaFuncSetAttribute(fused_pass_kernel<kPassAssign>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        DPR_CUDA(cudaFuncSetAttribute(fused_pass_kernel<kPassLloyd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        DPR_CUDA(cudaFuncSetAttribute(fused_pass_kernel<kPassSums>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        configured = true;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx().time_passes) {
        DPR_CUDA(cudaEventCreate(&e0));
        DPR_CUDA(cudaEventCreate(&e1));
        DPR_CUDA(cudaEventRecord(e0, stream));
    }
    // one instance per mode: the assignment-only pass carries no sums code (fewer registers spilled, smaller I-cache footprint)
    if (P.mode == kPassAssign) fused_pass_kernel<kPassAssign><<<grid, plan.warps * 32, plan.smem_bytes, stream>>>(P);
    else if (P.mode == kPassLloyd) fused_pass_kernel<kPassLloyd><<<grid, plan.warps * 32, plan.smem_bytes, stream>>>(P);
    else fused_pass_kernel<kPassSums><<<grid, plan.warps * 32, plan.smem_bytes, stream>>>(P);
    count_launch();
    if (e0) {
        DPR_CUDA(cudaEventRecord(e1, stream));
        ctx().pass_events.emplace_back(e0, e1);
    }
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
