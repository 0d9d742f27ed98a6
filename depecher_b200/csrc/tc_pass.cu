// tc_pass.cu — the hot kernel on the 5th-generation tensor cores: nearest-centre assignment fused with the
// per-cluster sums/counts, one persistent CTA per SM.
//
// Same contract as fused_pass.cu (it replaces Clusterer::allocate_clusters, /root/reference/src/Clusterer.cpp:34-87,
// and the scatter-add half of Clusterer::reevaluate_centers, :99-102) for the shapes that fit tensor memory; the
// distance step  s_k = ||c_k||^2 - 2 x.c_k  becomes D[128 cells x N centres] = X * (-2 C)^T on tcgen05.mma.
//
//   * A tile of 128 cells lands in shared memory with ONE bulk-async copy (TMA) — the resident layout in HBM is
//     already the K-major canonical operand layout (common.cuh).  A producer warp keeps a ring of tiles in flight.
//   * fp32 accuracy from tf32 products: x = xh + xl, -2c = ch + cl with the high parts the tf32 truncations
//     (kind::tf32 ignores the low 13 mantissa bits, so the raw fp32 word serves as xh).  Each of the four consumer
//     warps owns 32 cells = 32 TMEM lanes: a thread reads its cell's markers (LDS.128), computes xl, and parks xh and
//     xl in tensor memory (tcgen05.st) as the A operands.  One thread then issues xl*ch + xh*cl + xh*ch
//     (3 * ceil(d/8) tcgen05.mma, A from TMEM, B from shared memory) into a double-buffered TMEM accumulator.
//     Measured (profiles/microbench/tc_probe.cu): |error| <= 1e-6 * sum|x_j c_j|, the same order as an fp32 FMA chain.
//   * While the tensor core works on tile i, the consumer warps run the epilogue of tile i-1: tcgen05.ld brings a
//     thread its cell's N scores, the two smallest are kept with the centre index packed into the 5 low mantissa
//     bits, and a cell whose two best scores are closer than the error bound is re-decided in fp64 exactly as the
//     reference does — labels stay bit-exact.  Lloyd mode then adds the cell to the warp's private fp64 tables
//     (pass_device.cuh), exactly like the FFMA kernel.
#include "common.cuh"
#include "fused_pass.cuh"
#include "pass_device.cuh"

namespace dpr {

constexpr int kTcGroups = 2;                      // consumer groups; group g takes the CTA's tiles g, g + G, ...
constexpr int kTcWarps = 4 * kTcGroups;           // consumer warps: a group = 4 warps x 32 TMEM lanes = the 128 cells of a tile
constexpr int kTcThreads = (kTcWarps + 2) * 32;   // + the warp that issues the products + the producer warp
constexpr int kTcMaxStages = 8;

// ---------------------------------------------------------------------------------------------------
// tcgen05 PTX
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(kTcWarps * 32) : "memory"); }   // all consumer warps
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: 8 rows x 16 B core matrices; lbo = bytes between the two 16 B halves of a K=8 step,
// sbo = bytes between 8-row blocks
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float4& a, const float4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(__float_as_uint(a.x)),
                 "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)),
                 "r"(__float_as_uint(b.y)), "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[32], int o) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[o + 0]), "=r"(v[o + 1]), "=r"(v[o + 2]), "=r"(v[o + 3]), "=r"(v[o + 4]), "=r"(v[o + 5]), "=r"(v[o + 6]), "=r"(v[o + 7]),
          "=r"(v[o + 8]), "=r"(v[o + 9]), "=r"(v[o + 10]), "=r"(v[o + 11]), "=r"(v[o + 12]), "=r"(v[o + 13]), "=r"(v[o + 14]), "=r"(v[o + 15])
        : "r"(taddr)
        : "memory");
}
// one lane of a converged warp (the operands of the single-thread tcgen05 / bulk-copy instructions then stay
// warp-uniform for the compiler: no per-lane waterfall loop around them)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}\n"
        : "+r"(pred)
        : "r"(0xffffffffu));
    return pred != 0;
}
__device__ __forceinline__ float tf32_low(float v) { return v - __uint_as_float(__float_as_uint(v) & 0xffffe000u); }

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
template <int MODE>
__global__ void __launch_bounds__(kTcThreads, 1) tc_pass_kernel(const PassParams P, const TcPlan T) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_begin = P.cta_tile_start[blockIdx.x], t_end = P.cta_tile_start[blockIdx.x + 1];
    if (t_begin >= t_end) return;

    const int S = T.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [S] tile landed
    uint64_t* empty = full + kTcMaxStages;                // [S] tile no longer needed (one arrival per warp of the consuming group)
    uint64_t* a_ready = empty + kTcMaxStages;             // [G] the group's A operands are in tensor memory (one arrival per warp)
    uint64_t* mma_done = a_ready + 4;                     // [G] products into the group's accumulator finished
    uint64_t* tables_ready = mma_done + 4;                // the problem's centre tables are in shared memory
    uint32_t* tmem_word = reinterpret_cast<uint32_t*>(tables_ready + 1);
    CentreHeader* hdr = reinterpret_cast<CentreHeader*>(smem + T.off_hdr);
    float* cc_s = reinterpret_cast<float*>(smem + T.off_cc);
    int32_t* orig = reinterpret_cast<int32_t*>(smem + T.off_orig);
    float* bmat = reinterpret_cast<float*>(smem + T.off_b);
    float* stages = reinterpret_cast<float*>(smem + T.off_stages);
    const int d = T.d, groups = tile_groups(d), g8 = T.g8;
    const size_t tile_fl = tile_floats(d);
    const uint32_t tile_bytes = (uint32_t)(tile_fl * 4);
    const int npad = T.npad, d8 = 4 * g8;
    const uint32_t buf_cols = (uint32_t)(npad + 2 * d8);   // per group: accumulator, xh, xl

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        for (int g = 0; g < kTcGroups; ++g) {
            mbar_init(&a_ready[g], 4);
            mbar_init(&mma_done[g], 1);
        }
        mbar_init(tables_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)), "r"((uint32_t)T.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (g8 > groups)   // d % 8 in 1..4: the second half of the last K=8 step is a marker group no tile brings — zero, once
        for (int s = 0; s < S; ++s)
            for (int i = threadIdx.x; i < kGroupFloats; i += kTcThreads) stages[(size_t)s * T.stage_floats + (size_t)groups * kGroupFloats + i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_word;

    int p = 0;
    {
        int lo = 0, hi = P.n_problems - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.problems[mid].first_tile <= t_begin) lo = mid; else hi = mid - 1;
        }
        p = lo;
    }

    // ================================ producer warp ================================
    if (warp == kTcWarps + 1) {
        int tile = t_begin, s = 0, issued = 0;
        uint32_t phase = 0;
        while (tile < t_end) {
            const Problem* pr = P.problems + p;
            const int first_tile = pr->first_tile, p_end = min(t_end, first_tile + pr->n_tiles);
            if (pr->done || p_end <= tile) {
                tile = max(tile, p_end);
                ++p;
                continue;
            }
            const float* src = pr->x + (size_t)(tile - first_tile) * tile_fl;
            for (; tile < p_end; ++tile, src += tile_fl) {
                if (issued >= S) {
                    mbar_wait(&empty[s], (phase >> s) & 1u);
                    phase ^= 1u << s;
                }
                if (elect_one()) {
                    mbar_expect_tx(&full[s], tile_bytes);
                    bulk_g2s(smem_u32(stages + (size_t)s * T.stage_floats), src, tile_bytes, smem_u32(&full[s]));
                }
                __syncwarp();
                ++issued;
                s = (s + 1 == S) ? 0 : s + 1;
            }
            ++p;
        }
        return;
    }

    // ================================ the warp that issues the products ================================
    if (warp == kTcWarps) {
        int tile = t_begin;
        uint32_t it = 0, n_prob = 0;
        const uint32_t bh_addr = smem_u32(bmat), bl_addr = bh_addr + (uint32_t)(g8 * npad * 16);
        const uint32_t lbo = (uint32_t)npad * 16u;
        const uint64_t desc_h0 = umma_desc(bh_addr, lbo, 128), desc_l0 = umma_desc(bl_addr, lbo, 128);
        const uint64_t desc_step = (uint64_t)((2u * lbo) >> 4);   // one K=8 step = two marker groups further
        while (tile < t_end) {
            const Problem* pr = P.problems + p;
            const int p_end = min(t_end, pr->first_tile + pr->n_tiles);
            if (pr->done || p_end <= tile) {
                tile = max(tile, p_end);
                ++p;
                continue;
            }
            mbar_wait(tables_ready, n_prob & 1u);
            ++n_prob;
            const int N = hdr->tc_n;
            const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
            for (; tile < p_end; ++tile, ++it) {
                const uint32_t g = it % kTcGroups, k = it / kTcGroups;
                mbar_wait(&a_ready[g], k & 1u);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t acc = tmem + g * buf_cols, ah = acc + (uint32_t)npad, al = ah + (uint32_t)d8;
                    for (int s = 0; s < g8 / 2; ++s) {   // small terms first
                        umma_tf32_ts(acc, al + 8 * s, desc_h0 + s * desc_step, idesc, s > 0 ? 1u : 0u);
                        umma_tf32_ts(acc, ah + 8 * s, desc_l0 + s * desc_step, idesc, 1);
                    }
                    for (int s = 0; s < g8 / 2; ++s) umma_tf32_ts(acc, ah + 8 * s, desc_h0 + s * desc_step, idesc, 1);
                    umma_commit(&mma_done[g]);
                }
                __syncwarp();
            }
            ++p;
        }
        return;
    }

    // ================================ consumer warps ================================
    const int grp = warp >> 2, wq = warp & 3;   // group, and the TMEM lane quarter this warp may touch
    int* wscr = reinterpret_cast<int*>(smem + T.off_scratch + (size_t)warp * T.scratch_bytes);
    unsigned char* perm = reinterpret_cast<unsigned char*>(wscr + 2 * P.k_max + 2);
    const size_t wtab = (size_t)blockIdx.x * kTcWarps + warp;
    double* my_sums = P.warp_sums ? P.warp_sums + wtab * P.part_stride : nullptr;
    long long* my_counts = P.warp_counts ? P.warp_counts + wtab * P.k_max : nullptr;
    const int mycell_in_tile = wq * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)grp * buf_cols;
    const uint32_t a_hi = lane_base + (uint32_t)npad, a_lo = a_hi + (uint32_t)d8;

    uint32_t it = 0;        // tiles this CTA has walked past (all groups)
    uint32_t k_mine = 0;    // tiles this group has taken
    int tile = t_begin;
    while (tile < t_end) {
        const Problem pr = P.problems[p];
        const int p_end = min(t_end, pr.first_tile + pr.n_tiles);
        if (pr.done || p_end <= tile) {
            tile = max(tile, p_end);
            ++p;
            continue;
        }
        consumer_barrier();   // nobody still reads the previous problem's tables (every group has seen its last products complete)
        {
            const int* src = reinterpret_cast<const int*>(pr.hdr);
            int* dst = reinterpret_cast<int*>(hdr);
            for (int i = threadIdx.x; i < (int)(sizeof(CentreHeader) / 4); i += kTcWarps * 32) dst[i] = src[i];
            const float4* bsrc = reinterpret_cast<const float4*>(pr.tc_table + npad);
            float4* bdst = reinterpret_cast<float4*>(bmat);
            for (int i = threadIdx.x; i < 2 * g8 * npad; i += kTcWarps * 32) bdst[i] = bsrc[i];
            for (int i = threadIdx.x; i < npad; i += kTcWarps * 32) {
                cc_s[i] = pr.tc_table[i];
                orig[i] = pr.tc_orig[i];
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the B operand was written through the generic proxy
        consumer_barrier();
        if (threadIdx.x == 0) mbar_arrive(tables_ready);
        const int N = hdr->tc_n, n_act = hdr->k_act;
        const bool trivial = hdr->trivial != 0, force_exact = hdr->force_exact != 0;
        // error bound of a score: tf32 splits 3 * 2^-20, accumulation 2^-22 per product step, both relative to
        // sum|x_j c_j| <= (||x||^2 + ||c||^2) / 2 and doubled by the factor -2; cc rounding and the final add 3 * 2^-24;
        // twice that for a difference of two scores, plus the 5 masked index bits (2 * 2^-18) — see DESIGN.md §3
        const float eps = 3.f * 9.5367432e-7f + (float)(3 * (d8 / 8)) * 2.3841858e-7f + 3.f * 5.9604645e-8f;
        const float tau_c = 2.f * eps + 7.6293945e-6f;
        const float cc2 = 2.0f * hdr->cc_max;
        const bool delta = MODE == kPassLloyd && pr.delta;
        const bool need_old = (P.compare_old || delta) && pr.labels;
        long long changed = 0;
        const int count = p_end - tile;

        // this group's tiles of the range: j with (it + j) % G == grp
        for (int j = (int)((grp + kTcGroups - it % kTcGroups) % kTcGroups); j < count; j += kTcGroups, ++k_mine) {
            const uint32_t git = it + (uint32_t)j;             // ordinal of the tile in the CTA's walk: stage and phase of its landing
            const int s = (int)(git % (uint32_t)S);
            const float* slot = stages + (size_t)s * T.stage_floats;
            const int64_t mycell = (int64_t)(tile + j - pr.first_tile) * kSlotCells + mycell_in_tile;
            const bool live = mycell < pr.n;
            int oldl = 0;
            if (need_old && live) oldl = pr.labels[mycell];
            const float xx_tile = pr.tile_xx[tile + j - pr.first_tile];
            mbar_wait(&full[s], (git / (uint32_t)S) & 1u);
            // ---- split into tf32 high/low parts, parked in tensor memory as the A operands
            {
                const float4* st4 = reinterpret_cast<const float4*>(slot) + mycell_in_tile;
                for (int g = 0; g < g8; g += 2) {
                    const float4 v0 = st4[g * (kGroupFloats / 4)], v1 = st4[(g + 1) * (kGroupFloats / 4)];
                    tmem_st8(a_hi + 4 * g, v0, v1);
                    tmem_st8(a_lo + 4 * g, make_float4(tf32_low(v0.x), tf32_low(v0.y), tf32_low(v0.z), tf32_low(v0.w)),
                             make_float4(tf32_low(v1.x), tf32_low(v1.y), tf32_low(v1.z), tf32_low(v1.w)));
                }
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[grp]);
            }
            // ---- scores -> label (+ sums)
            mbar_wait(&mma_done[grp], k_mine & 1u);
            tc_fence_after();
            int lab = 0;
            if (!trivial) {
                float best = 3.3e38f, second = 3.3e38f;
                int bchunk = 0;
                for (int n0 = 0; n0 < N; n0 += 32) {
                    uint32_t v[32];
                    tmem_ld16(lane_base + n0, v, 0);
                    if (n0 + 16 < N) tmem_ld16(lane_base + n0 + 16, v, 16);
                    else {
#pragma unroll
                        for (int t = 16; t < 32; ++t) v[t] = 0u;   // columns beyond N: cc is +huge there
                    }
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    const float prev = best;
                    const float4* cc4 = reinterpret_cast<const float4*>(cc_s + n0);
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float4 c = cc4[q];
                        const float s0 = pack_idx(c.x + __uint_as_float(v[4 * q + 0]), 4 * q + 0);
                        const float s1 = pack_idx(c.y + __uint_as_float(v[4 * q + 1]), 4 * q + 1);
                        const float s2 = pack_idx(c.z + __uint_as_float(v[4 * q + 2]), 4 * q + 2);
                        const float s3 = pack_idx(c.w + __uint_as_float(v[4 * q + 3]), 4 * q + 3);
                        const float lo0 = fminf(s0, s1), hi0 = fmaxf(s0, s1);
                        second = fmin3(fmaxf(best, lo0), second, hi0);
                        best = fminf(best, lo0);
                        const float lo1 = fminf(s2, s3), hi1 = fmaxf(s2, s3);
                        second = fmin3(fmaxf(best, lo1), second, hi1);
                        best = fminf(best, lo1);
                    }
                    if (best < prev) bchunk = n0;
                }
                const float tau = tau_c * (xx_tile + cc2) + 1e-30f;
                const int a = bchunk + (int)(__float_as_uint(best) & 31u);
                const bool sure = (second - best > tau) && !force_exact;
                lab = orig[a < n_act ? a : 0];
                unsigned redo = __ballot_sync(0xffffffffu, !sure && live);
                if (redo && P.recheck_count && lane == 0) atomicAdd(P.recheck_count, (unsigned long long)__popc(redo));
                while (redo) {   // the whole warp re-decides each doubtful cell in fp64
                    const int src = __ffs(redo) - 1;
                    redo &= redo - 1;
                    const int v = exact_label_warp(slot, wq * 32 + src, d, pr.mu, orig, n_act, lane);
                    if (lane == src) lab = v;
                }
            }
            if (!live) lab = -1;
            const bool ch = lab >= 0 && lab != oldl;
            if (P.compare_old) changed += (int)ch;
            if (pr.labels && live) pr.labels[mycell] = lab;
            if (MODE == kPassLloyd) {
                const unsigned mch = __ballot_sync(0xffffffffu, ch);
                if (delta && __popc(mch) <= 8) {
                    unsigned m = mch;
                    while (m) {
                        const int src = __ffs(m) - 1;
                        m &= m - 1;
                        const int newL = __shfl_sync(0xffffffffu, lab, src);
                        const int oldL = __shfl_sync(0xffffffffu, oldl, src);
                        move_one(slot, d, lane, my_sums, my_counts, wq * 32 + src, newL, oldL);
                    }
                } else if (delta) {
                    const int key[2] = {ch ? 2 * lab : -1, ch ? 2 * oldl + 1 : -1};
                    const int cell[2] = {mycell_in_tile, mycell_in_tile};
                    slot_contributions<2>(slot, key, cell, d, pr.k, lane, wscr, perm, my_sums, my_counts);
                } else {
                    const int key[1] = {lab >= 0 ? 2 * lab : -1};
                    const int cell[1] = {mycell_in_tile};
                    slot_contributions<1>(slot, key, cell, d, pr.k, lane, wscr, perm, my_sums, my_counts);
                }
            }
            tc_fence_before();   // this warp's tcgen05.ld of the accumulator are complete before the next products overwrite it
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
        }
        it += (uint32_t)count;
        // ---- combine the warps' tables of this (CTA, problem) in fixed order; reset them for the next range
        if (P.cta_tables) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);
            long long* s_changed = reinterpret_cast<long long*>(smem + T.off_cc);   // the centre tables are no longer needed ...
            __threadfence();   // the warps' table updates are visible to the whole CTA
            consumer_barrier();   // ... once every warp is past its last tile
            if (lane == 0) s_changed[warp] = changed;
            consumer_barrier();
            const int64_t stride = P.part_stride;
            const int64_t entries = stride + P.k_max + 1;
            double* out = P.cta_tables + (size_t)(blockIdx.x + p) * entries;
            const size_t wbase = (size_t)blockIdx.x * kTcWarps;
            for (int64_t e = threadIdx.x; e < entries; e += kTcWarps * 32) {
                double acc = 0.0;
                if (e < stride) {
                    if (P.warp_sums) {
                        double v[kTcWarps];   // all loads in flight first, then the fixed-order sum
#pragma unroll
                        for (int w = 0; w < kTcWarps; ++w) v[w] = __ldcg(P.warp_sums + (wbase + w) * stride + e);
#pragma unroll
                        for (int w = 0; w < kTcWarps; ++w) {
                            acc += v[w];
                            if (v[w] != 0.0) __stcg(P.warp_sums + (wbase + w) * stride + e, 0.0);
                        }
                    }
                } else if (e < stride + P.k_max) {
                    if (P.warp_counts) {
                        long long t = 0;
#pragma unroll
                        for (int w = 0; w < kTcWarps; ++w) {
                            const long long v = __ldcg(P.warp_counts + (wbase + w) * P.k_max + (e - stride));
                            t += v;
                            if (v != 0) __stcg(P.warp_counts + (wbase + w) * P.k_max + (e - stride), 0LL);
                        }
                        acc = (double)t;
                    }
                } else {
                    long long t = 0;
                    for (int w = 0; w < kTcWarps; ++w) t += s_changed[w];
                    acc = (double)t;
                }
                out[e] = acc;
            }
        }
        tile = p_end;
        ++p;
    }
    tc_fence_before();
    consumer_barrier();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)T.tmem_cols));
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int plan_tc_pass(int d, int k_max, TcPlan* plan) {
    if (d < 1 || k_max < 1) return 0;
    const int npad = (k_max + 31) & ~31;
    const int g8 = 2 * ((d + 7) >> 3), d8 = 4 * g8;
    if (npad > 256) return 0;
    int cols = kTcGroups * (npad + 2 * d8), tmem_cols = 32;
    while (tmem_cols < cols) tmem_cols <<= 1;
    if (tmem_cols > 512) return 0;
    const size_t stage_bytes = (((size_t)g8 * kGroupFloats * 4) + 127) & ~(size_t)127;
    const size_t bars = 256;
    const size_t hdr_bytes = (sizeof(CentreHeader) + 127) & ~(size_t)127;
    const size_t cc_bytes = (size_t)npad * 4, orig_bytes = (size_t)npad * 4;
    const size_t scratch = (((size_t)(2 * k_max + 2) * 4 + 64) + 127) & ~(size_t)127;
    const size_t b_bytes = (size_t)2 * g8 * npad * 16;
    const size_t fixed = bars + hdr_bytes + cc_bytes + orig_bytes + scratch * kTcWarps + b_bytes;
    if (fixed + 3 * stage_bytes > (size_t)kSmemBudget) return 0;
    int S = (int)(((size_t)kSmemBudget - fixed) / stage_bytes);
    if (S > kTcMaxStages) S = kTcMaxStages;
    plan->d = d;
    plan->g8 = g8;
    plan->npad = npad;
    plan->tmem_cols = tmem_cols;
    plan->stages = S;
    plan->stage_floats = (int)(stage_bytes / 4);
    plan->off_hdr = (int)bars;
    plan->off_cc = plan->off_hdr + (int)hdr_bytes;
    plan->off_orig = plan->off_cc + (int)cc_bytes;
    plan->off_scratch = plan->off_orig + (int)orig_bytes;
    plan->scratch_bytes = (int)scratch;
    plan->off_b = plan->off_scratch + (int)(scratch * kTcWarps);
    plan->off_stages = plan->off_b + (int)b_bytes;
    plan->smem_bytes = (size_t)plan->off_stages + (size_t)S * stage_bytes;
    plan->table_floats = (size_t)npad + (size_t)2 * g8 * npad * 4;
    return 1;
}

int launch_tc_pass(const TcPlan& plan, const PassParams& P, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        DPR_CUDA(cudaFuncSetAttribute(tc_pass_kernel<kPassAssign>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        DPR_CUDA(cudaFuncSetAttribute(tc_pass_kernel<kPassLloyd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        configured = true;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx().time_passes) {
        DPR_CUDA(cudaEventCreate(&e0));
        DPR_CUDA(cudaEventCreate(&e1));
        DPR_CUDA(cudaEventRecord(e0, stream));
    }
    if (P.mode == kPassAssign) tc_pass_kernel<kPassAssign><<<grid, kTcThreads, plan.smem_bytes, stream>>>(P, plan);
    else tc_pass_kernel<kPassLloyd><<<grid, kTcThreads, plan.smem_bytes, stream>>>(P, plan);
    count_launch();
    if (e0) {
        DPR_CUDA(cudaEventRecord(e1, stream));
        ctx().pass_events.emplace_back(e0, e1);
    }
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
