// tc_pass.cu — the hot kernel on the 5th-generation tensor cores: nearest-centre assignment fused with the
// per-cluster sums/counts, one persistent CTA per SM.
//
// Same contract as fused_pass.cu (it replaces Clusterer::allocate_clusters, /root/reference/src/Clusterer.cpp:34-87,
// and the scatter-add half of Clusterer::reevaluate_centers, :99-102) for the shapes that fit tensor memory; the
// distance step  s_k = ||c_k||^2 - 2 x.c_k  becomes D[128 cells x N centres] = X * (-2 C)^T on tcgen05.mma.
//
//   * A tile of 128 cells lands in shared memory with ONE bulk-async copy (TMA): the resident layout in HBM is the
//     K-major canonical operand layout (common.cuh).  A producer warp keeps a ring of tiles in flight.
//   * fp32 accuracy from fp16 products: cells and centres are scaled by powers of two into fp16's range, then split
//     v = hi + lo with both parts fp16 (|v - hi - lo| <= 2^-22 |v| + 2^-25).  A consumer thread owns one cell = one
//     TMEM lane: it reads the cell's markers (LDS.128), splits them and parks xh, xl in tensor memory (tcgen05.st)
//     as A operands.  The centre table (centres.cu) stacks the high parts on top of the low parts in ONE B operand,
//     so the issuer warp needs two products per K=16 step: xh*[ch|cl] and xl*ch, accumulated in fp32 in TMEM.
//     Measured (profiles/microbench/tc_probe.cu): |error| <= 1e-6 * sum|x_j c_j|, the order of an fp32 FMA chain.
//   * G = 2..4 consumer groups of four warps (as many as tensor memory holds) take the CTA's tiles round-robin: while
//     one group waits for its products the others split or reduce.  tcgen05.ld brings a thread its cell's scores;
//     the two smallest are kept with the centre index packed into the 4 low mantissa bits, and a cell whose two
//     best scores are closer than the error bound is re-decided in fp64 exactly as the reference does — labels stay
//     bit-exact.  Lloyd mode: the first iteration of a fit adds every cell to fp64 tables in a fixed order (group tables in
//     shared memory); every later iteration only RECORDS the cells whose label changed in per-warp lists, and
//     delta_sums_kernel (below) adds their rows (+ to the new label, - to the old one) afterwards.
//   * Assignment-only sessions: up to four centre sets that score the same cells (the penalty search's scoring step) ride
//     on one landed tile (bundles): split once, multiplied with one set after the other.
//   * Single-thread instructions (tcgen05.mma, cp.async.bulk) are issued under elect.sync from converged warps so
//     that their operands stay in uniform registers.
//
// Build with -DDPR_TC_PROFILE to print per-phase cycle counts of one CTA (how the numbers in profiles/ were taken);
// the environment variable DPR_TC_GROUPS (2..4) overrides the number of consumer groups (tuning aid).
#include "common.cuh"
#include "fused_pass.cuh"
#include "pass_device.cuh"
#include "tc_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>

namespace dpr {

// G consumer groups per CTA; group g takes the CTA's tiles g, g + G, ...  A group = 4 warps x 32 TMEM lanes = the
// 128 cells of a tile.  Two more warps: the one that issues the products and the producer.
constexpr int kTcMaxGroups = 4;
constexpr int kTcMaxStages = 8;

// Lloyd full mode when the group tables do not fit in shared memory (large k * d): this warp's 32 cells into its private table in
// global memory by counting sort (out of line: rare)
__device__ __noinline__ void full_warp_sums(const float* __restrict__ slot, int d, int k, int lane, int lab, int mycell_in_tile, int* wscr,
                                            unsigned char* perm, double* my_sums, long long* my_counts) {
    const int key[1] = {lab >= 0 ? 2 * lab : -1};
    const int cell[1] = {mycell_in_tile};
    slot_contributions<1>(slot, key, cell, d, k, lane, wscr, perm, my_sums, my_counts);
}

// Lloyd full mode, one tile: warp w of the group takes the labels w, w + 4, ...; it finds a label's cells by ballot over the
// labels the four warps published and adds their rows (lane = marker, fp64 registers, highest cell first: a fixed order) into the
// group's table in shared memory.  `lb` is double-buffered by the caller (a warp can be one tile ahead of a slow neighbour at most).
// The counts come from the warps' own cells (match.any: one shared-memory integer add per label and warp, which commutes).
// The loop is bound by instruction issue (F2F + DADD per element, plus the bit scan): the address of a cell's second marker block
// is the first one's plus a constant, and a set bit costs FLO + clear + LEA.
__device__ __noinline__ void full_tile_sums(const float* __restrict__ slot, unsigned char* lb, double* gsum, long long* gcnt, int d, int k,
                                            int grp, int wq, int lane, int mycell_in_tile, int lab) {
    lb[mycell_in_tile] = (unsigned char)(lab >= 0 ? lab : 255);
    {   // this warp's 32 cells: one add per label present among them
        const unsigned peers = __match_any_sync(0xffffffffu, lab);
        if (lab >= 0 && (peers & ((1u << lane) - 1u)) == 0u) atomicAdd(reinterpret_cast<unsigned long long*>(gcnt) + lab, (unsigned long long)__popc(peers));
    }
    asm volatile("bar.sync %0, 128;" ::"r"(2 + grp) : "memory");
    const int l0 = lb[lane], l1 = lb[32 + lane], l2 = lb[64 + lane], l3 = lb[96 + lane];
    constexpr int kBlock1 = 8 * kGroupFloats;   // floats between a cell's marker j and marker j + 32 (8 marker groups)
    for (int jb = 0; jb < d; jb += 64) {
        const int j0 = jb + lane, j1 = jb + 32 + lane;
        const bool on0 = j0 < d, on1 = j1 < d;
        const float* r0 = slot + x_slot_index(0, on0 ? j0 : 0);   // marker j0 of cell 0 (clamped: the load is unconditional)
        for (int L = wq; L < k; L += 4) {
            const unsigned m[4] = {__ballot_sync(0xffffffffu, l0 == L), __ballot_sync(0xffffffffu, l1 == L), __ballot_sync(0xffffffffu, l2 == L),
                                   __ballot_sync(0xffffffffu, l3 == L)};
            if (!(m[0] | m[1] | m[2] | m[3])) continue;
            double a0 = 0.0, a1 = 0.0;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                unsigned mm = m[q];
                const uint32_t rq = smem_u32(r0 + 128 * q);   // cells 32 q .. (a shared-memory address: one LEA per cell)
                while (mm) {
                    int c;   // the highest set bit (bfind: one FLO; 31 - __clz costs the compiler two more instructions)
                    asm("bfind.u32 %0, %1;" : "=r"(c) : "r"(mm));
                    mm ^= 1u << c;
                    float x0, x1 = 0.f;
                    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x0) : "r"(rq + 16u * (uint32_t)c));
                    if (on1) asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(x1) : "r"(rq + 16u * (uint32_t)c), "n"(4 * kBlock1));
                    a0 += (double)x0;
                    a1 += (double)x1;   // (lanes without a second marker add 0)
                }
            }
            if (on0) gsum[L * d + j0] += a0;
            if (on1) gsum[L * d + j1] += a1;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------------------------------
struct SetRef {          // what the consumers need of problem p + s of a bundle besides its tables
    int32_t* labels;
    const double* mu;    // exact centres (global memory; used when they are not staged in shared memory)
};
template <int MODE, int kTcGroups, bool kMulti>
__global__ void __launch_bounds__((4 * kTcGroups + 2) * 32, 1) tc_pass_kernel(const PassParams P, const TcPlan T) {
    constexpr int kTcWarps = 4 * kTcGroups;
    constexpr bool kLloyd = MODE == kPassLloyd || MODE == kPassLloydDelta;
    constexpr bool kDeltaOnly = MODE == kPassLloydDelta;   // no problem of the launch needs full sums: that code is not in this instance
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    asm volatile("griddepcontrol.launch_dependents;");
    // (the tile ranges are not written by the kernels of a step — only by the host and by partition_kernel, a plain launch that has completed
    // before this grid is placed: safe to read before the wait, through L2)
    const int t_begin = __ldcg(P.cta_tile_start + blockIdx.x), t_end = __ldcg(P.cta_tile_start + blockIdx.x + 1);
    if (t_begin >= t_end) return;

    const int S = T.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [S] tile landed
    uint64_t* empty = full + kTcMaxStages;                // [S] tile no longer needed (one arrival per warp of the consuming group)
    uint64_t* a_ready = empty + kTcMaxStages;             // [G] the group's A operands are in tensor memory (one arrival per warp)
    uint64_t* mma_done = a_ready + kTcMaxGroups;                     // [G] products into the group's accumulator finished
    uint64_t* tables_ready = mma_done + kTcMaxGroups;                // the problem's centre tables are in shared memory
    uint64_t* acc_free = tables_ready + 1;                           // [G] bundles: the group has read its accumulator, the next set's products may overwrite it
    uint32_t* tmem_word = reinterpret_cast<uint32_t*>(acc_free + kTcMaxGroups);
    SetRef* set_ref = reinterpret_cast<SetRef*>(tmem_word + 2);      // [kTcMaxSets] label / exact-centre pointers of the bundle's problems
    float4* keyc = reinterpret_cast<float4*>(smem + 304);            // [kTcMaxSets] per centre set: near-tie bound in key units tk_a, tk_b, tk_c (see below), un-scaling factor
    // kMulti: several centre sets per tile (bundles; assignment-only sessions).  A separate instance: the single-set pass keeps its registers
    CentreHeader* hdr = reinterpret_cast<CentreHeader*>(smem + T.off_hdr);
    float* bmat = reinterpret_cast<float*>(smem + T.off_b);
    float* stages = reinterpret_cast<float*>(smem + T.off_stages);
    double* mu_s = reinterpret_cast<double*>(smem + T.off_mu);   // the problem's exact centres, for the fp64 re-decisions (when they fit)
    // Lloyd tables of a consumer group (when they fit): sums[k_max * d] doubles, counts[k_max], and the per-tile hand-over area
    const int d = T.d, groups = tile_groups(d), g16 = T.g8;   // g16: groups of 8 markers, whole K=16 steps
    const size_t tile_fl = tile_floats(d);
    const uint32_t tile_bytes = (uint32_t)(tile_fl * 4);
    const int npad = T.npad, acols = 4 * g16;              // 32-bit TMEM columns of one fp16 A operand (2 markers each)
    const uint32_t buf_cols = (uint32_t)(2 * npad + 2 * acols);   // per group: accumulator (high | low centre parts), xh, xl

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        for (int g = 0; g < kTcGroups; ++g) {
            mbar_init(&a_ready[g], 4);
            mbar_init(&mma_done[g], 1);
            mbar_init(&acc_free[g], 4);
        }
        mbar_init(tables_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)), "r"((uint32_t)T.tmem_cols));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (T.gtab && !kDeltaOnly)
        for (int i = threadIdx.x; i < (int)(T.gtab_bytes / 8) * kTcGroups; i += blockDim.x) reinterpret_cast<double*>(smem + T.off_gtab)[i] = 0.0;
    if (groups & 1)   // the split reads marker groups in pairs: the partner of an odd last group is never written by a copy — zero, once
        for (int s = 0; s < S; ++s)
            for (int i = threadIdx.x; i < kGroupFloats; i += blockDim.x) stages[(size_t)s * T.stage_floats + (size_t)groups * kGroupFloats + i] = 0.f;
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_word;
    // the set-up above (barriers, tensor-memory allocation, zero padding) runs while the previous kernel of the step is still finishing
    // on other SMs; everything below reads what it wrote
    asm volatile("griddepcontrol.wait;" ::: "memory");

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
                    mbar_wait_hint(&empty[s], (phase >> s) & 1u, T.wait_hint);
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
        uint32_t it = 0, n_prob = 0, ar_phase = 0, free_phase = 0;   // (phase bits: one per group)
        const uint32_t lbo = (uint32_t)npad * 32u;   // bytes between 8-marker groups of the (2*npad)-row B operand
        const uint64_t desc_b0 = umma_desc(smem_u32(bmat), lbo, 128);
        const uint64_t desc_step = (uint64_t)((2u * lbo) >> 4);   // one K=16 step = two marker groups further
        while (tile < t_end) {
            const Problem* pr = P.problems + p;
            const int p_end = min(t_end, pr->first_tile + pr->n_tiles);
            if (pr->done || p_end <= tile) {
                tile = max(tile, p_end);
                ++p;
                continue;
            }
            mbar_wait_hint(tables_ready, n_prob & 1u, T.wait_hint);
            ++n_prob;
            const uint32_t idesc0 = (1u << 4) | ((uint32_t)(128 >> 4) << 24);           // D fp32, A/B fp16, both K-major, M = 128
            if constexpr (!kMulti) {
                const int N = reinterpret_cast<const CentreHeader*>(smem + T.off_hdr)->tc_n;
                const uint32_t idesc = idesc0 | ((uint32_t)(N >> 3) << 17);   // N columns: xh * ch, xh * cl and xl * ch all accumulate into [0, N)
                const uint64_t desc_l0 = desc_b0 + (uint64_t)npad;            // the low parts: npad rows of 16 B further in every marker group
                for (; tile < p_end; ++tile, ++it) {
                    const uint32_t g = it % kTcGroups;
                    mbar_wait_hint(&a_ready[g], (ar_phase >> g) & 1u, T.wait_hint);
                    ar_phase ^= 1u << g;
                    tc_fence_after();
                    if (elect_one()) {
                        const uint32_t acc = tmem + g * buf_cols, ah = acc + (uint32_t)(2 * npad), al = ah + (uint32_t)acols;
                        for (int s = 0; s < g16 / 2; ++s) {
                            umma_f16_ts(acc, ah + 8 * s, desc_b0 + s * desc_step, idesc, s > 0 ? 1u : 0u);   // xh * ch
                            umma_f16_ts(acc, ah + 8 * s, desc_l0 + s * desc_step, idesc, 1);                  // xh * cl
                            umma_f16_ts(acc, al + 8 * s, desc_b0 + s * desc_step, idesc, 1);                  // xl * ch
                        }
                        umma_commit(&mma_done[g]);
                    }
                    __syncwarp();
                }
            } else {
                // bundles: every tile is multiplied with NS centre sets, one after the other into the same accumulator.  A group is
                // ready for its next products when its A operands are in place (first set of a tile) or when it has taken the
                // previous set's scores out of the accumulator; the groups reach these points in no particular order, so the warp
                // polls them in turn instead of waiting for one.
                const int NS = max(1, pr->bundle);
                const int count = p_end - tile;
                int left[kTcGroups], q_next[kTcGroups], events = 0;
#pragma unroll
                for (int g = 0; g < kTcGroups; ++g) {
                    const int j0 = (int)((g + kTcGroups - it % kTcGroups) % kTcGroups);   // the group's first tile of the range
                    left[g] = j0 < count ? (count - j0 + kTcGroups - 1) / kTcGroups : 0;
                    q_next[g] = 0;
                    events += left[g] * NS;
                }
                const uint64_t desc_set = (uint64_t)((uint32_t)T.set_bytes >> 4);   // the next set's B operand
                while (events > 0) {
                    bool any = false;
#pragma unroll
                    for (int g = 0; g < kTcGroups; ++g) {
                        if (left[g] == 0) continue;
                        const int q = q_next[g];
                        uint64_t* bar = q == 0 ? &a_ready[g] : &acc_free[g];
                        const uint32_t par = ((q == 0 ? ar_phase : free_phase) >> g) & 1u;
                        if (!mbar_test(bar, par)) continue;
                        if (q == 0) ar_phase ^= 1u << g; else free_phase ^= 1u << g;
                        tc_fence_after();
                        const int N = reinterpret_cast<const CentreHeader*>(smem + T.off_hdr + q * T.set_bytes)->tc_n;
                        const uint32_t idesc = idesc0 | ((uint32_t)(N >> 3) << 17);
                        if (elect_one()) {
                            const uint32_t acc = tmem + g * buf_cols, ah = acc + (uint32_t)(2 * npad), al = ah + (uint32_t)acols;
                            const uint64_t db = desc_b0 + q * desc_set, dl = db + (uint64_t)npad;
                            for (int s = 0; s < g16 / 2; ++s) {
                                umma_f16_ts(acc, ah + 8 * s, db + s * desc_step, idesc, s > 0 ? 1u : 0u);
                                umma_f16_ts(acc, ah + 8 * s, dl + s * desc_step, idesc, 1);
                                umma_f16_ts(acc, al + 8 * s, db + s * desc_step, idesc, 1);
                            }
                            umma_commit(&mma_done[g]);
                        }
                        __syncwarp();
                        if (q + 1 == NS) {
                            q_next[g] = 0;
                            --left[g];
                        } else {
                            q_next[g] = q + 1;
                        }
                        --events;
                        any = true;
                    }
                    if (!any) __nanosleep(40);
                }
                it += (uint32_t)count;
                tile = p_end;
            }
            ++p;
        }
        return;
    }

    // ================================ consumer warps ================================
    const int grp = warp >> 2, wq = warp & 3;   // group, and the TMEM lane quarter this warp may touch
    const int mycell_in_tile = wq * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)grp * buf_cols;
    const uint32_t a_hi = lane_base + (uint32_t)(2 * npad), a_lo = a_hi + (uint32_t)acols;
    const int steps8 = (d + 7) >> 3;   // 8-marker steps that hold data
    {   // zero padding of the last K=16 step, in both operands of this group's buffer
        const uint32_t z[4] = {0u, 0u, 0u, 0u};
        for (int c = 4 * steps8; c < acols; c += 4) {
            tmem_st4(a_hi + c, z);
            tmem_st4(a_lo + c, z);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }

#ifdef DPR_TC_PROFILE
    long long prof_t[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, prof_acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    long long prof_n = 0;
#define DPR_PROF(i) do { prof_t[i] = clock64(); if (i > 0) prof_acc[i] += prof_t[i] - prof_t[i - 1]; if (i == 8) ++prof_n; } while (0)
#else
#define DPR_PROF(i)
#endif
    uint32_t it = 0;        // tiles this CTA has walked past (all groups)
    uint32_t list_cursor = 0;   // entries this warp has appended to its list of moved cells (all problems of the CTA's range)
    int s_mine = grp % S;   // stage and landing parity of this group's next tile (its ordinal is grp, grp + G, ...)
    uint32_t par_mine = (uint32_t)(grp / S) & 1u;
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
        consumer_barrier<kTcWarps * 32>();   // nobody still reads the previous problem's tables (every group has seen its last products complete)
        const int NS = kMulti ? max(1, pr.bundle) : 1;   // centre sets that share this range's cells (problems p .. p + NS - 1)
        for (int q = 0; q < NS; ++q) {
            const Problem* pq = P.problems + p + q;
            unsigned char* base = smem + (size_t)q * T.set_bytes;
            const float* tab = q ? pq->tc_table : pr.tc_table;
            const int32_t* torig = q ? pq->tc_orig : pr.tc_orig;
            const double* mu_g = q ? pq->mu : pr.mu;
            const int kq = q ? pq->k : pr.k;
            const int* src = reinterpret_cast<const int*>(q ? pq->hdr : pr.hdr);
            int* dst = reinterpret_cast<int*>(base + T.off_hdr);
            for (int i = threadIdx.x; i < (int)(sizeof(CentreHeader) / 4); i += kTcWarps * 32) dst[i] = src[i];
            const float4* bsrc = reinterpret_cast<const float4*>(tab + npad);
            float4* bdst = reinterpret_cast<float4*>(base + T.off_b);
            for (int i = threadIdx.x; i < 2 * g16 * npad; i += kTcWarps * 32) bdst[i] = bsrc[i];   // g16 groups x 2*npad rows x 16 B
            float* cq = reinterpret_cast<float*>(base + T.off_cc);
            int32_t* oq = reinterpret_cast<int32_t*>(base + T.off_orig);
            const CentreHeader* hq = q ? pq->hdr : pr.hdr;
            // sortable integer keys (tc_device.cuh): every score of the set is bounded by ||x||^2 + 2 ||c||^2
            const float ccm = hq->cc_max;
            const float qq = key_quantum(*pr.xx_max + 2.0f * ccm);
            const float big_m = 12582912.f * qq;   // 1.5 * 2^23 * q  (0 without a usable scale: every cell is then decided exactly)
            const int kact_q = hq->k_act;
            for (int i = threadIdx.x; i < npad; i += kTcWarps * 32) {
                cq[i] = i < kact_q ? tab[i] + big_m : big_m + kKeyPad * qq;   // (padding columns: above every real score, inside the binade of M)
                oq[i] = torig[i];
            }
            if (threadIdx.x == 32 * q) {
                // near-tie bound of a pair of scores, in key units (q / 16): tau = tau_a (X + cc_max) + tau_b (X + 2 cc_max) + abs1 sqrt(X) + abs2 + 3 q,
                // X = the tile's bound of ||x||^2.  tau_a: the fp16 splits leave at most 3 * 2^-22 (proved: |v - hi - lo| <= 2^-22 |v| for
                // round-to-nearest parts, three cross terms); the fp32 accumulation inside the tensor core is allowed 2^-21 per product —
                // measured on this part at <= 2^-24 per step, also for same-sign terms (profiles/microbench/tc_probe_f16.cu) — with 3 products
                // per K = 16 step into one column; both relative to sum|x_j c_j| <= (||x||^2 + ||c||^2) / 2, doubled by the factor -2 and again
                // for a difference of two scores.  tau_b: the un-scaling fma and cc + M (relative to |score| <= ||x||^2 + 2 ||c||^2).  abs1 /
                // abs2: the absolute part of the split error (fp16 subnormals).  3 q: the keys' own resolution.
                const float tau_a = 2.f * (3.f * 2.3841858e-7f + (float)(3 * (g16 / 2) + 1) * 4.7683716e-7f);
                const float tau_b = 2.f * (4.f * 5.9604645e-8f);
                const float cc2 = 2.0f * ccm, iq16 = qq > 0.f ? 16.f / qq : 0.f;
                float4 kc;
                kc.x = (tau_a + tau_b) * iq16;
                kc.y = hq->tc_abs1 * iq16;
                kc.z = (hq->force_exact || !(qq > 0.f)) ? INFINITY : (tau_a * 0.5f * cc2 + tau_b * cc2 + hq->tc_abs2 + 3.f * qq) * iq16;
                kc.w = hq->tc_unscale;
                keyc[q] = kc;
            }
            if (T.mu_in_smem) {
                double* mq = reinterpret_cast<double*>(base + T.off_mu);
                for (int i = threadIdx.x; i < kq * d; i += kTcWarps * 32) mq[(i / d) * (d | 1) + i % d] = mu_g[i];   // odd row stride: lanes = rows hit different banks
            }
            if (kMulti && threadIdx.x == 0) {
                set_ref[q].labels = q ? pq->labels : pr.labels;
                set_ref[q].mu = mu_g;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the B operands were written through the generic proxy
        consumer_barrier<kTcWarps * 32>();
        if (threadIdx.x == 0) mbar_arrive(tables_ready);
        int N = hdr->tc_n, n_act = hdr->k_act;
        const double* mu_exact = T.mu_in_smem ? mu_s : pr.mu;
        const int mu_ld = T.mu_in_smem ? (d | 1) : d;
        bool trivial = hdr->trivial != 0;
        const float scale_x = hdr->tc_scale_x;   // (the same for every set of a bundle: it depends on the cells only)
        float4 kc = keyc[0];                      // tk_a, tk_b, tk_c, un-scaling factor of the current set
        uint32_t set_off = 0;   // byte offset of the current set's tables (bundles)
        int32_t* labels_q = pr.labels;
        const bool delta = kLloyd && (kDeltaOnly || pr.delta);   // (a session that lets finalize set it has allocated the lists)
        if (delta && lane == 0)   // where this range's entries begin: parked in the warp's scratch (free in delta mode) until the range ends
            *reinterpret_cast<uint32_t*>(smem + T.off_scratch + (size_t)warp * T.scratch_bytes) = list_cursor;
        const bool need_old = (P.compare_old || delta) && pr.labels;   // (bundles: every member has labels when the leader has)
        int changed = 0;   // per thread: at most one per tile
        const int count = p_end - tile;

        // this group's tiles of the range: j with (it + j) % G == grp
        for (int j = (int)((grp + kTcGroups - it % kTcGroups) % kTcGroups); j < count; j += kTcGroups, ++k_mine) {
            const int s = s_mine;
            const uint32_t full_par = par_mine;
            s_mine += kTcGroups;
            while (s_mine >= S) {
                s_mine -= S;
                par_mine ^= 1u;
            }
            const float* slot = stages + (size_t)s * T.stage_floats;
            const int64_t mycell = (int64_t)(tile + j - pr.first_tile) * kSlotCells + mycell_in_tile;
            const bool live = mycell < pr.n;
            const float xx_tile = pr.tile_xx[tile + j - pr.first_tile];
            int oldl_early = 0;   // one set per tile: the previous label is fetched here, behind the split
            if (!kMulti && need_old && live) oldl_early = labels_q[mycell];
            DPR_PROF(0);
            mbar_wait_hint(&full[s], full_par, T.wait_hint);
            DPR_PROF(1);
            // ---- split into tf32 high/low parts, parked in tensor memory as the A operands
            {
                const float4* st4 = reinterpret_cast<const float4*>(slot) + mycell_in_tile;
                if (scale_x == 1.f) split_tile<false>(st4, steps8, 1.f, a_hi, a_lo);
                else split_tile<true>(st4, steps8, scale_x, a_hi, a_lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[grp]);
            }
            DPR_PROF(2);
            // ---- scores -> label (+ sums), once per centre set of the bundle
            int lab = 0, oldl = 0;
            bool ch = false;
#pragma unroll 1
            for (int q = 0; q < NS; ++q) {
            if (kMulti && NS > 1) {   // this set's tables and constants
                const unsigned char* base = smem + (size_t)q * T.set_bytes;
                const CentreHeader* h = reinterpret_cast<const CentreHeader*>(base + T.off_hdr);
                N = h->tc_n;
                n_act = h->k_act;
                trivial = h->trivial != 0;
                kc = keyc[q];
                set_off = (uint32_t)q * (uint32_t)T.set_bytes;
                mu_exact = T.mu_in_smem ? reinterpret_cast<const double*>(base + T.off_mu) : set_ref[q].mu;
                labels_q = set_ref[q].labels;
                if (q > 0) ++k_mine;   // (the loop header counts one completion of mma_done per tile)
            }
            oldl = oldl_early;
            if (kMulti && need_old && live) oldl = labels_q[mycell];
            mbar_wait_hint(&mma_done[grp], k_mine & 1u, T.wait_hint);
            tc_fence_after();
            DPR_PROF(3);
            lab = 0;
            if (!trivial) {
                int best = 0x7fffffff, second = 0x7fffffff, bchunk = 0;
                for (int n0 = 0; n0 < N; n0 += 16) {   // N is a multiple of 16
                    uint32_t v[16];
                    tmem_ld16_raw(lane_base + n0, v);
                    tmem_wait_ld16(v);
                    DPR_PROF(4);
                    int b, sec;
                    top2_block16(v, reinterpret_cast<const float4*>(smem + set_off + T.off_cc) + (n0 >> 2), kc.w, T.key_mul, b, sec);
                    if (b < best) bchunk = n0;   // (strict: an equal key of a later block never displaces an earlier one)
                    second = __vimin3_s32(max(best, b), second, sec);
                    best = min(best, b);
                }
                const float tk = fmaf(kc.x, xx_tile, fmaf(kc.y, sqrtf(xx_tile), kc.z));
                const int a = bchunk + (best & 15);
                const bool sure = (float)(second - best) > tk;
                const int32_t* orig_q = reinterpret_cast<const int32_t*>(smem + set_off + T.off_orig);
                lab = orig_q[a < n_act ? a : 0];
                DPR_PROF(5);
                unsigned redo = __ballot_sync(0xffffffffu, !sure && live);
                if (redo && P.recheck_count && lane == 0) atomicAdd(P.recheck_count, (unsigned long long)__popc(redo));
                while (redo) {   // the whole warp re-decides each doubtful cell in fp64
                    const int src = __ffs(redo) - 1;
                    redo &= redo - 1;
                    const int v = exact_label_warp(slot, wq * 32 + src, d, mu_exact, mu_ld, orig_q, n_act, lane);
                    if (lane == src) lab = v;
                }
            }
            DPR_PROF(6);
            if (!live) lab = -1;
            ch = lab >= 0 && lab != oldl;
            if (P.compare_old) changed += (int)ch;
            if (labels_q && live) labels_q[mycell] = lab;
            if (kMulti && q + 1 < NS) {   // the accumulator may take the next set's products
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&acc_free[grp]);
            }
            }   // sets
            DPR_PROF(7);
            if (kLloyd && !kDeltaOnly && T.gtab && !delta) {
                // every cell contributes: the four warps of the group publish their labels, meet at the group's barrier, and update
                // the group's table in shared memory with plain read-modify-writes in a fixed order (no atomics, reproducible)
                double* gsum = reinterpret_cast<double*>(smem + T.off_gtab + (size_t)grp * T.gtab_bytes);   // this group's table
                unsigned char* glab = smem + T.off_glist + (size_t)grp * T.glist_bytes;                        // [2][128] labels of a whole tile
                full_tile_sums(slot, glab + (k_mine & 1u) * 128, gsum, reinterpret_cast<long long*>(gsum + P.part_stride), d, pr.k, grp, wq, lane,
                               mycell_in_tile, lab);
            } else if (kLloyd && delta) {
                // only the cells whose label changed contribute (+x to the new label, -x to the old one): the warp appends them to
                // its own list, in tile and lane order; delta_sums_kernel gathers their rows afterwards.  Nothing in this kernel
                // waits on a moved cell.
                const unsigned mch = __ballot_sync(0xffffffffu, ch);
                if (mch) {   // (addresses formed only here: nothing about the lists stays live across tiles but the cursor)
                    const size_t wl = (size_t)blockIdx.x * kTcWarps + warp;
                    uint2* my_list = P.move_lists + wl * P.move_cap;
                    const uint32_t at = list_cursor + (uint32_t)__popc(mch & ((1u << lane) - 1u));
                    if (ch) {
                        if ((int64_t)at < P.move_cap) my_list[at] = make_uint2((uint32_t)mycell, (uint32_t)lab | ((uint32_t)oldl << 16));
                        else *P.move_overflow = 1;
                    }
                    list_cursor += (uint32_t)__popc(mch);
                }
            } else if (kLloyd && !kDeltaOnly) {
                // every cell contributes but the group tables do not fit in shared memory: this warp's private table in global memory
                const size_t wtab = (size_t)blockIdx.x * kTcWarps + warp;
                int* wscr = reinterpret_cast<int*>(smem + T.off_scratch + (size_t)warp * T.scratch_bytes);
                unsigned char* perm = reinterpret_cast<unsigned char*>(wscr + 2 * P.k_max + 2);
                full_warp_sums(slot, d, pr.k, lane, lab, mycell_in_tile, wscr, perm, P.warp_sums + wtab * P.part_stride, P.warp_counts + wtab * P.k_max);
            }
            tc_fence_before();   // this warp's tcgen05.ld of the accumulator are complete before the next products overwrite it
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[s]);
            DPR_PROF(8);
        }
        it += (uint32_t)count;
        // ---- combine the warps' tables of this (CTA, problem) in fixed order; reset them for the next range
        if (P.cta_tables) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) changed += __shfl_xor_sync(0xffffffffu, changed, o);   // <= 32 * tiles of the range: fits an int
            long long* s_changed = reinterpret_cast<long long*>(smem + T.off_cc);   // the centre tables are no longer needed ...
            __threadfence();   // the warps' table updates are visible to the whole CTA
            consumer_barrier<kTcWarps * 32>();   // ... once every warp is past its last tile
            if (lane == 0) s_changed[warp] = (long long)changed;
            consumer_barrier<kTcWarps * 32>();
            const int64_t stride = P.part_stride;
            const int64_t entries = stride + P.k_max + 1;
            double* out = P.cta_tables + (size_t)(blockIdx.x + p) * entries;
            if (kDeltaOnly || delta) {
                // sums and counts of this (CTA, problem) come from delta_sums_kernel: hand it the list lengths
                if (lane == 0)
                    P.move_counts[(size_t)(blockIdx.x + p) * kTcWarps + warp] =
                        (int32_t)(list_cursor - *reinterpret_cast<const uint32_t*>(smem + T.off_scratch + (size_t)warp * T.scratch_bytes));
                if (threadIdx.x == 0) {
                    long long t = 0;
                    for (int w = 0; w < kTcWarps; ++w) t += s_changed[w];
                    out[entries - 1] = (double)t;
                }
            } else if constexpr (!kDeltaOnly) {
                const size_t wbase = (size_t)blockIdx.x * kTcWarps;
                const bool wt = !T.gtab && P.warp_sums;   // the per-warp tables in global memory were in use
                for (int64_t e = threadIdx.x; e < entries; e += kTcWarps * 32) {
                    double acc = 0.0;
                    if (e < stride) {
                        if (wt) {
                            double v[kTcWarps];   // all loads in flight first, then the fixed-order sum
#pragma unroll
                            for (int w = 0; w < kTcWarps; ++w) v[w] = __ldcg(P.warp_sums + (wbase + w) * stride + e);
#pragma unroll
                            for (int w = 0; w < kTcWarps; ++w) {
                                acc += v[w];
                                if (v[w] != 0.0) __stcg(P.warp_sums + (wbase + w) * stride + e, 0.0);
                            }
                        }
                        if (T.gtab)
                            for (int g = 0; g < kTcGroups; ++g) {
                                double* t = reinterpret_cast<double*>(smem + T.off_gtab + (size_t)g * T.gtab_bytes) + e;
                                acc += *t;
                                *t = 0.0;
                            }
                    } else if (e < stride + P.k_max) {
                        long long t = 0;
                        if (wt) {
#pragma unroll
                            for (int w = 0; w < kTcWarps; ++w) {
                                const long long v = __ldcg(P.warp_counts + (wbase + w) * P.k_max + (e - stride));
                                t += v;
                                if (v != 0) __stcg(P.warp_counts + (wbase + w) * P.k_max + (e - stride), 0LL);
                            }
                        }
                        if (T.gtab)
                            for (int g = 0; g < kTcGroups; ++g) {
                                long long* c = reinterpret_cast<long long*>(reinterpret_cast<double*>(smem + T.off_gtab + (size_t)g * T.gtab_bytes) + stride) + (e - stride);
                                t += *c;
                                *c = 0;
                            }
                        acc = (double)t;
                    } else {
                        long long t = 0;
                        for (int w = 0; w < kTcWarps; ++w) t += s_changed[w];
                        acc = (double)t;
                    }
                    out[e] = acc;
                }
            }
        }
        tile = p_end;
        ++p;
    }
#ifdef DPR_TC_PROFILE
    if (blockIdx.x == 3 && lane == 0 && (warp & 3) == 0)
        printf("warp %d tiles %lld: wait_full %lld split %lld wait_mma %lld tmem_ld %lld argmin %lld recheck %lld store %lld sums+release %lld\n", warp, prof_n, prof_acc[1] / prof_n, prof_acc[2] / prof_n,
               prof_acc[3] / prof_n, prof_acc[4] / prof_n, prof_acc[5] / prof_n, prof_acc[6] / prof_n, prof_acc[7] / prof_n, prof_acc[8] / prof_n);
#endif
    tc_fence_before();
    consumer_barrier<kTcWarps * 32>();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"((uint32_t)T.tmem_cols));
}

// ---------------------------------------------------------------------------------------------------
// delta_sums_kernel: the sums/counts of a delta iteration.  CTA c walks the same (CTA, problem) ranges as CTA c of the pass
// that just ran; warp a turns the lists a, a + A, ... of that CTA's consumer warps into its private fp64 table in shared
// memory — for each recorded cell, +row to the new label and -row to the old one, lane = marker, in list order — then the
// tables are added up in warp order and written where the pass itself writes the tables of a full iteration.  Every step has
// a fixed order, so the sums are bit-reproducible; the rows are gathered from HBM (ceil(d/4) sectors per moved cell), eight
// cells in flight per warp.
// ---------------------------------------------------------------------------------------------------
constexpr int kDeltaBatch = 16;
__global__ void __launch_bounds__(512, 1) delta_sums_kernel(const PassParams P, const int d, const int W, const int A, const int table_doubles) {
    extern __shared__ __align__(16) unsigned char dsm[];
    double* tabs = reinterpret_cast<double*>(dsm);
    __shared__ uint32_t cursor[4 * kTcMaxGroups];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    DPR_GRID_DEPENDENCY();
    const int t_begin = P.cta_tile_start[blockIdx.x], t_end = P.cta_tile_start[blockIdx.x + 1];
    if (t_begin >= t_end) return;
    if (threadIdx.x < 4 * kTcMaxGroups) cursor[threadIdx.x] = 0;
    int p = 0;
    {
        int lo = 0, hi = P.n_problems - 1;
        while (lo < hi) {
            const int mid = (lo + hi + 1) >> 1;
            if (P.problems[mid].first_tile <= t_begin) lo = mid; else hi = mid - 1;
        }
        p = lo;
    }
    const int64_t stride = P.part_stride;
    const int64_t entries = stride + P.k_max + 1;
    const size_t tile_fl = tile_floats(d);
    int tile = t_begin;
    while (tile < t_end) {
        const Problem* pr = P.problems + p;
        const int p_end = min(t_end, pr->first_tile + pr->n_tiles);
        if (pr->done || p_end <= tile) {
            tile = max(tile, p_end);
            ++p;
            continue;
        }
        if (pr->delta) {   // (uniform over the CTA)
            for (int i = threadIdx.x; i < A * table_doubles; i += blockDim.x) tabs[i] = 0.0;
            __syncthreads();
            if (warp < A) {
                double* tab = tabs + (size_t)warp * table_doubles;
                long long* cnt = reinterpret_cast<long long*>(tab + stride);
                const float* __restrict__ x = pr->x;
                const float* __restrict__ xr = pr->xrows;   // cell-major copy (large matrices): a row is contiguous
                const int d4 = 4 * tile_groups(d);
                for (int l = warp; l < W; l += A) {
                    const int n_e = P.move_counts[(size_t)(blockIdx.x + p) * W + l];
                    const uint32_t pos_l = cursor[l];   // position of this range's first entry in list l
                    const uint2* __restrict__ list = P.move_lists + ((size_t)blockIdx.x * W + l) * P.move_cap + pos_l;
                    uint2 ent_next = make_uint2(0u, 0u);
                    if (lane < min(kDeltaBatch, n_e)) ent_next = list[lane];
                    for (int e0 = 0; e0 < n_e; e0 += kDeltaBatch) {
                        const int m = min(kDeltaBatch, n_e - e0);
                        const uint2 ent = ent_next;
                        if (e0 + kDeltaBatch + lane < n_e && lane < kDeltaBatch) ent_next = list[e0 + kDeltaBatch + lane];   // next batch's entries: in flight with the gathers
                        for (int jb = 0; jb < d; jb += 64) {
                            const int j0 = jb + lane, j1 = j0 + 32;
                            const bool on0 = j0 < d, on1 = j1 < d;
                            const int c0 = on0 ? j0 : 0, c1 = on1 ? j1 : 0;   // clamped: the gathers are unconditional (entries past the batch read cell 0)
                            const size_t jo0 = xr ? (size_t)c0 : (size_t)(c0 >> 2) * kGroupFloats + (c0 & 3);
                            const size_t jo1 = xr ? (size_t)c1 : (size_t)(c1 >> 2) * kGroupFloats + (c1 & 3);
                            float v0[kDeltaBatch], v1[kDeltaBatch];
#pragma unroll
                            for (int q = 0; q < kDeltaBatch; ++q) {   // all gathers of the batch in flight
                                const uint32_t cell = __shfl_sync(0xffffffffu, ent.x, q);
                                const float* row = xr ? xr + (size_t)cell * d4 : x + (size_t)(cell >> 7) * tile_fl + (size_t)(cell & 127u) * 4;
                                v0[q] = __ldg(row + jo0);
                                v1[q] = __ldg(row + jo1);
                            }
#pragma unroll
                            for (int q = 0; q < kDeltaBatch; ++q) {
                                const uint32_t lab = __shfl_sync(0xffffffffu, ent.y, q);
                                if (q < m) {
                                    // the four read-modify-writes of an entry are independent (new != old label, two different markers)
                                    double* tn = tab + (size_t)(lab & 0xffffu) * d;
                                    double* to = tab + (size_t)(lab >> 16) * d;
                                    const double n0 = on0 ? tn[j0] : 0.0, o0 = on0 ? to[j0] : 0.0;
                                    const double n1 = on1 ? tn[j1] : 0.0, o1 = on1 ? to[j1] : 0.0;
                                    if (on0) {
                                        tn[j0] = n0 + (double)v0[q];
                                        to[j0] = o0 - (double)v0[q];
                                    }
                                    if (on1) {
                                        tn[j1] = n1 + (double)v1[q];
                                        to[j1] = o1 - (double)v1[q];
                                    }
                                    if (jb == 0 && lane == 0) {
                                        cnt[lab & 0xffffu] += 1;
                                        cnt[lab >> 16] -= 1;
                                    }
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) cursor[l] += (uint32_t)n_e;
                }
            }
            __syncthreads();
            double* out = P.cta_tables + (size_t)(blockIdx.x + p) * entries;
            for (int64_t e = threadIdx.x; e < stride + P.k_max; e += blockDim.x) {
                if (e < stride) {
                    double acc = 0.0;
                    for (int a = 0; a < A; ++a) acc += tabs[(size_t)a * table_doubles + e];
                    out[e] = acc;
                } else {
                    long long t = 0;
                    for (int a = 0; a < A; ++a) t += reinterpret_cast<const long long*>(tabs + (size_t)a * table_doubles)[e];
                    out[e] = (double)t;
                }
            }
            __syncthreads();
        }
        tile = p_end;
        ++p;
    }
}


// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int plan_tc_pass(int d, int k_max, int sets_wanted, TcPlan* plan) {
    if (d < 1 || k_max < 1) return 0;
    const int npad = (k_max + 31) & ~31;
    const int g16 = 2 * ((d + 15) >> 4), acols = 4 * g16;
    if (2 * npad > 256) return 0;   // the widest product has npad + N <= 2 * npad columns
    int G = kTcMaxGroups;   // as many consumer groups as tensor memory holds (each: accumulator + xh + xl)
    while (G > 2 && G * (2 * npad + 2 * acols) > 512) --G;
    if (const char* e = getenv("DPR_TC_GROUPS")) {   // tuning aid
        const int want = atoi(e);
        if (want >= 2 && want < G) G = want;
    }
    int cols = G * (2 * npad + 2 * acols), tmem_cols = 32;
    while (tmem_cols < cols) tmem_cols <<= 1;
    if (tmem_cols > 512) return 0;
    const size_t stage_bytes = (((size_t)(tile_groups(d) + (tile_groups(d) & 1)) * kGroupFloats * 4) + 127) & ~(size_t)127;   // the tile (+ one zero group when its group count is odd)
    const size_t bars = 384;   // mbarriers, the tensor-memory address, the bundle's label / centre pointers
    const size_t hdr_bytes = (sizeof(CentreHeader) + 127) & ~(size_t)127;
    const size_t cc_bytes = (size_t)npad * 4, orig_bytes = (size_t)npad * 4;
    const size_t scratch = (((size_t)(2 * k_max + 2) * 4 + 64) + 127) & ~(size_t)127;
    const size_t b_bytes = (size_t)g16 * 2 * npad * 16;
    const size_t fixed0 = bars + scratch * 4 * G;
    size_t set_bytes = hdr_bytes + cc_bytes + orig_bytes + b_bytes;   // one centre set: header, ||c||^2, column -> row map, B operand
    if (fixed0 + set_bytes + 3 * stage_bytes > (size_t)kSmemBudget) return 0;
    // the exact centres next to the tables when they are small
    const size_t mu_bytes = (((size_t)k_max * (d | 1) * 8) + 127) & ~(size_t)127;
    const bool mu_in_smem = mu_bytes <= 32 * 1024 && fixed0 + set_bytes + mu_bytes + 4 * stage_bytes <= (size_t)kSmemBudget;
    if (mu_in_smem) set_bytes += mu_bytes;
    // centre sets resident at once (bundles of an assignment-only session): as many as leave 4 stages, at most kTcMaxSets
    int sets = std::max(1, std::min(sets_wanted, kTcMaxSets));
    if (const char* e = getenv("DPR_TC_SETS")) sets = std::max(1, std::min(sets, atoi(e)));   // tuning aid
    while (sets > 1 && fixed0 + (size_t)sets * set_bytes + 4 * stage_bytes > (size_t)kSmemBudget) --sets;
    size_t fixed = fixed0 + (size_t)sets * set_bytes;
    // Lloyd tables of the consumer groups in shared memory, when at least 4 stages remain (Lloyd sessions hold one set)
    const size_t gtab_bytes = (((size_t)k_max * d + k_max) * 8 + 127) & ~(size_t)127;
    const size_t glist_bytes = 256;   // [2][128] labels
    const bool gtab = sets_wanted <= 1 && k_max <= 255 && fixed + (gtab_bytes + glist_bytes) * G + 4 * stage_bytes <= (size_t)kSmemBudget;
    const size_t off_gtab = fixed;
    if (gtab) fixed += (gtab_bytes + glist_bytes) * G;
    int S = (int)(((size_t)kSmemBudget - fixed) / stage_bytes);
    if (S > kTcMaxStages) S = kTcMaxStages;
    plan->d = d;
    plan->groups = G;
    plan->g8 = g16;
    plan->wait_hint = 2000;
    plan->key_mul = 16u;
    plan->npad = npad;
    plan->tmem_cols = tmem_cols;
    plan->stages = S;
    plan->stage_floats = (int)(stage_bytes / 4);
    plan->off_scratch = (int)bars;
    plan->scratch_bytes = (int)scratch;
    plan->off_hdr = (int)fixed0;
    plan->off_cc = plan->off_hdr + (int)hdr_bytes;
    plan->off_orig = plan->off_cc + (int)cc_bytes;
    plan->off_b = plan->off_orig + (int)orig_bytes;
    plan->mu_in_smem = mu_in_smem ? 1 : 0;
    plan->off_mu = plan->off_b + (int)b_bytes;
    plan->sets = sets;
    plan->set_bytes = (int)set_bytes;
    plan->gtab = gtab ? 1 : 0;
    plan->gtab_bytes = (int)gtab_bytes;
    plan->glist_bytes = (int)glist_bytes;
    plan->off_gtab = (int)off_gtab;
    plan->off_glist = (int)(off_gtab + gtab_bytes * G);
    plan->off_stages = (int)fixed;
    plan->smem_bytes = (size_t)plan->off_stages + (size_t)S * stage_bytes;
    plan->table_floats = (size_t)npad + b_bytes / 4;
    return 1;
}

int launch_tc_pass(const TcPlan& plan, const PassParams& P, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
#define DPR_TC_ATTR(G)                                                                                                                   \
    DPR_CUDA(cudaFuncSetAttribute(tc_pass_kernel<kPassAssign, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget)); \
    DPR_CUDA(cudaFuncSetAttribute(tc_pass_kernel<kPassAssign, G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));  \
    DPR_CUDA(cudaFuncSetAttribute(tc_pass_kernel<kPassLloyd, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));      \
    DPR_CUDA(cudaFuncSetAttribute(tc_pass_kernel<kPassLloydDelta, G, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        DPR_TC_ATTR(2) DPR_TC_ATTR(3) DPR_TC_ATTR(4)
#undef DPR_TC_ATTR
        configured = true;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx().time_passes) {
        DPR_CUDA(cudaEventCreate(&e0));
        DPR_CUDA(cudaEventCreate(&e1));
        DPR_CUDA(cudaEventRecord(e0, stream));
    }
    const int threads = (4 * plan.groups + 2) * 32;
#define DPR_TC_LAUNCH(G)                                                                                                          \
    if (plan.groups == G) {                                                                                                       \
        if (P.mode == kPassAssign && P.bundles) tc_pass_kernel<kPassAssign, G, true><<<grid, threads, plan.smem_bytes, stream>>>(P, plan); \
        else if (P.mode == kPassAssign) tc_pass_kernel<kPassAssign, G, false><<<grid, threads, plan.smem_bytes, stream>>>(P, plan);       \
        else if (P.all_delta) DPR_CUDA(launch_dependent(tc_pass_kernel<kPassLloydDelta, G, false>, dim3(grid), dim3(threads), plan.smem_bytes, stream, P, plan)); \
        else DPR_CUDA(launch_dependent(tc_pass_kernel<kPassLloyd, G, false>, dim3(grid), dim3(threads), plan.smem_bytes, stream, P, plan));   \
    }
    DPR_TC_LAUNCH(2) DPR_TC_LAUNCH(3) DPR_TC_LAUNCH(4)
#undef DPR_TC_LAUNCH
    count_launch();
    if (e0) {
        DPR_CUDA(cudaEventRecord(e1, stream));
        ctx().pass_events.emplace_back(e0, e1);
    }
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

int launch_delta_sums(const TcPlan& plan, const PassParams& P, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        DPR_CUDA(cudaFuncSetAttribute(delta_sums_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget - 1024));
        configured = true;
    }
    const int W = 4 * plan.groups;
    const size_t table_bytes = ((size_t)P.part_stride + P.k_max) * 8;
    const size_t budget = (size_t)kSmemBudget - 1024;
    int A = (int)std::min<size_t>((size_t)W, budget / table_bytes);
    if (A < 1) return fail(DPR_ERR_INVALID, "delta sums: one k x d table does not fit in shared memory");
    DPR_CUDA(launch_dependent(delta_sums_kernel, dim3(grid), dim3(512), (size_t)A * table_bytes, stream, P, plan.d, W, A, (int)(table_bytes / 8)));
    count_launch();
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
