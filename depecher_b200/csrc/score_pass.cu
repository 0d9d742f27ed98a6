// score_pass.cu — the scoring pass of the penalty search on the 5th-generation tensor cores: the SAME cells assigned to many
// centre sets (Clusterer::optimize_param scores every fitted pair on the full data, /root/reference/src/Clusterer.cpp:370-371:
// 2 sets per penalty and worker, 220 per set of depeche()'s default sweep), optionally with the K x K contingency counts of each
// (ret1, ret2) pair taken inside the pass so that no label ever goes to HBM (cluster_distance's inputs, :376 and :267-305).
//
// Same arithmetic as tc_pass.cu (fp16-split operands, fp32 accumulation, exact fp64 re-decision of every near tie: labels are
// bit-exact), different shape — the round-1 bundled instance of tc_pass_kernel ran at 29 % of its HBM roofline (0.237-0.245 ms per
// set at 10M x 40, K = 30), bound by a chain of latencies per (tile, set) inside a consumer group:
//   * the centre sets of a bundle (up to 4) are STACKED along N: one B operand of sum(n_q) rows, so one round of products per
//     tile serves every set (one wait per tile instead of one per set); a set only occupies round16(active centres) columns;
//   * the three partial products xh*ch, xh*cl, xl*ch accumulate into the SAME columns (3 tcgen05.mma per K = 16 step), so a
//     score is one TMEM column, not the sum of two halves: half the tcgen05.ld traffic, and the add disappears;
//   * roles: one group of four warps only SPLITS tiles (fp32 -> fp16 high/low parts into one of two A buffers in tensor memory),
//     three groups of four warps only REDUCE (tcgen05.ld -> two smallest scores -> label), one accumulator each, so the products of
//     tile t+1 and t+2 run while tile t is reduced; one warp issues the products, one keeps the ring of bulk copies in flight.
//     512 TMEM columns = 3 accumulators x 128 + 2 x (xh + xl).  The roles run under different register limits (setmaxnreg:
//     112 / 88 / 56 registers per thread for reduce / split / service, 480 per sub-partition), each with its own copy of the code;
//   * the reduction is built for the ALU pipe (min / max / logic at half the FMA pipe's rate) and for instruction-level
//     parallelism (three reducing warps per scheduler: dependent-issue latency, not a pipe, is what binds): scores become sortable
//     integer keys with the column number inside on the FMA pipe (one fma + one shift-add), the smallest key of a block of 16 is a
//     tree of 3-input mins, the second smallest comes from an unsigned-min trick (add + min fuse into VIADDMNMX), and two blocks
//     are reduced side by side without a branch between them.
// Measured (profiles/r02_summary.md): 0.154 ms per set (4 sets per tile) = 2.85 TB/s of the 176 B per cell and 4 sets = 44 % of the
// HBM roofline; tensor-memory reads are NOT the limit (profiles/microbench/ldtm_probe.cu: 215-370 B/cycle/SM, the pass needs ~25).
#include "common.cuh"
#include "fused_pass.cuh"
#include "pass_device.cuh"
#include "tc_device.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

namespace dpr {

constexpr int kScAccs = 3;          // accumulators = reducing groups
constexpr int kScABufs = 2;         // A-operand buffers (tiles split but not yet multiplied)
constexpr int kScAccCols = 128;     // TMEM columns per accumulator: the stacked sets of a bundle
constexpr int kScConsumerWarps = 4 * (kScAccs + 1);   // 12 reducing + 4 splitting
constexpr int kScThreads = (kScConsumerWarps + 4) * 32;   // + one warpgroup of service warps: products, bulk copies, two idle
// Registers per thread after the prologue (setmaxnreg works on warpgroups = 4 consecutive warps; every sub-partition of the SM holds one
// warp of each warpgroup, so 3 * reduce + split + service <= 512): the reducing warps hold two 16-column blocks of scores
constexpr int kScRegsReduce = 112, kScRegsSplit = 88, kScRegsService = 56;
static_assert(kScAccs * kScRegsReduce + kScRegsSplit + kScRegsService <= 480, "the CTA owns 5 warps x 96 registers per SM sub-partition");

struct ScoreSet {          // one centre set of the current bundle (shared memory); the first 32 bytes are what the end of a set's reduction reads
    int32_t col0, n_act;   // its first accumulator column; active centres
    float tk_a, tk_b;      // near-tie bound in key units: tk_a * X + tk_b * sqrt(X) + tk_c, X = the tile's bound of ||x||^2 (+inf: always exact)
    float tk_c;
    int32_t ncols, trivial, pad_;   // columns (a multiple of 16); fewer than two active rows: every label 0
    int32_t* labels;       // nullable: scoring with fused contingency counts stores no labels
    const double* mu;      // exact centres (global memory; only the rare fp64 re-decisions read them)
};

static_assert(sizeof(ScoreSet) == 48, "the reducer reads a set's first 16 bytes as one float4");

// A set of at most two 16-column blocks as the reducer's short path reads it (two 16-byte loads).  A set with one block gets a second
// one made of padding: the same accumulator columns with factor 0 and the ||c||^2 of padding columns, so both blocks of every set go
// through the same straight-line code.
struct FastSet {
    uint32_t cols;         // column of block A | column of block B << 8 | set number << 16 | first ||c||^2 entry of block B << 24
    float unscale_a, unscale_b;
    uint32_t col0_nact;    // the set's first column | active centres << 8
    float tk_a, tk_b, tk_c;   // near-tie bound in key units (ScoreSet)
    uint32_t pad_;
};
static_assert(sizeof(FastSet) == 32, "two float4 loads");

// the two smallest of a sorted pair (a0 <= a1) and a sorted pair (b0 <= b1), sorted
__device__ __forceinline__ void merge2(float a0, float a1, float b0, float b1, float& lo, float& hi) {
    lo = fminf(a0, b0);
    hi = fmin3(fmaxf(a0, b0), a1, b1);
}

#ifdef DPR_SC_PROFILE
#define SC_PROF_DECL long long pt_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pa_[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pn_ = 0
#define SC_PROF(i) do { pt_[i] = clock64(); if (i > 0) pa_[i] += pt_[i] - pt_[i - 1]; } while (0)
#define SC_PROF_TILE ++pn_
#else
#define SC_PROF_DECL
#define SC_PROF(i)
#define SC_PROF_TILE
#endif

template <bool kPairs>
__global__ void __launch_bounds__(kScThreads, 1) score_pass_kernel(const PassParams P, const ScorePlan T) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int t_begin = P.cta_tile_start[blockIdx.x], t_end = P.cta_tile_start[blockIdx.x + 1];
    if (t_begin >= t_end) return;

    const int S = T.stages;
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);   // [S] tile landed
    uint64_t* empty = full + kScMaxStages;                // [S] tile reduced (one arrival per warp of the reducing group)
    uint64_t* a_ready = empty + kScMaxStages;             // [2] A buffer written (one arrival per splitting warp)
    uint64_t* a_free = a_ready + kScABufs;                // [2] the products that read the A buffer are complete
    uint64_t* acc_full = a_free + kScABufs;               // [3] products into the accumulator complete
    uint64_t* acc_free = acc_full + kScAccs;              // [3] the accumulator has been read (one arrival per reducing warp)
    uint64_t* tables_ready = acc_free + kScAccs;          // the bundle's centre tables are in shared memory
    uint32_t* tmem_word = reinterpret_cast<uint32_t*>(tables_ready + 1);
    int32_t* n_total_s = reinterpret_cast<int32_t*>(tmem_word + 1);   // accumulator columns the current bundle uses
    int32_t* n_chunks_s = n_total_s + 1;                               // blocks of 16 columns that hold scores (trivial sets have none)
    uint2* chunk_s = reinterpret_cast<uint2*>(n_chunks_s + 2);   // (8-byte aligned)   // [<= 8] .x: column | set << 8 | first of its set << 12 | last << 13; .y: its set's un-scaling factor
    int32_t* n_fast_s = n_chunks_s + 1;                          // >= 0: every non-trivial set of the bundle has at most two blocks — that many FastSets; -1: the general walk
    FastSet* fast_s = reinterpret_cast<FastSet*>(smem + 384);    // [kScMaxSets] (the chunk list ends at 328)
    ScoreSet* sets = reinterpret_cast<ScoreSet*>(smem + T.off_sets);
    float* cc_s = reinterpret_cast<float*>(smem + T.off_cc);          // [b_rows] ||c||^2 per accumulator column (padding: +huge)
    int32_t* orig_s = reinterpret_cast<int32_t*>(smem + T.off_orig);  // [b_rows] column -> row of its set's mu (padding: -1)
    float4* bh = reinterpret_cast<float4*>(smem + T.off_bh);          // stacked high parts, K-major canonical layout
    float4* bl = reinterpret_cast<float4*>(smem + T.off_bl);          // stacked low parts
    unsigned int* pair_s = reinterpret_cast<unsigned int*>(smem + T.off_pairs);   // [sets / 2][kk * kk] contingency counts of this CTA
    float* stages = reinterpret_cast<float*>(smem + T.off_stages);
    const int d = T.d, groups = tile_groups(d), k16 = T.k16;
    const size_t tile_fl = tile_floats(d);
    const uint32_t tile_bytes = (uint32_t)(tile_fl * 4);
    const int acols = T.acols;                                        // TMEM columns of one fp16 A operand (2 markers each)
    const uint32_t a_base = (uint32_t)(kScAccs * kScAccCols);
    const int kk = T.kk, kk2 = kk * kk;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 4);
        }
        for (int b = 0; b < kScABufs; ++b) {
            mbar_init(&a_ready[b], 4);
            mbar_init(&a_free[b], 1);
        }
        for (int c = 0; c < kScAccs; ++c) {
            mbar_init(&acc_full[c], 1);
            mbar_init(&acc_free[c], 4);
        }
        mbar_init(tables_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_word)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (kPairs)
        for (int i = threadIdx.x; i < (T.sets / 2) * kk2; i += blockDim.x) pair_s[i] = 0u;
    if (groups & 1)   // the split reads marker groups in pairs: the partner of an odd last group is never written by a copy — zero, once
        for (int s = 0; s < S; ++s)
            for (int i = threadIdx.x; i < kGroupFloats; i += blockDim.x) stages[(size_t)s * T.stage_floats + (size_t)groups * kGroupFloats + i] = 0.f;
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
    int tile = t_begin;

    // ---- hand the registers to the warps that need them
    if (warp >= kScConsumerWarps) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kScRegsService));
        if (warp >= kScConsumerWarps + 2) return;   // (the warpgroup's two spare warps)
    }

    // ================================ producer warp: the ring of bulk copies ================================
    if (warp == kScConsumerWarps + 1) {
        int s = 0, issued = 0;
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
    if (warp == kScConsumerWarps) {
        uint32_t seq = 0, n_range = 0, c = 0;     // tiles walked; ranges walked; accumulator of tile seq (= seq % 3)
        uint32_t u3 = 0;                          // use number of accumulator c (= seq / 3)
        const uint32_t lbo = (uint32_t)T.b_rows * 16u;   // bytes between the 8-marker groups of a stacked B operand
        const uint64_t desc_h0 = umma_desc(smem_u32(bh), lbo, 128), desc_l0 = umma_desc(smem_u32(bl), lbo, 128);
        const uint64_t desc_step = (uint64_t)((2u * lbo) >> 4);   // one K = 16 step = two marker groups further
        SC_PROF_DECL;
        while (tile < t_end) {
            const Problem* pr = P.problems + p;
            const int p_end = min(t_end, pr->first_tile + pr->n_tiles);
            if (pr->done || p_end <= tile) {
                tile = max(tile, p_end);
                ++p;
                continue;
            }
            mbar_wait_hint(tables_ready, n_range & 1u, T.wait_hint);
            ++n_range;
            const uint32_t n_total = (uint32_t)*reinterpret_cast<volatile int32_t*>(n_total_s);
            const uint32_t idesc = (1u << 4) | ((uint32_t)(128 >> 4) << 24) | ((n_total >> 3) << 17);   // D fp32, A/B fp16, both K-major, M = 128, N = n_total
            for (; tile < p_end; ++tile, ++seq) {
                const uint32_t b = seq & 1u, ub = seq >> 1;
                SC_PROF(0);
                mbar_wait_hint(&a_ready[b], ub & 1u, T.wait_hint);
                SC_PROF(1);
                if (u3 > 0) mbar_wait_hint(&acc_free[c], (u3 - 1u) & 1u, T.wait_hint);
                SC_PROF(2);
                tc_fence_after();
                if (elect_one()) {
                    const uint32_t acc = tmem + c * (uint32_t)kScAccCols;
                    const uint32_t ah = tmem + a_base + b * (uint32_t)(2 * acols), al = ah + (uint32_t)acols;
                    for (int s = 0; s < k16; ++s) {
                        umma_f16_ts(acc, ah + 8 * s, desc_h0 + s * desc_step, idesc, s > 0 ? 1u : 0u);   // xh * ch
                        umma_f16_ts(acc, ah + 8 * s, desc_l0 + s * desc_step, idesc, 1);                  // xh * cl
                        umma_f16_ts(acc, al + 8 * s, desc_h0 + s * desc_step, idesc, 1);                  // xl * ch
                    }
                    umma_commit(&acc_full[c]);
                    umma_commit(&a_free[b]);
                }
                __syncwarp();
                SC_PROF(3);
                SC_PROF_TILE;
                if (++c == kScAccs) {
                    c = 0;
                    ++u3;
                }
            }
            ++p;
        }
#ifdef DPR_SC_PROFILE
        if (blockIdx.x == 3 && lane == 0 && pn_) printf("issuer  tiles %lld: wait_a_ready %lld wait_acc_free %lld issue %lld\n", pn_, pa_[1] / pn_, pa_[2] / pn_, pa_[3] / pn_);
#endif
        return;
    }

    // ================================ consumer warps: 3 reducing groups + 1 splitting group ================================
    // The two roles run under different register limits (setmaxnreg): each gets its own copy of the walk over the ranges, so that the
    // compiler never has to fit the reducing code under the splitting warps' limit.
    const int grp = warp >> 2, wq = warp & 3;   // group; TMEM lane quarter this warp may touch
    auto consumer = [&](auto role) {
    constexpr bool splitter = decltype(role)::value;
    const int mycell_in_tile = wq * 32 + lane;
    const uint32_t lane_base = tmem + ((uint32_t)(wq * 32) << 16);
    const int steps8 = (d + 7) >> 3;   // 8-marker steps that hold data
    if constexpr (splitter) {   // zero padding of the last K = 16 step, in both parts of both A buffers: written once
        const uint32_t z[4] = {0u, 0u, 0u, 0u};
        for (int b = 0; b < kScABufs; ++b)
            for (int c = 4 * steps8; c < acols; c += 4) {
                tmem_st4(lane_base + a_base + (uint32_t)(b * 2 * acols + c), z);
                tmem_st4(lane_base + a_base + (uint32_t)(b * 2 * acols + acols + c), z);
            }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    }
    // error bound of a score: see tc_pass.cu.  Here 3 * k16 products accumulate into one column (2^-21 allowed per step, measured
    // <= 2^-24), and there is no sum of two accumulator halves.
    const float tau_a = 2.f * (3.f * 2.3841858e-7f + (float)(3 * k16 + 1) * 4.7683716e-7f);
    const float tau_b = 2.f * (4.f * 5.9604645e-8f);   // (the index packing has its own, absolute term: ScoreSet::tau_q)

    SC_PROF_DECL;
    uint32_t seq = 0;               // tiles this CTA has walked past
    int s_idx = 0;                  // stage of tile seq (= seq % S)
    uint32_t s_use = 0;             // seq / S
    uint32_t r3 = 0;                // seq % 3
    int prev_pair0 = -1, prev_pairs = 0;   // the pair tables in shared memory belong to these global pairs (kPairs)
    const int tid = threadIdx.x;    // < 512

    while (tile < t_end) {
        const Problem pr = P.problems[p];
        const int p_end = min(t_end, pr.first_tile + pr.n_tiles);
        if (pr.done || p_end <= tile) {
            tile = max(tile, p_end);
            ++p;
            continue;
        }
        const int NS = max(1, pr.bundle);   // centre sets that share this range's cells (problems p .. p + NS - 1)
        consumer_barrier<kScConsumerWarps * 32>();   // every group is past the previous range: its tables and counts are final
        if (kPairs && prev_pairs > 0) {   // this CTA's counts of the previous bundle's pairs -> the global tables
            for (int e = tid; e < prev_pairs * kk2; e += kScConsumerWarps * 32) {
                const unsigned int v = pair_s[e];
                if (v) {
                    atomicAdd(P.pair_tables + (size_t)(prev_pair0 + e / kk2) * kk2 + (e % kk2), (unsigned long long)v);
                    pair_s[e] = 0u;
                }
            }
        }
        prev_pair0 = p >> 1;
        prev_pairs = NS >> 1;
        {   // the bundle's tables: stacked B operands, ||c||^2 and the column -> row maps, per-set constants
            int col0[kScMaxSets + 1];
            col0[0] = 0;
            float ccm = 0.f;   // largest ||c||^2 of the bundle
#pragma unroll
            for (int q = 0; q < kScMaxSets; ++q) {
                col0[q + 1] = col0[q] + (q < NS ? P.problems[p + q].hdr->tc_n : 0);
                if (q < NS) ccm = fmaxf(ccm, P.problems[p + q].hdr->cc_max);
            }
            // sortable integer keys with the column number inside (tc_device.cuh: key_quantum, top2_block16); one scale for the bundle
            const float qk = key_quantum(*pr.xx_max + 2.0f * ccm);   // (tc_device.cuh)
            const bool q_ok = qk > 0.f;
            const float qq = q_ok ? qk : 1.f;
            const float big_m = q_ok ? 12582912.f * qq : 0.f;     // 1.5 * 2^23 * q
#pragma unroll
            for (int q = 0; q < kScMaxSets; ++q) {
                if (q >= NS) break;
                const Problem* pq = P.problems + p + q;
                const int nq = col0[q + 1] - col0[q], npad = pq->tc_npad;
                const float* tab = pq->tc_table;
                const float4* bsrc = reinterpret_cast<const float4*>(tab + npad);   // per 8-marker group: npad high rows, npad low rows (16 B each)
                for (int i = tid; i < 2 * k16 * nq; i += kScConsumerWarps * 32) {
                    const int g = i / nq, r = i - g * nq;
                    bh[(size_t)g * T.b_rows + col0[q] + r] = bsrc[(size_t)g * 2 * npad + r];
                    bl[(size_t)g * T.b_rows + col0[q] + r] = bsrc[(size_t)g * 2 * npad + npad + r];
                }
                for (int i = tid; i < nq; i += kScConsumerWarps * 32) {
                    // (padding columns: above every real score and still inside the binade of M)
                    cc_s[col0[q] + i] = i < pq->hdr->k_act ? tab[i] + big_m : big_m + kKeyPad * qq;
                    orig_s[col0[q] + i] = pq->tc_orig[i];
                }
                if (tid == 32 * q) {
                    const CentreHeader* h = pq->hdr;
                    const float cc2 = 2.0f * h->cc_max, iq16 = 16.f / qq;
                    ScoreSet st;
                    st.col0 = col0[q];
                    st.n_act = h->k_act;
                    st.ncols = nq;
                    st.trivial = h->trivial;
                    st.pad_ = 0;
                    // tau = tau_a (X + cc2 / 2) + tau_b (X + cc2) + abs1 sqrt(X) + abs2 + 3 q   (tc_pass.cu's bound without its index-bit
                    // term, plus the 3 q of the keys), divided by the key unit q / 16
                    st.tk_a = (tau_a + tau_b) * iq16;
                    st.tk_b = h->tc_abs1 * iq16;
                    st.tk_c = (h->force_exact || !q_ok) ? INFINITY : (tau_a * 0.5f * cc2 + tau_b * cc2 + h->tc_abs2 + 3.f * qq) * iq16;
                    st.labels = pq->labels;
                    st.mu = pq->mu;
                    sets[q] = st;
                }
            }
            if (tid < 16) cc_s[T.b_rows + tid] = big_m + kKeyPad * qq;   // a block of padding columns (second block of a one-block set)
            if (tid == 0) {
                *n_total_s = col0[NS];
                int nch = 0, nf = 0;
                bool fast_ok = true;
                for (int q = 0; q < NS; ++q) {
                    const CentreHeader* h = P.problems[p + q].hdr;
                    if (h->trivial) continue;
                    for (int c = col0[q]; c < col0[q + 1]; c += 16)
                        chunk_s[nch++] = make_uint2((uint32_t)c | ((uint32_t)q << 8) | (c == col0[q] ? 1u << 12 : 0u) | (c + 16 == col0[q + 1] ? 1u << 13 : 0u),
                                                    __float_as_uint(h->tc_unscale));
                    const int nq = col0[q + 1] - col0[q];
                    if (nq > 32) {
                        fast_ok = false;
                        continue;
                    }
                    const bool two = nq == 32;
                    const float cc2 = 2.0f * h->cc_max, iq16 = 16.f / qq;
                    FastSet f;
                    f.cols = (uint32_t)col0[q] | ((uint32_t)(two ? col0[q] + 16 : col0[q]) << 8) | ((uint32_t)q << 16) | ((uint32_t)(two ? col0[q] + 16 : T.b_rows) << 24);
                    f.unscale_a = h->tc_unscale;
                    f.unscale_b = two ? h->tc_unscale : 0.f;
                    f.col0_nact = (uint32_t)col0[q] | ((uint32_t)h->k_act << 8);
                    f.tk_a = (tau_a + tau_b) * iq16;      // (as in ScoreSet below)
                    f.tk_b = h->tc_abs1 * iq16;
                    f.tk_c = (h->force_exact || !q_ok) ? INFINITY : (tau_a * 0.5f * cc2 + tau_b * cc2 + h->tc_abs2 + 3.f * qq) * iq16;
                    f.pad_ = 0u;
                    fast_s[nf++] = f;
                }
                *n_chunks_s = nch;
                *n_fast_s = fast_ok ? nf : -1;
            }
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the B operands were written through the generic proxy
        consumer_barrier<kScConsumerWarps * 32>();
        if (tid == 0) mbar_arrive(tables_ready);
        const float scale_x = pr.hdr->tc_scale_x;   // (the same for every set of a bundle: it depends on the cells only)

        if constexpr (splitter) {
            // ---- fp32 cells -> fp16 high / low parts, parked in tensor memory as the A operands of tile seq
            for (; tile < p_end; ++tile, ++seq) {
                const uint32_t b = seq & 1u, ub = seq >> 1;
                SC_PROF(0);
                mbar_wait_hint(&full[s_idx], s_use & 1u, T.wait_hint);
                SC_PROF(1);
                if (ub > 0) mbar_wait_hint(&a_free[b], (ub - 1u) & 1u, T.wait_hint);
                SC_PROF(2);
                tc_fence_after();
                const float* slot = stages + (size_t)s_idx * T.stage_floats;
                const float4* st4 = reinterpret_cast<const float4*>(slot) + mycell_in_tile;
                const uint32_t a_hi = lane_base + a_base + b * (uint32_t)(2 * acols), a_lo = a_hi + (uint32_t)acols;
                if (scale_x == 1.f) split_tile<false>(st4, steps8, 1.f, a_hi, a_lo);
                else split_tile<true>(st4, steps8, scale_x, a_hi, a_lo);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&a_ready[b]);
                SC_PROF(3);
                SC_PROF_TILE;
                if (++s_idx == S) {
                    s_idx = 0;
                    ++s_use;
                }
            }
        } else {
            // ---- scores of tile seq (accumulator grp) -> labels of every set (+ contingency counts of the pairs)
            // this group's tiles of the range: every third one (tile seq goes to accumulator seq % 3).  The tile's norm bound comes from
            // global memory: fetched one own tile ahead, so that its latency is never waited for
            const int count = p_end - tile, t0 = tile - pr.first_tile;
            const int j0 = (int)((uint32_t)grp + kScAccs - r3) % kScAccs;
            int s_mine = s_idx + j0;
            while (s_mine >= S) s_mine -= S;
            uint32_t u_mine = (seq + (uint32_t)j0) / kScAccs;
            float xx_next = j0 < count ? pr.tile_xx[t0 + j0] : 0.f;
            for (int j = j0; j < count; j += kScAccs, ++u_mine) {
                {
                    const float* slot = stages + (size_t)s_mine * T.stage_floats;
                    const int s_cur = s_mine;
                    s_mine += kScAccs;
                    while (s_mine >= S) s_mine -= S;
                    const int64_t mycell = (int64_t)(t0 + j) * kSlotCells + mycell_in_tile;
                    const bool live = mycell < pr.n;
                    const float xx_tile = xx_next;
                    if (j + kScAccs < count) xx_next = pr.tile_xx[t0 + j + kScAccs];
                    const float sq_xx = sqrtf(xx_tile);
                    SC_PROF(0);
                    mbar_wait_hint(&acc_full[grp], u_mine & 1u, T.wait_hint);
                    tc_fence_after();
                    SC_PROF(1);
                    // The bundle's scores are walked in blocks of 16 accumulator columns (chunk list built with the tables; the sets of a
                    // bundle follow one another, trivial sets have no block).  The tcgen05.ld of block c + 1 is in flight while block c is
                    // reduced: two register buffers, the loop unrolled by two.
                    uint32_t labs = 0;    // labels of the bundle's sets, 8 bits each (k <= 64 in this kernel)
                    unsigned doubt = 0;   // bit q: this cell's decision for set q needs the exact path
                    int best = 0x7fffffff, second = 0x7fffffff, bcol = 0;
                    const int nch = *n_chunks_s;
                    const uint32_t acc_base = lane_base + (uint32_t)(grp * kScAccCols);
                    auto release_acc = [&]() {   // every score of the tile is in registers: the accumulator may take the products of tile seq + 3
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&acc_free[grp]);
                    };
                    // one block of 16 keys -> (smallest, second smallest).  The smallest key b: a tree of 3-input mins (8 instructions).  The
                    // second smallest: the keys are distinct (column number inside), so key - b - 1 as an UNSIGNED number is huge for the
                    // smallest key and (key - b - 1) for every other one: second = b + 1 + unsigned-min of those (add + min fuse into one
                    // instruction, VIADDMNMX).  No branch inside: two blocks are reduced side by side (independent instruction streams for
                    // the scheduler — with three reducing warps per sub-partition the pass is bound by dependent-issue latency, not by a pipe).
                    auto top2 = [&](const uint32_t (&v)[16], const uint32_t cc_col, const float unscale, int& b, int& sec) {
                        top2_block16(v, reinterpret_cast<const float4*>(cc_s + cc_col), unscale, T.key_mul, b, sec);
                    };
                    // a block's result into the running (best, second) of its set; at the set's last block: the label, and whether the two
                    // best scores are safely apart
                    auto fold = [&](const uint2 ci, int b, int sec) {
                        const uint32_t info = ci.x;   // column | set << 8 | first block of its set << 12 | last << 13
                        const bool first = (info & (1u << 12)) != 0;
                        const int col = (int)(info & 255u);
                        if (first || b < best) bcol = col;   // (strict: an equal key of a later block never displaces an earlier one)
                        second = first ? sec : __vimin3_s32(max(best, b), second, sec);
                        best = first ? b : min(best, b);
                        if (info & (1u << 13)) {
                            const int q = (int)((info >> 8) & 15u);
                            const float4 s0 = *reinterpret_cast<const float4*>(sets + q);   // col0, n_act, tk_a, tk_b
                            const float tk = fmaf(s0.z, xx_tile, fmaf(s0.w, sq_xx, sets[q].tk_c));
                            const int a = bcol - __float_as_int(s0.x) + (best & 15);
                            labs |= (uint32_t)orig_s[__float_as_int(s0.x) + (a < __float_as_int(s0.y) ? a : 0)] << (8 * q);
                            if (!((float)(second - best) > tk)) doubt |= 1u << q;
                        }
                    };
                    // blocks two at a time: both loads in flight, one wait, both tournaments in one straight-line stretch (three or four at
                    // a time need more than the 112 registers of a reducing warp: measured 0.158-0.162 ms per set with their spills, 0.154 with two)
                    uint32_t va[16], vb[16];
                    const int nf = *n_fast_s;
                    if (nf >= 0) {
                        // the short path (every set of the bundle has at most 32 columns: the penalty search's K = 30): one set per round, its
                        // two blocks reduced side by side and merged directly — no per-block bookkeeping
                        for (int f = 0; f < nf; ++f) {
                            const uint4 f0 = *reinterpret_cast<const uint4*>(fast_s + f);
                            const float4 f1 = *(reinterpret_cast<const float4*>(fast_s + f) + 1);
                            const uint32_t col_a = f0.x & 255u, col_b = (f0.x >> 8) & 255u;
                            tmem_ld16_raw(acc_base + col_a, va);
                            tmem_ld16_raw(acc_base + col_b, vb);
                            tmem_wait_ld16(va);
                            tmem_wait_ld16(vb);
                            if (f + 1 == nf) release_acc();
                            int ba, sa, bb, sb;
                            top2(va, col_a, __uint_as_float(f0.y), ba, sa);
                            top2(vb, f0.x >> 24, __uint_as_float(f0.z), bb, sb);
                            const int bst = min(ba, bb);                       // (an equal key of block B never displaces block A's: strict below)
                            const int sec = __vimin3_s32(max(ba, bb), sa, sb);
                            const int col0_q = (int)(f0.w & 255u), n_act_q = (int)(f0.w >> 8);
                            const int a = (int)(bb < ba ? col_b : col_a) - col0_q + (bst & 15);
                            const uint32_t sh = (f0.x >> 13) & 0x18u;          // 8 * set number
                            labs |= (uint32_t)orig_s[col0_q + (a < n_act_q ? a : 0)] << sh;
                            const float tk = fmaf(f1.x, xx_tile, fmaf(f1.y, sq_xx, f1.z));
                            if (!((float)(sec - bst) > tk)) doubt |= 1u << (sh >> 3);
                        }
                        if (nf == 0) release_acc();
                    } else {
                        for (int c = 0; c < nch; c += 2) {
                            const uint2 ca = chunk_s[c], cb = chunk_s[c + 1 < nch ? c + 1 : c];
                            tmem_ld16_raw(acc_base + (ca.x & 255u), va);
                            tmem_ld16_raw(acc_base + (cb.x & 255u), vb);
                            tmem_wait_ld16(va);
                            tmem_wait_ld16(vb);
                            if (c + 2 >= nch) release_acc();
                            int ba, sa, bb, sb;
                            top2(va, ca.x & 255u, __uint_as_float(ca.y), ba, sa);
                            top2(vb, cb.x & 255u, __uint_as_float(cb.y), bb, sb);
                            fold(ca, ba, sa);
                            if (c + 1 < nch) fold(cb, bb, sb);
                        }
                        if (nch == 0) release_acc();
                    }
                    SC_PROF(2);
                    // near ties: the whole warp re-decides the cell in fp64 with the reference's arithmetic (exact_label_warp)
                    if (__any_sync(0xffffffffu, doubt != 0 && live)) {
#pragma unroll
                        for (int q = 0; q < kScMaxSets; ++q) {
                            if (q >= NS) break;
                            unsigned redo = __ballot_sync(0xffffffffu, ((doubt >> q) & 1u) && live);
                            if (!redo) continue;
                            if (P.recheck_count && lane == 0) atomicAdd(P.recheck_count, (unsigned long long)__popc(redo));
                            const ScoreSet* st = sets + q;
                            while (redo) {
                                const int src = __ffs(redo) - 1;
                                redo &= redo - 1;
                                const int v = exact_label_warp_inline(slot, wq * 32 + src, d, st->mu, d, orig_s + st->col0, st->n_act, lane);
                                if (lane == src) labs = (labs & ~(255u << (8 * q))) | ((uint32_t)v << (8 * q));
                            }
                        }
                    }
                    SC_PROF(3);
                    if (live) {
#pragma unroll
                        for (int q = 0; q < kScMaxSets; ++q) {
                            if (q >= NS) break;
                            int32_t* out = sets[q].labels;
                            if (out) out[mycell] = (int32_t)((labs >> (8 * q)) & 255u);
                        }
                        if (kPairs) {   // (ret1, ret2) of a penalty cell are neighbours in the bundle
                            if (NS >= 2) atomicAdd(&pair_s[(labs & 255u) * kk + ((labs >> 8) & 255u)], 1u);
                            if (NS >= 4) atomicAdd(&pair_s[kk2 + ((labs >> 16) & 255u) * kk + (labs >> 24)], 1u);
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&empty[s_cur]);
                    SC_PROF(4);
                    SC_PROF_TILE;
                }
            }
            seq += (uint32_t)count;
            r3 = seq % kScAccs;
            s_idx = (int)((uint32_t)(s_idx + count) % (uint32_t)S);
            tile = p_end;
        }
        tile = p_end;
        ++p;
    }
#ifdef DPR_SC_PROFILE
    if (blockIdx.x == 3 && lane == 0 && wq == 0 && pn_) {
        if (splitter) printf("splitter tiles %lld: wait_full %lld wait_a_free %lld split %lld\n", pn_, pa_[1] / pn_, pa_[2] / pn_, pa_[3] / pn_);
        else printf("reducer %d tiles %lld: wait_acc_full %lld ld+argmin(all sets) %lld release+recheck %lld store+pairs %lld\n", grp, pn_, pa_[1] / pn_, pa_[2] / pn_, pa_[3] / pn_, pa_[4] / pn_);
    }
#endif
    tc_fence_before();
    consumer_barrier<kScConsumerWarps * 32>();
    if (kPairs && prev_pairs > 0) {
        for (int e = tid; e < prev_pairs * kk2; e += kScConsumerWarps * 32) {
            const unsigned int v = pair_s[e];
            if (v) atomicAdd(P.pair_tables + (size_t)(prev_pair0 + e / kk2) * kk2 + (e % kk2), (unsigned long long)v);
        }
    }
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
    }
    };   // consumer
    if (grp == kScAccs) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kScRegsSplit));
        consumer(std::true_type{});
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kScRegsReduce));
        consumer(std::false_type{});
    }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
int plan_score_pass(int d, int k_max, int sets_wanted, int fuse_kk, ScorePlan* plan) {
    static const bool off = getenv("DPR_SCORE_OLD") != nullptr;   // A/B aid: keep the bundled instance of tc_pass_kernel
    if (off || d < 1 || k_max < 2 || sets_wanted < 1) return 0;
    const int k16 = (d + 15) >> 4;
    if (k16 > 4) return 0;   // beside the 3 accumulators tensor memory holds 2 A buffers of 2 x 8 * k16 columns
    const int ncol = (k_max + 15) & ~15;
    if (ncol > 64) return 0;
    const int sets = sets_wanted < 2 ? 1 : (kScAccCols / ncol >= 4 && sets_wanted >= 4) ? 4 : 2;   // (one set: a plain assignment — the pass is then bound by the splitting group, i.e. close to HBM)
    const int b_rows = sets * ncol;
    const int acols = 8 * k16;
    const size_t stage_bytes = (((size_t)(tile_groups(d) + (tile_groups(d) & 1)) * kGroupFloats * 4) + 127) & ~(size_t)127;
    const size_t bars = 640;   // mbarriers, counters, the chunk list and the FastSets
    const size_t sets_bytes = 512;   // kScMaxSets x ScoreSet (56 B)
    const size_t cc_bytes = ((size_t)(b_rows + 16) * 4 + 127) & ~(size_t)127;   // + one block of padding columns
    const size_t b_bytes = (size_t)2 * k16 * b_rows * 16;
    size_t pair_bytes = 0;
    int kk = 0;
    if (fuse_kk > 0) {
        pair_bytes = (((size_t)(sets / 2) * fuse_kk * fuse_kk * 4) + 127) & ~(size_t)127;
        if (pair_bytes <= 40 * 1024) kk = fuse_kk; else pair_bytes = 0;   // wide tables: the labels go to HBM and contingency_kernel counts them
    }
    const size_t fixed = bars + sets_bytes + 2 * cc_bytes + 2 * b_bytes + pair_bytes;
    if (fixed + 5 * stage_bytes > (size_t)kSmemBudget) return 0;
    int S = (int)(((size_t)kSmemBudget - fixed) / stage_bytes);
    if (S > kScMaxStages) S = kScMaxStages;
    plan->d = d;
    plan->k16 = k16;
    plan->ncol_max = ncol;
    plan->sets = sets;
    plan->b_rows = b_rows;
    plan->acols = acols;
    plan->wait_hint = 2000;
    plan->key_mul = 16u;
    plan->kk = kk;
    plan->stages = S;
    plan->stage_floats = (int)(stage_bytes / 4);
    plan->off_sets = (int)bars;
    plan->off_cc = plan->off_sets + (int)sets_bytes;
    plan->off_orig = plan->off_cc + (int)cc_bytes;
    plan->off_bh = plan->off_orig + (int)cc_bytes;
    plan->off_bl = plan->off_bh + (int)b_bytes;
    plan->off_pairs = plan->off_bl + (int)b_bytes;
    plan->off_stages = plan->off_pairs + (int)pair_bytes;
    plan->smem_bytes = (size_t)plan->off_stages + (size_t)S * stage_bytes;
    return 1;
}

int launch_score_pass(const ScorePlan& plan, const PassParams& P, int grid, cudaStream_t stream) {
    static bool configured = false;
    if (!configured) {
        DPR_CUDA(cudaFuncSetAttribute(score_pass_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        DPR_CUDA(cudaFuncSetAttribute(score_pass_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBudget));
        configured = true;
    }
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (ctx().time_passes) {
        DPR_CUDA(cudaEventCreate(&e0));
        DPR_CUDA(cudaEventCreate(&e1));
        DPR_CUDA(cudaEventRecord(e0, stream));
    }
    if (P.pair_tables && plan.kk > 0) score_pass_kernel<true><<<grid, kScThreads, plan.smem_bytes, stream>>>(P, plan);
    else score_pass_kernel<false><<<grid, kScThreads, plan.smem_bytes, stream>>>(P, plan);
    count_launch();
    if (e0) {
        DPR_CUDA(cudaEventRecord(e1, stream));
        ctx().pass_events.emplace_back(e0, e1);
    }
    DPR_CUDA(cudaGetLastError());
    return DPR_OK;
}

}  // namespace dpr
