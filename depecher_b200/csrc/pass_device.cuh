// pass_device.cuh — device helpers shared by the two pass kernels (fused_pass.cu: FFMA path, tc_pass.cu: tcgen05 path):
// mbarrier / bulk-copy PTX, the exact fp64 decision, and the Lloyd-table updates.
#pragma once
#include "common.cuh"

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
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* src, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ float fmin3(float a, float b, float c) {
    float r;
    asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
    return r;
}

// ---------------------------------------------------------------------------------------------------
// distance micro-kernel for one chunk of KC centres: 2 cells per lane
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ float pack_idx(float s, int idx) {
    return __uint_as_float((__float_as_uint(s) & 0xFFFFFFE0u) | (unsigned)idx);
}

// ---------------------------------------------------------------------------------------------------
// exact decision for one cell, by the whole warp: the reference's arithmetic (Clusterer.cpp:56-66) in fp64 —
// d_k = sqrt(sum_t ((-x_t) + mu_kt)^2), separate multiply and add (no contraction: the reference is built
// for baseline x86-64, which has no FMA), strict-< first minimum over the active rows in order.  Lane a
// takes active slots a, a+32, ...; the (distance, slot) pairs are then reduced lexicographically, which is
// the sequential first-minimum scan (a NaN distance is never chosen unless it is the first one, like there).
// ---------------------------------------------------------------------------------------------------
static __device__ __forceinline__ int exact_label_warp_inline(const float* __restrict__ slot, int cell, int d, const double* __restrict__ mu, int ld,
                                                              const int32_t* __restrict__ orig, int n_slots, int lane) {   // ld: doubles between rows of mu
    double bv = INFINITY;
    int bi = 0x7fffffff;
    bool first_nan = false;
    for (int a = lane; a < n_slots; a += 32) {
        const int row = orig[a];
        if (row < 0) continue;
        const double* m = mu + (size_t)row * ld;
        double s = 0.0;
        int t = 0;
        // four markers = one 16-byte group of the cell: the loads, differences and squares of a group are independent of one
        // another and of the running sum, only the four additions into s form the (left-to-right) chain
        const float4* x4 = reinterpret_cast<const float4*>(slot + cell * 4);
#pragma unroll 2
        for (; t + 4 <= d; t += 4) {
            const float4 xv = x4[(t >> 2) * (kGroupFloats / 4)];
            const double d0 = __dadd_rn(-(double)xv.x, m[t]), d1 = __dadd_rn(-(double)xv.y, m[t + 1]);
            const double d2 = __dadd_rn(-(double)xv.z, m[t + 2]), d3 = __dadd_rn(-(double)xv.w, m[t + 3]);
            const double q0 = __dmul_rn(d0, d0), q1 = __dmul_rn(d1, d1), q2 = __dmul_rn(d2, d2), q3 = __dmul_rn(d3, d3);
            s = __dadd_rn(s, q0);
            s = __dadd_rn(s, q1);
            s = __dadd_rn(s, q2);
            s = __dadd_rn(s, q3);
        }
        for (; t < d; ++t) {
            const double diff = __dadd_rn(-(double)slot[x_slot_index(cell, t)], m[t]);
            s = __dadd_rn(s, __dmul_rn(diff, diff));
        }
        const double dist = sqrt(s);
        if (a == 0 && dist != dist) first_nan = true;
        if (dist < bv) {
            bv = dist;
            bi = a;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov < bv || (ov == bv && oi < bi)) {
            bv = ov;
            bi = oi;
        }
    }
    first_nan = __shfl_sync(0xffffffffu, (int)first_nan, 0) != 0;
    if (first_nan || bi == 0x7fffffff) bi = 0;   // nothing compares below a NaN / every distance is NaN or +inf: the first row stays
    return n_slots > 0 ? orig[bi] : 0;
}
// out of line: the path is rare and the pass kernels are short of registers (score_pass.cu, whose warp roles run under different
// register limits, inlines it into the role that calls it instead)
static __device__ __noinline__ int exact_label_warp(const float* __restrict__ slot, int cell, int d, const double* __restrict__ mu, int ld,
                                                    const int32_t* __restrict__ orig, int n_slots, int lane) {
    return exact_label_warp_inline(slot, cell, d, mu, ld, orig, n_slots, lane);
}

__device__ __forceinline__ void red_add(double* p, double v) { atomicAdd(p, v); }
__device__ __forceinline__ void red_add(long long* p, long long v) { atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)v); }

// ---------------------------------------------------------------------------------------------------
// Lloyd mode: signed contributions of one slot into the warp's private table.
// Each lane brings up to NE entries: (cell, label, sign); key = 2*label + (sign < 0).  off[0..2k] ints and
// perm[0..32*NE-1] bytes are the warp's scratch.  No atomics: one leader per key per instruction.
// ---------------------------------------------------------------------------------------------------
template <int NE>
__device__ __forceinline__ void slot_contributions(const float* __restrict__ slot, const int (&key)[NE], const int (&cell)[NE], int d,
                                                   int k, int lane, int* off, unsigned char* perm, double* __restrict__ sums,
                                                   long long* __restrict__ counts) {
    const unsigned full = 0xffffffffu;
    const int nk = 2 * k;
    bool any = false;
#pragma unroll
    for (int r = 0; r < NE; ++r) any |= key[r] >= 0;
    if (!__any_sync(full, any)) return;
    for (int i = lane; i <= nk; i += 32) off[i] = 0;
    __syncwarp();
    int pos[NE];
#pragma unroll
    for (int r = 0; r < NE; ++r) {
        const int L = key[r];
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
    // inclusive scan: off[i] = number of entries with key < i   (entry i held the count of key i-1)
    int carry = 0;
    for (int b = 0; b <= nk; b += 32) {
        const int i = b + lane;
        int s = (i <= nk) ? off[i] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(full, s, o);
            if (lane >= o) s += t;
        }
        if (i <= nk) off[i] = carry + s;
        carry += __shfl_sync(full, s, 31);
    }
    __syncwarp();
#pragma unroll
    for (int r = 0; r < NE; ++r)
        if (key[r] >= 0) perm[off[key[r]] + pos[r]] = (unsigned char)cell[r];
    __syncwarp();
    for (int L = lane; L < k; L += 32) {
        const int c = (off[2 * L + 1] - off[2 * L]) - (off[2 * L + 2] - off[2 * L + 1]);
        if (c) red_add(&counts[L], (long long)c);
    }
    for (int jb = 0; jb < d; jb += 32) {
        const int j = jb + lane;
        const bool on = j < d;
        const float* row = slot + x_slot_index(0, on ? j : 0);   // this lane's marker: cell c sits 4*c floats further
        for (int Lb = 0; Lb < k; Lb += 32) {
            const int Li = Lb + lane;
            const bool present = (Li < k) && (off[2 * Li + 2] > off[2 * Li]);
            unsigned pm = __ballot_sync(full, present);
            while (pm) {
                const int L = Lb + __ffs(pm) - 1;
                pm &= pm - 1;
                const int s = off[2 * L], mid = off[2 * L + 1], e = off[2 * L + 2];
                double* tcell = sums + (size_t)L * d + j;
                double a0 = 0.0, a1 = 0.0;
                int p = s;
                for (; p + 1 < mid; p += 2) {
                    a0 += (double)row[4 * perm[p]];
                    a1 += (double)row[4 * perm[p + 1]];
                }
                if (p < mid) a0 += (double)row[4 * perm[p]];
                double m0 = 0.0;
                for (p = mid; p < e; ++p) m0 += (double)row[4 * perm[p]];
                if (on) red_add(tcell, (a0 + a1) - m0);
            }
        }
    }
    __syncwarp();
}

// ---------------------------------------------------------------------------------------------------
// Lloyd delta mode, few moved cells: cell c leaves label oldL for newL.  Lane j owns markers j, j+32, ...
// The additions are reductions without return value (RED.ADD.F64): nothing waits on the L2 round trip.  They stay
// deterministic because a table belongs to one warp and a given address is only ever touched by one lane of it —
// operations of one thread on one address take effect in program order — so every table entry sees a fixed
// sequence of fp64 additions.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void move_one(const float* __restrict__ slot, int d, int lane, double* __restrict__ sums,
                                         long long* __restrict__ counts, int c, int newL, int oldL) {
    double* ta = sums + (size_t)newL * d;
    double* tb = sums + (size_t)oldL * d;
    for (int j = lane; j < d; j += 32) {
        const double v = (double)slot[x_slot_index(c, j)];
        red_add(ta + j, v);
        red_add(tb + j, -v);
    }
    if (lane < 2) red_add(&counts[lane ? oldL : newL], lane ? -1LL : 1LL);
}

}  // namespace dpr
