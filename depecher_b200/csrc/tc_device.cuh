// tc_device.cuh — tcgen05 / tensor-memory / mbarrier PTX wrappers and the fp16 operand split shared by the tensor-core kernels
// (tc_pass.cu: the Lloyd / assignment pass; score_pass.cu: the scoring pass of the penalty search)
#pragma once
#include <cuda_fp16.h>

#include "common.cuh"
#include "pass_device.cuh"

namespace dpr {

// ---------------------------------------------------------------------------------------------------
// tcgen05 PTX
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
template <int THREADS>
__device__ __forceinline__ void consumer_barrier() { asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory"); }   // all consumer warps
// wait with a suspend-time hint: the warp sleeps in hardware between polls instead of burning issue slots
__device__ __forceinline__ void mbar_wait_hint(uint64_t* bar, uint32_t parity, uint32_t hint_ns) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity), "r"(hint_ns)
        : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {   // non-blocking
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, no swizzle: 8 rows x 16 B core matrices; lbo = bytes between the two 16 B halves of a K=8 step,
// sbo = bytes between 8-row blocks
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void umma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&w)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]),
                 "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7])
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
// 16 accumulator columns of this warp's 32 lanes; the registers are valid after tmem_wait_ld16 (which names them, so that no use
// can be scheduled above the wait)
__device__ __forceinline__ void tmem_ld16_raw(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
          "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld16(uint32_t (&v)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]), "+r"(v[9]), "+r"(v[10]),
                   "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15]));
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
// v (already scaled) = hi + lo + r with hi, lo fp16 and |r| <= 2^-22 |v| + 2^-25; two values per 32-bit TMEM column
// (the element with the lower K index in the low half)
__device__ __forceinline__ void split_pair(float v0, float v1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(v0, v1);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ---------------------------------------------------------------------------------------------------
// Sortable integer keys (tc_pass.cu, score_pass.cu).  With q a power of two in (2^-21, 2^-20] x (the largest |score| under the
// centre set) and M = 1.5 * 2^23 * q,
//     w   = fma(unscale, D, cc + M)     the score rounded to a multiple of q, plus M: every w lies in the binade of M, so its bit
//                                       pattern is  const + score / q  (one rounding)
//     key = bits(w) * 16 + i            i = the column's number in its block of 16; int32 arithmetic wraps, but the keys of one
//                                       binade fill one 2^27-aligned block of the integers: they never straddle the sign change
// Signed integer min / max on the keys order the scores to within q and carry the winner's column; a difference of two keys
// times q / 16 is the difference of the scores up to 3 q (cc + M rounds cc: q / 2; the fma: q / 2; twice; the column numbers: < q).
// key_scale: q for a set whose largest |score| is bounded by smax (0: no usable scale — decide every cell exactly)
__device__ __forceinline__ float key_quantum(float smax) {
    int e2 = 0;
    frexpf(smax, &e2);                                   // smax < 2^e2
    return (smax > 0.f && smax < 1e30f && e2 > -90) ? ldexpf(1.f, e2 - 21) : 0.f;   // |score| / q < 2^21: M +- 2^22 q is the binade of M
}
// ||c||^2 + M of a padding column, in units of q above M: above every real score (< 2^21), inside the binade (< 2^22)
constexpr float kKeyPad = 4194240.f;   // 2^22 - 64
// one block of 16 accumulator columns -> (smallest key, second smallest key).  cc4: the block's ||c||^2 + M (padding columns:
// M + 2^20 q).  The smallest: a tree of 3-input mins.  The second smallest: the keys are distinct (column number inside), so
// key - b - 1 as an UNSIGNED number is huge for the smallest key and (key - b - 1) for every other one: second = b + 1 +
// unsigned-min of those (add + min fuse into one instruction, VIADDMNMX).  mul16 = 16 as a run-time value: a register factor stays
// an IMAD on the FMA pipe, a constant 16 becomes LEA on the half-rate ALU pipe that the mins already load.
__device__ __forceinline__ void top2_block16(const uint32_t (&v)[16], const float4* __restrict__ cc4, float unscale, uint32_t mul16, int& b, int& sec) {
    int w[16];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float4 cq = cc4[j];
        w[4 * j + 0] = (int)(__float_as_uint(fmaf(unscale, __uint_as_float(v[4 * j + 0]), cq.x)) * mul16 + (unsigned)(4 * j + 0));
        w[4 * j + 1] = (int)(__float_as_uint(fmaf(unscale, __uint_as_float(v[4 * j + 1]), cq.y)) * mul16 + (unsigned)(4 * j + 1));
        w[4 * j + 2] = (int)(__float_as_uint(fmaf(unscale, __uint_as_float(v[4 * j + 2]), cq.z)) * mul16 + (unsigned)(4 * j + 2));
        w[4 * j + 3] = (int)(__float_as_uint(fmaf(unscale, __uint_as_float(v[4 * j + 3]), cq.w)) * mul16 + (unsigned)(4 * j + 3));
    }
    b = min(__vimin3_s32(__vimin3_s32(w[0], w[1], w[2]), __vimin3_s32(w[3], w[4], w[5]), __vimin3_s32(w[6], w[7], w[8])),
            __vimin3_s32(__vimin3_s32(w[9], w[10], w[11]), __vimin3_s32(w[12], w[13], w[14]), w[15]));
    const int nb1 = -(b + 1);
    unsigned u[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) u[j] = (unsigned)(w[j] + nb1);
    const unsigned um = min(__vimin3_u32(__vimin3_u32(u[0], u[1], u[2]), __vimin3_u32(u[3], u[4], u[5]), __vimin3_u32(u[6], u[7], u[8])),
                            __vimin3_u32(__vimin3_u32(u[9], u[10], u[11]), __vimin3_u32(u[12], u[13], u[14]), u[15]));
    sec = (int)((unsigned)(b + 1) + um);
}

// index of a column inside its 16-column block, in the 4 low mantissa bits of the score
__device__ __forceinline__ float pack_idx4(float s, int idx) { return __uint_as_float((__float_as_uint(s) & 0xFFFFFFF0u) | (unsigned)idx); }
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&w)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
}
// one cell's markers -> fp16 high / low parts in tensor memory; 8 markers (2 marker groups of the tile, 4 TMEM columns
// per operand) per step.  st4 points at the cell's 16 bytes inside marker group 0.  Columns past the last step belong
// to the zero padding of the last K=16 step: written once at kernel start, never again.
template <bool SCALED>
__device__ __forceinline__ void split_group(const float4* __restrict__ st4, float scale, uint32_t& hi0, uint32_t& hi1, uint32_t& lo0, uint32_t& lo1) {
    float4 v = *st4;   // 4 markers of the cell
    if (SCALED) {
        v.x *= scale; v.y *= scale; v.z *= scale; v.w *= scale;
    }
    split_pair(v.x, v.y, hi0, lo0);
    split_pair(v.z, v.w, hi1, lo1);
}
template <bool SCALED>
__device__ __forceinline__ void split_tile(const float4* __restrict__ st4, int steps, float scale, uint32_t a_hi, uint32_t a_lo) {
    constexpr int G4 = kGroupFloats / 4;   // float4 stride between marker groups
    int g = 0;
    for (; g + 1 < steps; g += 2, st4 += 4 * G4, a_hi += 8, a_lo += 8) {   // 16 markers -> 8 columns per operand
        uint32_t hi[8], lo[8];
#pragma unroll
        for (int q = 0; q < 4; ++q) split_group<SCALED>(st4 + q * G4, scale, hi[2 * q], hi[2 * q + 1], lo[2 * q], lo[2 * q + 1]);
        tmem_st8(a_hi, hi);
        tmem_st8(a_lo, lo);
    }
    if (g < steps) {   // 8 more markers -> 4 columns
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 2; ++q) split_group<SCALED>(st4 + q * G4, scale, hi[2 * q], hi[2 * q + 1], lo[2 * q], lo[2 * q + 1]);
        tmem_st4(a_hi, hi);
        tmem_st4(a_lo, lo);
    }
}

}  // namespace dpr
