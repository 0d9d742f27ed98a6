/*
 * dpr_philox.h — the randomness contract shared by the CUDA library and the CPU oracle.
 *
 * The reference seeds libc rand() with time(0)+offset (src/Clusterer.cpp:25,
 * src/Clusterer.h:51-53), so its runs are unrepeatable by construction.  The
 * replacement takes an explicit 64-bit seed and draws every random number from
 * Philox4x32-10 addressed by (purpose, a, b, c): a pure function of its inputs,
 * so host, device and oracle agree bit for bit without sharing any state.
 *
 * Draw catalogue (the only places the reference consumes randomness):
 *   DPR_RNG_BOOT  a = fit id, b = draw i           -> row = w0 % n      (Clusterer.cpp:314-316)
 *   DPR_RNG_SEED  a = fit id, b = 0                -> first centre = w0 % n            (:156)
 *                 a = fit id, b = t (t >= 1)       -> u = w0 * 2^-32, x = u * cum      (:136)
 *   DPR_RNG_PAIR  a = pair id, b = num, c = attempt-> (i, j) = (w0 % n, w1 % n), redraw while i == j (:291-295)
 *
 * Usable from C, C++ and CUDA (define DPR_HD before including for __host__ __device__).
 */
#ifndef DPR_PHILOX_H
#define DPR_PHILOX_H

#include <stdint.h>

#ifndef DPR_HD
#ifdef __CUDACC__
#define DPR_HD __host__ __device__ __forceinline__
#else
#define DPR_HD static inline
#endif
#endif

enum { DPR_RNG_BOOT = 1, DPR_RNG_SEED = 2, DPR_RNG_PAIR = 3 };

typedef struct dpr_u32x4 { uint32_t w[4]; } dpr_u32x4;

DPR_HD void dpr_mulhilo32(uint32_t a, uint32_t b, uint32_t* hi, uint32_t* lo) {
    uint64_t p = (uint64_t)a * (uint64_t)b;
    *hi = (uint32_t)(p >> 32);
    *lo = (uint32_t)p;
}

/* Philox4x32-10 (Salmon et al., SC'11): counter (c0..c3), key (k0,k1). */
DPR_HD dpr_u32x4 dpr_philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                   uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0, lo0, hi1, lo1;
        dpr_mulhilo32(M0, c0, &hi0, &lo0);
        dpr_mulhilo32(M1, c2, &hi1, &lo1);
        uint32_t n0 = hi1 ^ c1 ^ k0;
        uint32_t n1 = lo1;
        uint32_t n2 = hi0 ^ c3 ^ k1;
        uint32_t n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += W0; k1 += W1;
    }
    dpr_u32x4 out;
    out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
    return out;
}

/* One addressed draw: four 32-bit words for (seed, purpose, a, b, c). */
DPR_HD dpr_u32x4 dpr_rng_draw(uint64_t seed, uint32_t purpose, uint32_t a, uint32_t b, uint32_t c) {
    return dpr_philox4x32_10(a, b, c, purpose, (uint32_t)seed, (uint32_t)(seed >> 32));
}

/* u in [0,1): the reference's rand()/RAND_MAX is in [0,1]; the closed end only selects
 * the degenerate map[x] branch of element_from_vector (Clusterer.cpp:141-143). */
DPR_HD double dpr_rng_unit(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

#endif /* DPR_PHILOX_H */
