// tc_probe.cu — does a 3xTF32 tcgen05 product reproduce fp32-grade dot products for the distance step?
//
// One CTA, one tile: 128 cells x d markers (A) against N centres (B).  A comes from TMEM (mode 0) or from shared
// memory (mode 1) in the K-major no-swizzle canonical layout; B always from shared memory.  The three products
// xl*ch, xh*cl, xh*ch are accumulated into one TMEM accumulator.  The host compares against fp64 and reports the
// largest error relative to S = sum_j |x_j||c_j|, and whether feeding raw fp32 words instead of explicitly
// masked ones changes the result (i.e. whether kind::tf32 truncates the low 13 mantissa bits).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe tc_probe.cu && ./tc_probe
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                        \
    do {                                                                             \
        cudaError_t e_ = (x);                                                        \
        if (e_ != cudaSuccess) {                                                     \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__);  \
            exit(1);                                                                 \
        }                                                                            \
    } while (0)

constexpr int M = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) |
           (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}

// element (row, k) of a K-major no-swizzle operand: 8 rows x 4 elements (16 B) per core matrix
__host__ __device__ inline int op_index(int row, int k, int lbo_floats) { return (k >> 2) * lbo_floats + (row >> 3) * 32 + (row & 7) * 4 + (k & 3); }

// number of batches committed before `variant` in the throughput section (parity of the mbarrier)
__device__ __forceinline__ int variant_count_done(int variant, int N) {
    int c = 0;
    for (int v = 0; v < variant; ++v) {
        const int Nv = v < 2 ? N : (v < 4 ? 2 * N : 4 * N);
        if (Nv > 256) continue;
        ++c;
    }
    return c;
}

// mode bit 0: A from shared memory instead of TMEM;  bit 1: feed raw x instead of masked xh
__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ x, const float* __restrict__ c, int d, int N, int mode, float* out, long long* clk) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int d8 = (d + 7) & ~7;
    const int b_lbo = (N / 8) * 32;          // floats between 4-marker groups of B
    const int a_lbo = 16 * 32 + 4;           // floats between 4-marker groups of A (padded by 16 B)
    float* bh = reinterpret_cast<float*>(smem);
    float* bl = bh + (d8 / 4) * b_lbo;
    float* ah = bl + (d8 / 4) * b_lbo;
    float* al = ah + (d8 / 4) * a_lbo;
    const int tid = threadIdx.x, warp = tid >> 5;

    for (int i = tid; i < (d8 / 4) * b_lbo * 2 + (d8 / 4) * a_lbo * 2; i += 128) bh[i] = 0.f;
    __syncthreads();
    for (int i = tid; i < N * d; i += 128) {
        const int n = i / d, k = i % d;
        const float v = c[i];
        const float h = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        bh[op_index(n, k, b_lbo)] = h;
        bl[op_index(n, k, b_lbo)] = v - h;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"((mode & 8) ? 2u : 1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s;
    const uint32_t acc_col = 0, ah_col = 256, al_col = 384;

    const long long c0 = clock64();
    // this thread's cell row: split and park as the A operand
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < d8; k0 += 8) {
        uint32_t h[8], l[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int k = k0 + q;
            const float v = k < d ? x[(size_t)tid * d + k] : 0.f;
            const float vh = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
            h[q] = (mode & 2) ? __float_as_uint(v) : __float_as_uint(vh);
            l[q] = __float_as_uint(v - vh);
            ah[op_index(tid, k, a_lbo)] = __uint_as_float(h[q]);
            al[op_index(tid, k, a_lbo)] = __uint_as_float(l[q]);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + ah_col + k0), "r"(h[0]),
                     "r"(h[1]), "r"(h[2]), "r"(h[3]), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7])
                     : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(lane_addr + al_col + k0), "r"(l[0]),
                     "r"(l[1]), "r"(l[2]), "r"(l[3]), "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7])
                     : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    const long long c1 = clock64();
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes of the smem operands -> tensor core
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // instruction descriptor: D fp32, A/B tf32, both K-major, N, M=128
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
    const long long c2 = clock64();
    long long c3 = 0;
    if (tid == 0 || ((mode & 8) && tid == 32)) {
        const uint32_t acc_shift = tid == 32 ? 128u : 0u;   // second issuer: its own accumulator
        uint32_t accumulate = 0;
        for (int term = 0; term < 3; ++term) {   // xl*ch, xh*cl, xh*ch
            const float* bsrc = term == 1 ? bl : bh;
            const float* asrc = term == 0 ? al : ah;
            const uint32_t acol = term == 0 ? al_col : ah_col;
            for (int s = 0; s < d8 / 8; ++s) {
                const uint64_t bdesc = make_desc(smem_u32(bsrc + 2 * s * b_lbo), b_lbo * 4, 128);
                if (mode & 1) {
                    const uint64_t adesc = make_desc(smem_u32(asrc + 2 * s * a_lbo), a_lbo * 4, 128);
                    mma_ss(tmem + acc_col, adesc, bdesc, idesc, accumulate);
                } else {
                    mma_ts(tmem + acc_col + acc_shift + ((mode & 4) ? (uint32_t)(term * 64 + (s & 1) * 32) : 0u), tmem + acol + 8 * s, bdesc, idesc, (mode & 4) ? (uint32_t)(s > 1) : accumulate);
                }
                accumulate = 1;
            }
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
        c3 = clock64();
    }
    {   // everyone waits for the products
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    const long long c4 = clock64();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (tid == 0) {
        clk[0] = c1 - c0;   // split + tcgen05.st + wait::st (includes the global loads of x here)
        clk[1] = c3 - c2;   // issue of all products + commit
        clk[2] = c4 - c2;   // issue .. products complete (mbarrier observed)
    }
    for (int n0 = 0; n0 < N; n0 += 8) {
        uint32_t v[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                     : "r"(lane_addr + acc_col + n0)
                     : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 8; ++q) out[(size_t)tid * N + n0 + q] = __uint_as_float(v[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // ---- throughput of back-to-back products issued from uniform registers: variants of (A source, N)
    if (warp == 0 && mode == 0) {
        for (int variant = 0; variant < 6; ++variant) {
            const bool ss = variant & 1;
            const int Nv = variant < 2 ? N : (variant < 4 ? 2 * N : 4 * N);
            if (Nv > 256) continue;
            const uint32_t idv = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(Nv >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
            const uint64_t bd0 = make_desc(smem_u32(bh), b_lbo * 4, 128), ad0 = make_desc(smem_u32(ah), a_lbo * 4, 128);
            const long long t0 = clock64();
            uint32_t pred = 0;
            asm volatile("{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\telect.sync rx|px, %1;\n\t@px mov.s32 %0, 1;\n\t}\n" : "+r"(pred) : "r"(0xffffffffu));
            if (pred) {
                for (int rep = 0; rep < 8; ++rep)
                    for (int s = 0; s < d8 / 8; ++s) {
                        if (ss) mma_ss(tmem + acc_col, ad0 + (uint64_t)s * ((2 * a_lbo * 4) >> 4), bd0, idv, 1);
                        else mma_ts(tmem + acc_col, tmem + ah_col + 8 * s, bd0, idv, 1);
                    }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
            }
            __syncwarp();
            const long long t1 = clock64();
            {
                uint32_t done = 0;
                const uint32_t par = (uint32_t)((variant_count_done(variant, N) + 1) & 1);
                while (!done)
                    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(par) : "memory");
            }
            const long long t2 = clock64();
            if (tid == 0) {
                clk[4 + 2 * variant] = t1 - t0;
                clk[5 + 2 * variant] = t2 - t0;
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    const int ds[] = {40, 15, 50, 8, 3};
    const int Ns[] = {32, 16, 64, 256, 48};
    srand(7);
    for (int t = 0; t < 5; ++t) {
        const int d = ds[t], N = Ns[t], d8 = (d + 7) & ~7;
        std::vector<float> x((size_t)M * d), c((size_t)N * d);
        for (auto& v : x) v = (float)((rand() / (double)RAND_MAX - 0.5) * 8.0);
        for (auto& v : c) v = (float)((rand() / (double)RAND_MAX - 0.5) * 6.0);
        float *dx, *dc, *dout;
        long long* dclk;
        CK(cudaMalloc(&dclk, 256));
        CK(cudaMemset(dclk, 0, 256));
        CK(cudaMalloc(&dx, x.size() * 4));
        CK(cudaMalloc(&dc, c.size() * 4));
        CK(cudaMalloc(&dout, (size_t)M * N * 4));
        CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
        const size_t smem = (size_t)(d8 / 4) * ((N / 8) * 32 * 2 + (16 * 32 + 4) * 2) * 4;
        CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        std::vector<float> res[5];
        for (int mode = 0; mode <= (N <= 32 ? 8 : 3); mode = (mode == 4 ? 8 : mode + 1)) {
            CK(cudaMemset(dout, 0xff, (size_t)M * N * 4));
            probe<<<1, 128, smem>>>(dx, dc, d, N, mode, dout, dclk);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            long long hclk[3];
            CK(cudaMemcpy(hclk, dclk, 24, cudaMemcpyDeviceToHost));
            printf("   cycles: split+st %lld, issue %lld, issue..done %lld\n", hclk[0], hclk[1], hclk[2]);
            if (mode == 0) {
                long long h2[16];
                CK(cudaMemcpy(h2, dclk, 128, cudaMemcpyDeviceToHost));
                const char* names[6] = {"TS N", "SS N", "TS 2N", "SS 2N", "TS 4N", "SS 4N"};
                for (int v = 0; v < 6; ++v)
                    if (h2[5 + 2 * v]) printf("   %d products %s: issue %lld cycles, done %lld cycles (%.1f per product)\n", 8 * (d8 / 8), names[v], h2[4 + 2 * v], h2[5 + 2 * v], (double)h2[5 + 2 * v] / (8 * (d8 / 8)));
            }
            if (mode >= 4) continue;
            res[mode].resize((size_t)M * N);
            CK(cudaMemcpy(res[mode].data(), dout, (size_t)M * N * 4, cudaMemcpyDeviceToHost));
            double worst = 0, worst32 = 0;
            for (int i = 0; i < M; ++i)
                for (int n = 0; n < N; ++n) {
                    double ref = 0, S = 0;
                    float f32 = 0.f;
                    for (int k = 0; k < d; ++k) {
                        ref += (double)x[(size_t)i * d + k] * c[(size_t)n * d + k];
                        S += std::fabs((double)x[(size_t)i * d + k] * c[(size_t)n * d + k]);
                        f32 = fmaf(x[(size_t)i * d + k], c[(size_t)n * d + k], f32);
                    }
                    worst = std::max(worst, std::fabs(res[mode][(size_t)i * N + n] - ref) / S);
                    worst32 = std::max(worst32, std::fabs((double)f32 - ref) / S);
                }
            printf("d=%2d N=%3d mode=%d (%s A, %s xh): max |err|/S = %.3e   (plain fp32 fma chain: %.3e; 2^-24 = 5.96e-8)\n", d, N, mode,
                   (mode & 1) ? "smem" : "TMEM", (mode & 2) ? "raw " : "masked", worst, worst32);
        }
        size_t diff_raw = 0, diff_ss = 0;
        for (size_t i = 0; i < res[0].size(); ++i) {
            diff_raw += res[0][i] != res[2][i];
            diff_ss += res[0][i] != res[1][i];
        }
        printf("   raw vs masked xh: %zu of %zu entries differ;  TMEM-A vs smem-A: %zu differ\n", diff_raw, res[0].size(), diff_ss);
        cudaFree(dx);
        cudaFree(dc);
        cudaFree(dout);
    }
    return 0;
}
