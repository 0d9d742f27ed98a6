// tc_probe_f16.cu — error of the shipped product scheme: x = xh + xl, c = ch + cl (all fp16), D = xh*[ch|cl] + xl*ch on
// tcgen05.mma kind::f16 (fp32 accumulate, A operands in tensor memory, B N-stacked in shared memory), against fp64.
// Reports max |D0 + D1 - x.c| / sum|x_j c_j| for random-sign, all-positive (truncation bias adds up) and wide-range rows.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tc_probe_f16 tc_probe_f16.cu && ./tc_probe_f16
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                       \
    do {                                                                            \
        cudaError_t e_ = (x);                                                       \
        if (e_ != cudaSuccess) {                                                    \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); \
            exit(1);                                                                \
        }                                                                           \
    } while (0)

constexpr int M = 128;
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
    return (uint64_t)((addr >> 4) & 0x3fff) | ((uint64_t)((lbo >> 4) & 0x3fff) << 16) | ((uint64_t)((sbo >> 4) & 0x3fff) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, {%5, %5, %5, %5}, p;\n\t}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
        : "memory");
}

// N centres (multiple of 16, <= 64), d markers; B operand: rows [0,N) high parts, [N,2N) low parts
__global__ void __launch_bounds__(128, 1) probe(const float* __restrict__ x, const float* __restrict__ c, int d, int N, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(8) uint64_t bar;
    const int d16 = (d + 15) & ~15, g16 = d16 / 8;
    __half* b = reinterpret_cast<__half*>(smem);   // half (row, k) at (k/8) * 16*N + 8*row + k%8  (2N rows)
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < g16 * 2 * N * 8; i += 128) b[i] = __float2half(0.f);
    __syncthreads();
    for (int i = tid; i < N * d; i += 128) {
        const int n = i / d, k = i % d;
        const float v = c[i];
        const __half hi = __float2half_rn(v);
        const __half lo = __float2half_rn(v - __half2float(hi));
        const size_t at = (size_t)(k / 8) * (16 * N) + (size_t)n * 8 + (k % 8);
        b[at] = hi;
        b[at + (size_t)N * 8] = lo;
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_base_s, ah_col = 256, al_col = 384;
    const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);
    for (int k0 = 0; k0 < d16; k0 += 8) {   // 8 markers -> 4 columns per operand
        uint32_t h[4], l[4];
        for (int q = 0; q < 4; ++q) {
            const int k = k0 + 2 * q;
            const float v0 = k < d ? x[(size_t)tid * d + k] : 0.f, v1 = k + 1 < d ? x[(size_t)tid * d + k + 1] : 0.f;
            const __half2 hh = __floats2half2_rn(v0, v1);
            const float2 hf = __half22float2(hh);
            const __half2 ll = __floats2half2_rn(v0 - hf.x, v1 - hf.y);
            h[q] = *reinterpret_cast<const uint32_t*>(&hh);
            l[q] = *reinterpret_cast<const uint32_t*>(&ll);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lane_addr + ah_col + k0 / 2), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
        asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(lane_addr + al_col + k0 / 2), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
    }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t idesc0 = (1u << 4) | ((uint32_t)(M >> 4) << 24);
    if (tid == 0) {
        const uint32_t lbo = (uint32_t)N * 32u;
        for (int s = 0; s < g16 / 2; ++s) {
            const uint64_t bd = make_desc(smem_u32(b) + s * 2 * lbo, lbo, 128);
            mma_ts(tmem, tmem + ah_col + 8 * s, bd, idesc0 | ((uint32_t)((2 * N) >> 3) << 17), s > 0);
            mma_ts(tmem, tmem + al_col + 8 * s, bd, idesc0 | ((uint32_t)(N >> 3) << 17), 1);
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    {
        uint32_t done = 0;
        while (!done)
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0u) : "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    for (int n0 = 0; n0 < N; n0 += 8) {
        uint32_t v[8], w[8];
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]) : "r"(lane_addr + n0) : "memory");
        asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                     : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]) : "r"(lane_addr + N + n0) : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        for (int q = 0; q < 8; ++q) out[(size_t)tid * N + n0 + q] = __uint_as_float(v[q]) + __uint_as_float(w[q]);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

int main() {
    const int ds[] = {40, 15, 64, 50, 7};
    const int Ns[] = {32, 16, 64, 32, 48};
    const char* kinds[3] = {"random sign", "all positive", "wide range"};
    srand(11);
    auto u = []() { return rand() / (double)RAND_MAX; };
    for (int t = 0; t < 5; ++t)
        for (int kind = 0; kind < 3; ++kind) {
            const int d = ds[t], N = Ns[t], d16 = (d + 15) & ~15;
            std::vector<float> x((size_t)M * d), c((size_t)N * d);
            for (size_t i = 0; i < x.size(); ++i) {
                double v = kind == 1 ? 0.5 + 7.0 * u() : (u() - 0.5) * 16.0;
                if (kind == 2) v *= std::pow(2.0, (int)(u() * 12) - 6);
                x[i] = (float)v;
            }
            for (size_t i = 0; i < c.size(); ++i) {
                double v = kind == 1 ? 0.5 + 5.0 * u() : (u() - 0.5) * 12.0;
                if (kind == 2) v *= std::pow(2.0, (int)(u() * 12) - 6);
                c[i] = (float)v;
            }
            float *dx, *dc, *dout;
            CK(cudaMalloc(&dx, x.size() * 4));
            CK(cudaMalloc(&dc, c.size() * 4));
            CK(cudaMalloc(&dout, (size_t)M * N * 4));
            CK(cudaMemcpy(dx, x.data(), x.size() * 4, cudaMemcpyHostToDevice));
            CK(cudaMemcpy(dc, c.data(), c.size() * 4, cudaMemcpyHostToDevice));
            const size_t smem = (size_t)(d16 / 8) * 2 * N * 16;
            CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            probe<<<1, 128, smem>>>(dx, dc, d, N, dout);
            CK(cudaGetLastError());
            CK(cudaDeviceSynchronize());
            std::vector<float> r((size_t)M * N);
            CK(cudaMemcpy(r.data(), dout, r.size() * 4, cudaMemcpyDeviceToHost));
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
                    worst = std::max(worst, std::fabs(r[(size_t)i * N + n] - ref) / S);
                    worst32 = std::max(worst32, std::fabs((double)f32 - ref) / S);
                }
            const int steps = d16 / 16;
            printf("d=%2d N=%2d %-12s: max |err|/S = %.3e  (fp32 fma chain %.3e)   shipped bound eps = %.3e\n", d, N, kinds[kind], worst, worst32,
                   3 * 2.3841858e-7 + (2 * steps + 1) * 2.3841858e-7 + 4 * 5.9604645e-8);
            cudaFree(dx);
            cudaFree(dc);
            cudaFree(dout);
        }
    return 0;
}
