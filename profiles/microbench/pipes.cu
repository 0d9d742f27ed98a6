// Microbenchmark of the issue/pipe rates that bound the distance micro-kernel (design input, not product).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

template <int NACC>
__global__ void k_ffma(float* out, float a, float b, int iters) {
    float acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_ffma2(float* out, float a, float b, int iters) {
    float2 acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 av = make_float2(a, a * 1.0001f), bv = make_float2(b, b * 0.999f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = __ffma2_rn(acc[i], av, bv);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// FFMA mixed with ALU work (min/max tracking) to see whether FMNMX steals FMA issue
template <int NACC>
__global__ void k_ffma_mix(float* out, float a, float b, int iters) {
    float acc[NACC];
    float best = 1e30f, second = 1e30f;
#pragma unroll
    for (int i = 0; i < NACC; ++i) acc[i] = threadIdx.x * 1e-3f + i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) acc[i] = fmaf(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < NACC / 8; ++i) {  // 1 min/max pair per 8 FFMA
            second = fminf(second, fmaxf(acc[i], best));
            best = fminf(best, acc[i]);
        }
    }
    float s = best + second;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// broadcast LDS.128 + FFMA: the centre-operand pattern (every lane reads the same 16 bytes)
template <int R, int KC, bool PACKED>
__global__ void k_tile(float* out, const float* __restrict__ cen, int d, int iters) {
    extern __shared__ float4 sm[];
    float4* cs = sm;                       // [d][KC/4] float4
    float4* xs = sm + d * (KC / 4);        // [d][blockDim.x] float4 (4 cells per thread)
    for (int i = threadIdx.x; i < d * (KC / 4); i += blockDim.x) cs[i] = reinterpret_cast<const float4*>(cen)[i];
    for (int i = threadIdx.x; i < d * blockDim.x; i += blockDim.x) xs[i] = make_float4(i * 1e-4f, 1.f, 2.f, 3.f);
    __syncthreads();
    float2 acc[R][KC / 2];
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < KC / 2; ++k) acc[r][k] = make_float2(0.f, 0.f);
    for (int it = 0; it < iters; ++it) {
#pragma unroll 2
        for (int j = 0; j < d; ++j) {
            float4 xv = xs[j * blockDim.x + threadIdx.x];
            float x[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
            for (int q = 0; q < KC / 4; ++q) {
                float4 c = cs[j * (KC / 4) + q];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (PACKED) {
                        float2 xx = make_float2(x[r], x[r]);
                        acc[r][2 * q] = __ffma2_rn(xx, make_float2(c.x, c.y), acc[r][2 * q]);
                        acc[r][2 * q + 1] = __ffma2_rn(xx, make_float2(c.z, c.w), acc[r][2 * q + 1]);
                    } else {
                        acc[r][2 * q].x = fmaf(x[r], c.x, acc[r][2 * q].x);
                        acc[r][2 * q].y = fmaf(x[r], c.y, acc[r][2 * q].y);
                        acc[r][2 * q + 1].x = fmaf(x[r], c.z, acc[r][2 * q + 1].x);
                        acc[r][2 * q + 1].y = fmaf(x[r], c.w, acc[r][2 * q + 1].y);
                    }
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int r = 0; r < R; ++r)
#pragma unroll
        for (int k = 0; k < KC / 2; ++k) s += acc[r][k].x + acc[r][k].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_dadd(double* out, const float* __restrict__ in, int iters) {
    double acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = i;
    float v = in[threadIdx.x & 31];
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { acc[i] += (double)v; v = __int_as_float(__float_as_int(v) ^ (it & 1)); }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float time_ms(F f, int reps = 5) {
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        cudaEventRecord(a);
        f();
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int sms = p.multiProcessorCount;
    printf("device %s sms %d clock %d kHz\n", p.name, sms, p.clockRate);
    float* out; double* dout; float* cen;
    CK(cudaMalloc(&out, sizeof(float) * sms * 32 * 1024));
    CK(cudaMalloc(&dout, sizeof(double) * sms * 32 * 1024));
    CK(cudaMalloc(&cen, 1 << 20));
    CK(cudaMemset(cen, 0, 1 << 20));
    const int iters = 20000;
    for (int warps = 4; warps <= 16; warps *= 2) {
        int thr = warps * 32;
        {
            float ms = time_ms([&] { k_ffma<32><<<sms, thr>>>(out, 1.0001f, 0.5f, iters); });
            double fma = (double)sms * thr * 32.0 * iters;
            printf("ffma   scalar warps/SM %2d: %.1f GFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", warps, fma / ms * 1e-6, fma / ms * 1e-6 / sms / 1.965);
        }
        {
            float ms = time_ms([&] { k_ffma2<32><<<sms, thr>>>(out, 1.0001f, 0.5f, iters); });
            double fma = (double)sms * thr * 64.0 * iters;
            printf("ffma2  packed warps/SM %2d: %.1f GFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", warps, fma / ms * 1e-6, fma / ms * 1e-6 / sms / 1.965);
        }
        {
            float ms = time_ms([&] { k_ffma_mix<32><<<sms, thr>>>(out, 1.0001f, 0.5f, iters); });
            double fma = (double)sms * thr * 32.0 * iters;
            printf("ffma+fmnmx(1:4)  warps/SM %2d: %.1f GFMA/s\n", warps, fma / ms * 1e-6);
        }
        {
            float ms = time_ms([&] { k_dadd<<<sms, thr>>>(dout, cen, iters); });
            double ops = (double)sms * thr * 8.0 * iters;
            printf("f2f+dadd warps/SM %2d: %.1f Gop/s (%.1f /clk/SM)\n", warps, ops / ms * 1e-6, ops / ms * 1e-6 / sms / 1.965);
        }
    }
    const int d = 40, it2 = 400;
#define TILE(R, KC, P, THR)                                                                                     \
    {                                                                                                           \
        size_t sh = sizeof(float4) * (d * (KC / 4) + d * THR);                                                  \
        cudaFuncSetAttribute(k_tile<R, KC, P>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);           \
        float ms = time_ms([&] { k_tile<R, KC, P><<<sms, THR, sh>>>(out, cen, d, it2); });                      \
        double fma = (double)sms * THR * R * KC * d * it2;                                                      \
        printf("tile R=%d KC=%2d %s thr %3d: %.1f GFMA/s (%.1f FMA/clk/SM @1.965)\n", R, KC, P ? "packed" : "scalar", THR, \
               fma / ms * 1e-6, fma / ms * 1e-6 / sms / 1.965);                                                 \
    }
    TILE(4, 16, false, 128) TILE(4, 16, true, 128) TILE(4, 16, false, 256) TILE(4, 16, true, 256)
    TILE(4, 8, false, 128) TILE(4, 8, true, 128) TILE(4, 8, false, 256) TILE(4, 8, true, 256)
    TILE(4, 32, false, 128) TILE(4, 32, true, 128) TILE(4, 32, true, 256)
    TILE(2, 16, true, 256) TILE(2, 32, true, 256) TILE(2, 32, false, 256)
    CK(cudaDeviceSynchronize());
    printf("done\n");
    return 0;
}
