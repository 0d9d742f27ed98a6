// Design probe: 4 cells per lane (128-cell slots, 8 warps/SM, up to 255 registers) against the product's 2 cells per lane
// (64-cell slots, 16 warps/SM, 128 registers): the same inner FFMA2 loop + two-min fold on slots resident in shared memory.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o r4_bench r4_bench.cu && ./r4_bench
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ float fmin3(float a, float b, float c) { float r; asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float pack_idx(float s, int idx) { return __uint_as_float((__float_as_uint(s) & 0xFFFFFFE0u) | (unsigned)idx); }

template <int KC, int R>
__device__ __forceinline__ void chunk(const float* __restrict__ slot, const float* __restrict__ tab, int d, int lane, float (&best)[R], float (&second)[R]) {
    constexpr int NP = KC / 2, KCP = (KC + 3) & ~3, ROW = 32 * R;
    float2 acc[R][NP];
    const float2* cc = reinterpret_cast<const float2*>(tab);
#pragma unroll
    for (int t = 0; t < NP; ++t) {
        const float2 c = cc[t];
#pragma unroll
        for (int r = 0; r < R; ++r) acc[r][t] = c;
    }
    const float4* cw = reinterpret_cast<const float4*>(tab + KCP);
#pragma unroll 2
    for (int j = 0; j < d; ++j) {
        float x[R];
        if (R == 4) {
            const float4 xv = reinterpret_cast<const float4*>(slot)[j * (ROW / 4) + (lane ^ ((j >> 2) & 7))];
            x[0] = xv.x; x[1] = xv.y; x[2] = xv.z; x[3] = xv.w;
        } else {
            const float2 xv = reinterpret_cast<const float2*>(slot)[j * (ROW / 2) + (lane ^ ((j >> 2) & 15))];
            x[0] = xv.x; x[1] = xv.y;
        }
#pragma unroll
        for (int q = 0; q < KCP / 4; ++q) {
            const float4 c = cw[j * (KCP / 4) + q];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float2 xr = make_float2(x[r], x[r]);
                acc[r][2 * q] = __ffma2_rn(xr, make_float2(c.x, c.y), acc[r][2 * q]);
                if (2 * q + 1 < NP) acc[r][2 * q + 1] = __ffma2_rn(xr, make_float2(c.z, c.w), acc[r][2 * q + 1]);
            }
        }
    }
#pragma unroll
    for (int t = 0; t < NP; ++t)
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const float a = pack_idx(acc[r][t].x, 2 * t), b = pack_idx(acc[r][t].y, 2 * t + 1);
            const float lo = fminf(a, b), hi = fmaxf(a, b);
            second[r] = fmin3(fmaxf(best[r], lo), second[r], hi);
            best[r] = fminf(best[r], lo);
        }
}

template <int KC, int R, int THREADS>
__global__ void __launch_bounds__(THREADS, 1) k(float* out, int d, int slots_per_warp) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KCP = (KC + 3) & ~3, ROW = 32 * R;
    float* table = sm;
    float* slot = sm + (d + 1) * KCP + warp * d * ROW;
    for (int i = threadIdx.x; i < (d + 1) * KCP; i += blockDim.x) table[i] = 0.001f * (i % 97);
    for (int i = lane; i < d * ROW; i += 32) slot[i] = 0.01f * ((i * 7 + warp) % 113);
    __syncthreads();
    float acc_out = 0.f;
    for (int s = 0; s < slots_per_warp; ++s) {
        float best[R], second[R];
#pragma unroll
        for (int r = 0; r < R; ++r) best[r] = second[r] = 3e38f;
        chunk<KC, R>(slot, table, d, lane, best, second);
#pragma unroll
        for (int r = 0; r < R; ++r) acc_out += best[r] + second[r];
        table[0] = acc_out * 1e-30f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc_out;
}

template <int KC, int R, int THREADS>
void run(int d, int sms, float* out) {
    constexpr int KCP = (KC + 3) & ~3;
    const int warps = THREADS / 32;
    const size_t shm = sizeof(float) * ((size_t)(d + 1) * KCP + (size_t)warps * d * 32 * R);
    cudaFuncSetAttribute(k<KC, R, THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    const int slots = 200 * 2 / R;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k<KC, R, THREADS><<<sms, THREADS, shm>>>(out, d, slots);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        k<KC, R, THREADS><<<sms, THREADS, shm>>>(out, d, slots);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double fma = (double)sms * warps * slots * 32.0 * R * KC * d;
    printf("R=%d KC=%2d d=%2d warps=%2d smem=%3zuKB: %.2f TFMA/s   err=%s\n", R, KC, d, warps, shm >> 10, fma / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    run<30, 2, 512>(40, sms, out);
    run<30, 4, 256>(40, sms, out);
    run<30, 4, 320>(40, sms, out);
    run<16, 4, 512>(40, sms, out);
    run<16, 4, 256>(40, sms, out);
    run<20, 2, 512>(15, sms, out);
    run<20, 4, 256>(15, sms, out);
    run<20, 4, 512>(15, sms, out);
    run<8, 2, 512>(40, sms, out);
    run<8, 4, 512>(40, sms, out);
    run<8, 4, 256>(40, sms, out);
    run<22, 2, 512>(40, sms, out);
    run<22, 4, 256>(40, sms, out);
    return 0;
}
