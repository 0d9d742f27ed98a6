// Ceiling of the compute phase of fused_pass_kernel: the product's own dist_chunk<KC, true> (inner FFMA2 loop + two-min
// fold) on slots that are already in shared memory — no TMA, no labels, no rechecks.  Design input, not product.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o chunk_bench chunk_bench.cu && ./chunk_bench
#include <cstdio>
#include "../../depecher_b200/csrc/fused_pass.cu"

namespace dpr {
Context& ctx() { static Context c; return c; }
int fail(int code, const std::string&) { return code; }
void set_error(const std::string&) {}
int ensure_ready() { return 0; }
}
using namespace dpr;

// software-pipelined variant: operands of marker j+1 are loaded while marker j is computed (two register sets)
template <int KC>
__device__ __forceinline__ void load_ops(const float2* xs, const float4* cw, int j, int lane, float2& xv, float4 (&c)[(KC + 3) / 4]) {
    constexpr int KCP = (KC + 3) & ~3;
    xv = xs[j * (kRowStride / 2) + (lane ^ (j & 15))];
#pragma unroll
    for (int q = 0; q < KCP / 4; ++q) c[q] = cw[j * (KCP / 4) + q];
}
template <int KC>
__device__ __forceinline__ void fma_ops(float2 (&acc)[2][KC / 2], const float2 xv, const float4 (&c)[(KC + 3) / 4], float (&xx)[2]) {
    constexpr int NP = KC / 2, KCP = (KC + 3) & ~3;
    const float2 x0 = make_float2(xv.x, xv.x), x1 = make_float2(xv.y, xv.y);
#pragma unroll
    for (int q = 0; q < KCP / 4; ++q) {
        acc[0][2 * q] = __ffma2_rn(x0, make_float2(c[q].x, c[q].y), acc[0][2 * q]);
        acc[1][2 * q] = __ffma2_rn(x1, make_float2(c[q].x, c[q].y), acc[1][2 * q]);
        if (2 * q + 1 < NP) {
            acc[0][2 * q + 1] = __ffma2_rn(x0, make_float2(c[q].z, c[q].w), acc[0][2 * q + 1]);
            acc[1][2 * q + 1] = __ffma2_rn(x1, make_float2(c[q].z, c[q].w), acc[1][2 * q + 1]);
        }
    }
    xx[0] = fmaf(xv.x, xv.x, xx[0]);
    xx[1] = fmaf(xv.y, xv.y, xx[1]);
}
template <int KC>
__device__ __forceinline__ void dist_chunk_sp(const float* __restrict__ slot, const float* __restrict__ tab, int d, int lane,
                                              float (&xx)[2], float (&best)[2], float (&second)[2]) {
    constexpr int NP = KC / 2, KCP = (KC + 3) & ~3;
    float2 acc[2][NP];
    const float2* cc = reinterpret_cast<const float2*>(tab);
#pragma unroll
    for (int t = 0; t < NP; ++t) acc[0][t] = acc[1][t] = cc[t];
    const float2* xs = reinterpret_cast<const float2*>(slot);
    const float4* cw = reinterpret_cast<const float4*>(tab + KCP);
    float2 xa, xb;
    float4 ca[(KC + 3) / 4], cb[(KC + 3) / 4];
    load_ops<KC>(xs, cw, 0, lane, xa, ca);
    int j = 0;
    for (; j + 2 <= d; j += 2) {
        load_ops<KC>(xs, cw, j + 1, lane, xb, cb);
        fma_ops<KC>(acc, xa, ca, xx);
        if (j + 2 < d) load_ops<KC>(xs, cw, j + 2, lane, xa, ca);
        fma_ops<KC>(acc, xb, cb, xx);
    }
    if (j < d) fma_ops<KC>(acc, xa, ca, xx);
#pragma unroll
    for (int t = 0; t < NP; ++t) {
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const float a = pack_idx(acc[r][t].x, 2 * t);
            const float b = pack_idx(acc[r][t].y, 2 * t + 1);
            const float lo = fminf(a, b), hi = fmaxf(a, b);
            second[r] = fmin3(fmaxf(best[r], lo), second[r], hi);
            best[r] = fminf(best[r], lo);
        }
    }
}

template <int KC, bool FOLD, int SWZ, bool XX, int UNR = 2>
__device__ __forceinline__ void dist_chunk_var(const float* __restrict__ slot, const float* __restrict__ tab, int d, int lane,
                                               float (&xx)[2], float (&best)[2], float (&second)[2]) {
    constexpr int NP = KC / 2, KCP = (KC + 3) & ~3;
    float2 acc[2][NP];
    const float2* cc = reinterpret_cast<const float2*>(tab);
#pragma unroll
    for (int t = 0; t < NP; ++t) acc[0][t] = acc[1][t] = cc[t];
    const float2* xs = reinterpret_cast<const float2*>(slot);
    const float4* cw = reinterpret_cast<const float4*>(tab + KCP);
#pragma unroll UNR
    for (int j = 0; j < d; ++j) {
        const float2 xv = xs[j * (kRowStride / 2) + (SWZ == 1 ? (lane ^ (j & 15)) : SWZ == 2 ? (lane ^ ((j >> 2) & 15)) : lane)];
        const float2 x0 = make_float2(xv.x, xv.x), x1 = make_float2(xv.y, xv.y);
#pragma unroll
        for (int q = 0; q < KCP / 4; ++q) {
            const float4 c = cw[j * (KCP / 4) + q];
            acc[0][2 * q] = __ffma2_rn(x0, make_float2(c.x, c.y), acc[0][2 * q]);
            acc[1][2 * q] = __ffma2_rn(x1, make_float2(c.x, c.y), acc[1][2 * q]);
            if (2 * q + 1 < NP) {
                acc[0][2 * q + 1] = __ffma2_rn(x0, make_float2(c.z, c.w), acc[0][2 * q + 1]);
                acc[1][2 * q + 1] = __ffma2_rn(x1, make_float2(c.z, c.w), acc[1][2 * q + 1]);
            }
        }
        if (XX) { xx[0] = fmaf(xv.x, xv.x, xx[0]); xx[1] = fmaf(xv.y, xv.y, xx[1]); }
    }
    if (FOLD) {
#pragma unroll
        for (int t = 0; t < NP; ++t)
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const float a = pack_idx(acc[r][t].x, 2 * t), b = pack_idx(acc[r][t].y, 2 * t + 1);
                const float lo = fminf(a, b), hi = fmaxf(a, b);
                second[r] = fmin3(fmaxf(best[r], lo), second[r], hi);
                best[r] = fminf(best[r], lo);
            }
    } else {
#pragma unroll
        for (int t = 0; t < NP; ++t)
#pragma unroll
            for (int r = 0; r < 2; ++r) best[r] += acc[r][t].x + acc[r][t].y;
    }
}

template <int KC, int VAR>
__global__ void __launch_bounds__(512, 1) k_var(float* out, int d, int slots_per_warp) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KCP = (KC + 3) & ~3;
    float* table = sm;
    float* slot = sm + (d + 1) * KCP + warp * d * kRowStride;
    for (int i = threadIdx.x; i < (d + 1) * KCP; i += blockDim.x) table[i] = 0.001f * (i % 97);
    for (int i = lane; i < d * kRowStride; i += 32) slot[i] = 0.01f * ((i * 7 + warp) % 113);
    __syncthreads();
    float acc_out = 0.f;
    for (int s = 0; s < slots_per_warp; ++s) {
        float xx[2] = {0.f, 0.f}, best[2] = {3e38f, 3e38f}, second[2] = {3e38f, 3e38f};
        if (VAR == 0) dist_chunk_var<KC, true, 1, true>(slot, table, d, lane, xx, best, second);
        if (VAR == 1) dist_chunk_var<KC, false, 1, true>(slot, table, d, lane, xx, best, second);
        if (VAR == 2) dist_chunk_var<KC, true, 0, true>(slot, table, d, lane, xx, best, second);
        if (VAR == 3) dist_chunk_var<KC, false, 0, false>(slot, table, d, lane, xx, best, second);
        if (VAR == 4) dist_chunk_var<KC, true, 2, true, 4>(slot, table, d, lane, xx, best, second);
        if (VAR == 5) dist_chunk_var<KC, true, 2, true, 2>(slot, table, d, lane, xx, best, second);
        if (VAR == 6) dist_chunk_var<KC, true, 1, true, 4>(slot, table, d, lane, xx, best, second);
        if (VAR == 7) dist_chunk_var<KC, true, 1, true, 1>(slot, table, d, lane, xx, best, second);
        if (VAR == 8) dist_chunk_var<KC, true, 2, false, 4>(slot, table, d, lane, xx, best, second);
        acc_out += best[0] + best[1] + second[0] + second[1] + xx[0] + xx[1];
        table[0] = acc_out * 1e-30f;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc_out;
}
template <int KC, int VAR>
void runv(int d, int warps, int sms, float* out, const char* name) {
    constexpr int KCP = (KC + 3) & ~3;
    const size_t shm = sizeof(float) * ((size_t)(d + 1) * KCP + (size_t)warps * d * kRowStride);
    cudaFuncSetAttribute(k_var<KC, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    const int slots = 200;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_var<KC, VAR><<<sms, warps * 32, shm>>>(out, d, slots);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        k_var<KC, VAR><<<sms, warps * 32, shm>>>(out, d, slots);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double fma = (double)sms * warps * slots * 64.0 * KC * d;
    printf("variant %-28s KC=%2d d=%2d warps=%2d: %.2f TFMA/s  err=%s\n", name, KC, d, warps, fma / best * 1e-9, cudaGetErrorString(cudaGetLastError()));
}

template <int KC, bool FOLD_ONLY>
__global__ void __launch_bounds__(512, 1) k_chunk(float* out, int d, int slots_per_warp) {
    extern __shared__ __align__(128) float sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int KCP = (KC + 3) & ~3;
    float* table = sm;                                  // [cc | w[d][KCP]]
    float* slot = sm + (d + 1) * KCP + warp * d * kRowStride;
    for (int i = threadIdx.x; i < (d + 1) * KCP; i += blockDim.x) table[i] = 0.001f * (i % 97);
    for (int i = lane; i < d * kRowStride; i += 32) slot[i] = 0.01f * ((i * 7 + warp) % 113);
    __syncthreads();
    float acc_out = 0.f;
    for (int s = 0; s < slots_per_warp; ++s) {
        float xx[2] = {0.f, 0.f}, best[2] = {3e38f, 3e38f}, second[2] = {3e38f, 3e38f};
        if (FOLD_ONLY) dist_chunk_sp<KC>(slot, table, d, lane, xx, best, second);
        else dist_chunk<KC>(slot, table, d, lane, best, second);   // the product's micro-kernel
        acc_out += best[0] + best[1] + second[0] + second[1] + xx[0] + xx[1];
        table[0] = acc_out * 1e-30f;   // keep iterations dependent
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc_out;
}

template <int KC, bool SP = false>
void run(int d, int warps, int sms, float* out) {
    constexpr int KCP = (KC + 3) & ~3;
    const size_t shm = sizeof(float) * ((size_t)(d + 1) * KCP + (size_t)warps * d * kRowStride);
    cudaFuncSetAttribute(k_chunk<KC, SP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shm);
    const int slots = 200;
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    k_chunk<KC, SP><<<sms, warps * 32, shm>>>(out, d, slots);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(a);
        k_chunk<KC, SP><<<sms, warps * 32, shm>>>(out, d, slots);
        cudaEventRecord(b);
        cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    const double fma = (double)sms * warps * slots * 64.0 * KC * d;
    const double cells = (double)sms * warps * slots * 64.0;
    printf("%s KC=%2d d=%2d warps=%2d: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM @1.965)  -> %.3f ms per 10M cells   err=%s\n", SP ? "pipelined" : "baseline ", KC, d, warps, best,
           fma / best * 1e-9, fma / best * 1e-6 / sms / 1.965, 1e7 / cells * best, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    float* out;
    cudaMalloc(&out, sizeof(float) * sms * 512);
    for (int warps : {4, 8, 12, 16}) run<30>(40, warps, sms, out);
    runv<30, 0>(40, 16, sms, out, "fold+swizzle+xx (no peel)");
    runv<30, 1>(40, 16, sms, out, "no fold");
    runv<30, 2>(40, 16, sms, out, "no swizzle");
    runv<30, 3>(40, 16, sms, out, "loop only");
    runv<30, 4>(40, 16, sms, out, "fold+swz(j>>2)+xx unroll4");
    runv<30, 5>(40, 16, sms, out, "fold+swz(j>>2)+xx unroll2");
    runv<30, 6>(40, 16, sms, out, "fold+swz(j)+xx unroll4");
    runv<30, 7>(40, 16, sms, out, "fold+swz(j)+xx unroll1");
    runv<30, 8>(40, 16, sms, out, "fold+swz(j>>2) unroll4 no xx");
    runv<22, 4>(40, 16, sms, out, "fold+swz(j>>2)+xx unroll4");
    runv<22, 0>(40, 16, sms, out, "fold+swizzle+xx (no peel)");
    runv<20, 4>(15, 16, sms, out, "fold+swz(j>>2)+xx unroll4");
    runv<20, 0>(15, 16, sms, out, "fold+swizzle+xx (no peel)");
    runv<8, 4>(40, 16, sms, out, "fold+swz(j>>2)+xx unroll4");
    runv<8, 0>(40, 16, sms, out, "fold+swizzle+xx (no peel)");
    runv<32, 3>(40, 16, sms, out, "loop only");
    runv<30, 3>(40, 8, sms, out, "loop only");
    run<32>(40, 16, sms, out);
    run<16>(40, 16, sms, out);
    run<8>(40, 16, sms, out);
    run<20>(15, 16, sms, out);
    run<22>(40, 16, sms, out);
    run<30, true>(40, 16, sms, out);
    run<24, true>(40, 16, sms, out);
    run<22, true>(40, 16, sms, out);
    run<16, true>(40, 16, sms, out);
    run<14, true>(40, 16, sms, out);
    run<16, true>(40, 8, sms, out);
    run<8, true>(40, 16, sms, out);
    run<20, true>(15, 16, sms, out);
    run<10, true>(15, 16, sms, out);
    return 0;
}
