// ldtm_probe.cu — tensor-memory read throughput of tcgen05.ld on B200 (sm_100a), the resource that bounds the scoring pass
// (depecher_b200/csrc/score_pass.cu): W warps of one CTA per SM read their lane quarters of TMEM in a loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ldtm_probe ldtm_probe.cu && ./ldtm_probe
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int NCOL>
__device__ __forceinline__ uint32_t ld_cols(uint32_t taddr);
template <>
__device__ __forceinline__ uint32_t ld_cols<16>(uint32_t taddr) {
    uint32_t v[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= v[i];
    return x;
}
template <>
__device__ __forceinline__ uint32_t ld_cols<32>(uint32_t taddr) {
    uint32_t v[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
                   "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 32; ++i) x ^= v[i];
    return x;
}
// two x16 loads in flight before the wait
template <>
__device__ __forceinline__ uint32_t ld_cols<1616>(uint32_t taddr) {
    uint32_t v[16], w[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]),
                   "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]), "=r"(w[8]), "=r"(w[9]), "=r"(w[10]),
                   "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                 : "r"(taddr + 16) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    uint32_t x = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) x ^= v[i] ^ w[i];
    return x;
}

template <int NCOL>
__global__ void __launch_bounds__(512, 1) probe(int warps, int reps, long long* cycles, uint32_t* sink) {
    __shared__ uint32_t tmem_word;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&tmem_word)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_word;
    uint32_t acc = 0;
    long long t0 = 0, t1 = 0;
    if (warp < warps) {
        const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int cols = NCOL == 1616 ? 32 : NCOL;
        __syncwarp();
        t0 = clock64();
        for (int r = 0; r < reps; ++r) acc ^= ld_cols<NCOL>(base + (uint32_t)(((r * cols) + (warp >> 2) * 64) & 255));
        t1 = clock64();
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp < warps && (threadIdx.x & 31) == 0) cycles[blockIdx.x * 16 + warp] = t1 - t0;
    if (acc == 0x12345678u) sink[0] = acc;
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(512u));
}

template <int NCOL>
void run(const char* name, int warps) {
    long long* d_c;
    uint32_t* d_s;
    cudaMalloc(&d_c, sizeof(long long) * 148 * 16);
    cudaMalloc(&d_s, 4);
    const int reps = 20000;
    probe<NCOL><<<148, 512>>>(warps, 100, d_c, d_s);
    probe<NCOL><<<148, 512>>>(warps, reps, d_c, d_s);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
    long long h[148 * 16];
    cudaMemcpy(h, d_c, sizeof(h), cudaMemcpyDeviceToHost);
    long long mx = 0;
    for (int w = 0; w < warps; ++w) mx = h[w] > mx ? h[w] : mx;
    const int cols = NCOL == 1616 ? 32 : NCOL;
    const double bytes = (double)warps * reps * cols * 32 * 4;
    printf("%-22s %2d warps: %7.1f cycles per load per warp, %6.1f B/cycle/SM\n", name, warps, (double)mx / reps, bytes / mx);
    cudaFree(d_c);
    cudaFree(d_s);
}

int main() {
    for (int w : {1, 4, 8, 12, 16}) run<16>("32x32b.x16", w);
    for (int w : {1, 4, 8, 12, 16}) run<32>("32x32b.x32", w);
    for (int w : {4, 12, 16}) run<1616>("2 x 32x32b.x16 + wait", w);
    return 0;
}
