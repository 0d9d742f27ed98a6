// api.cu — the C ABI (include/depecher_b200.h): context, resident datasets and the entry points that
// replace the reference's four Rcpp exports (/root/reference/src/InterfaceUtils.cpp:66-161).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "api_internal.cuh"
#include "aux_kernels.cuh"
#include "comm.cuh"
#include "common.cuh"
#include "scale.cuh"
#include "session.cuh"
#if defined(__x86_64__)
#include <immintrin.h>
#endif
#include "../../include/dpr_philox.h"

namespace dpr {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }
int fail(int code, const std::string& msg) {
    g_error = msg;
    return code;
}
Context& ctx() {
    static Context c;
    return c;
}
int ensure_ready() {
    Context& c = ctx();
    if (c.ready) return DPR_OK;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n < 1)
        return fail(DPR_ERR_NO_DEVICE, std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0") +
                                           " (this library has no CPU fallback)");
    if (c.device >= n) return fail(DPR_ERR_NO_DEVICE, "device ordinal out of range");
    DPR_CUDA(cudaSetDevice(c.device));
    cudaDeviceProp prop;
    DPR_CUDA(cudaGetDeviceProperties(&prop, c.device));
    if (prop.major != 10)
        return fail(DPR_ERR_NO_DEVICE, std::string("device '") + prop.name + "' is sm_" + std::to_string(prop.major) + std::to_string(prop.minor) +
                                           "; the kernels are built for sm_100a (B200) only");
    c.sms = prop.multiProcessorCount;
    DPR_CUDA(cudaStreamCreateWithFlags(&c.own_stream, cudaStreamNonBlocking));
    if (!c.stream) c.stream = c.own_stream;
    {   // keep freed device memory in the stream-ordered pool: a penalty sweep allocates and frees multi-GB label and
        // partial-sum buffers per call, and going back to the driver for them costs more than the kernels
        cudaMemPool_t pool;
        DPR_CUDA(cudaDeviceGetDefaultMemPool(&pool, c.device));
        unsigned long long keep = ~0ULL;
        DPR_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
    }
    c.ready = true;
    return DPR_OK;
}

// stream-ordered device memory (see ensure_ready): allocation and release are ordered on the library's stream
int dev_alloc(void** p, size_t bytes) {
    DPR_CUDA(cudaMallocAsync(p, bytes ? bytes : 1, ctx().stream));
    return DPR_OK;
}
void dev_free(void* p) {
    if (p) cudaFreeAsync(p, ctx().stream);
}

void bench_session_forget(dpr_dataset* ds);

// ---- small RAII helpers for temporaries ----------------------------------------------------------
struct DevBuf {
    void* p = nullptr;
    ~DevBuf() { dev_free(p); }
    int alloc(size_t bytes, bool zero = false) {
        DPR_TRY(dev_alloc(&p, bytes));
        if (zero) DPR_CUDA(cudaMemsetAsync(p, 0, bytes ? bytes : 1, ctx().stream));
        return DPR_OK;
    }
    template <typename T> T* as() { return reinterpret_cast<T*>(p); }
};

static void colmajor_to_rowmajor(const double* cm, int rows, int cols, std::vector<double>& rm) {
    rm.resize((size_t)rows * cols);
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) rm[(size_t)i * cols + j] = cm[(size_t)j * rows + i];
}
static void rowmajor_to_colmajor(const std::vector<double>& rm, int rows, int cols, double* cm) {
    for (int i = 0; i < rows; ++i)
        for (int j = 0; j < cols; ++j) cm[(size_t)j * rows + i] = rm[(size_t)i * cols + j];
}

static int dataset_alloc(int64_t n, int32_t d, dpr_dataset** out) {
    DPR_TRY(ensure_ready());
    if (n < 1 || d < 1 || !out) return fail(DPR_ERR_INVALID, "dataset needs n >= 1, d >= 1");
    std::unique_ptr<dpr_dataset> ds(new dpr_dataset());
    ds->n = n;
    ds->d = d;
    ds->n_pad = (n + kSlotCells - 1) / kSlotCells * kSlotCells;
    const size_t bytes = sizeof(float) * (size_t)(ds->n_pad / kSlotCells) * tile_floats(d);
    DPR_TRY(dev_alloc((void**)&ds->x, bytes));
    DPR_CUDA(cudaMemsetAsync(ds->x, 0, bytes, ctx().stream));
    DPR_TRY(dev_alloc((void**)&ds->tile_xx, sizeof(float) * (size_t)(ds->n_pad / kSlotCells + 1)));
    *out = ds.release();
    return DPR_OK;
}
// after the cells are in place: the per-tile norm bounds the fused pass reads
static int dataset_finish(dpr_dataset* ds) {
    DPR_TRY(launch_tile_norms(ds->x, ds->n_pad / kSlotCells, ds->d, ds->tile_xx, ctx().stream));
    return launch_max_finite(ds->tile_xx, ds->n_pad / kSlotCells, ds->tile_xx + ds->n_pad / kSlotCells, ctx().stream);
}
static int dataset_labels(dpr_dataset* ds) {
    if (!ds->labels) DPR_TRY(dev_alloc((void**)&ds->labels, sizeof(int32_t) * (size_t)ds->n_pad));
    return DPR_OK;
}

// host -> device.  R hands over pageable double memory; copying that directly is staged by the driver at a few
// GB/s and moves 8 bytes per value.  Instead host threads narrow each chunk to fp32 into pinned buffers (kept for the
// life of the process), the fp32 chunk crosses PCIe asynchronously, and a kernel re-tiles it into the resident layout
// while the threads already narrow the next chunk.
constexpr int kStageBufs = 4;   // a ring: the host threads fill buffer c % 4 while the copies of the chunks before it are in flight
struct UploadStage {
    float* pinned[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
    float* dev[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t done[kStageBufs] = {nullptr, nullptr, nullptr, nullptr};
    int64_t cap = 0;
};
static UploadStage g_stage;
static int upload_stage(int64_t elems) {
    if (g_stage.cap >= elems) return DPR_OK;
    for (int b = 0; b < kStageBufs; ++b) {
        if (g_stage.pinned[b]) cudaFreeHost(g_stage.pinned[b]);
        if (g_stage.dev[b]) cudaFree(g_stage.dev[b]);
        DPR_CUDA(cudaMallocHost(&g_stage.pinned[b], sizeof(float) * elems));
        DPR_CUDA(cudaMalloc(&g_stage.dev[b], sizeof(float) * elems));
        if (!g_stage.done[b]) DPR_CUDA(cudaEventCreateWithFlags(&g_stage.done[b], cudaEventDisableTiming));
    }
    g_stage.cap = elems;
    return DPR_OK;
}

// double -> float of one contiguous piece.  With AVX: four values per convert and streaming (non-temporal) stores, so the
// pinned destination is written without first being read into the cache (it is only ever read by the DMA engine).
#if defined(__x86_64__)
__attribute__((target("avx"))) static void narrow_piece_avx(const double* src, float* dst, int64_t a, int64_t b) {
    int64_t i = a;
    for (; i < b && (reinterpret_cast<uintptr_t>(dst + i) & 31); ++i) dst[i] = (float)src[i];
    for (; i + 8 <= b; i += 8) {
        const __m128 lo = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i));
        const __m128 hi = _mm256_cvtpd_ps(_mm256_loadu_pd(src + i + 4));
        _mm256_stream_ps(dst + i, _mm256_set_m128(hi, lo));
    }
    for (; i < b; ++i) dst[i] = (float)src[i];
    _mm_sfence();
}
#endif
#if defined(__x86_64__)
__attribute__((target("avx"))) static void copy_piece_avx(const double* src, double* dst, int64_t count) {
    int64_t i = 0;
    for (; i < count && (reinterpret_cast<uintptr_t>(dst + i) & 31); ++i) dst[i] = src[i];
    for (; i + 8 <= count; i += 8) {
        _mm256_stream_pd(dst + i, _mm256_loadu_pd(src + i));
        _mm256_stream_pd(dst + i + 4, _mm256_loadu_pd(src + i + 4));
    }
    for (; i < count; ++i) dst[i] = src[i];
    _mm_sfence();
}
#endif
// pageable -> pinned copy of doubles with streaming stores (the destination is only ever read by the DMA engine)
static void copy_piece(const double* src, double* dst, int64_t count) {
#if defined(__x86_64__)
    static const bool avx = __builtin_cpu_supports("avx");
    if (avx) return copy_piece_avx(src, dst, count);
#endif
    std::memcpy(dst, src, sizeof(double) * (size_t)count);
}
// double -> float, reporting whether every value survived the trip unchanged (a NaN counts as changed: the caller then ships doubles)
#if defined(__x86_64__)
__attribute__((target("avx"))) static bool narrow_check_avx(const double* src, float* dst, int64_t a, int64_t b) {
    int64_t i = a;
    bool ok = true;
    for (; i < b && (reinterpret_cast<uintptr_t>(dst + i) & 31); ++i) {
        dst[i] = (float)src[i];
        ok &= (double)dst[i] == src[i];
    }
    __m256d all = _mm256_castsi256_pd(_mm256_set1_epi64x(-1));
    for (; i + 8 <= b; i += 8) {
        const __m256d x0 = _mm256_loadu_pd(src + i), x1 = _mm256_loadu_pd(src + i + 4);
        const __m128 lo = _mm256_cvtpd_ps(x0), hi = _mm256_cvtpd_ps(x1);
        all = _mm256_and_pd(all, _mm256_and_pd(_mm256_cmp_pd(_mm256_cvtps_pd(lo), x0, _CMP_EQ_OQ), _mm256_cmp_pd(_mm256_cvtps_pd(hi), x1, _CMP_EQ_OQ)));
        _mm256_stream_ps(dst + i, _mm256_set_m128(hi, lo));
    }
    ok &= _mm256_movemask_pd(all) == 0xf;
    for (; i < b; ++i) {
        dst[i] = (float)src[i];
        ok &= (double)dst[i] == src[i];
    }
    _mm_sfence();
    return ok;
}
#endif
static bool narrow_check(const double* src, float* dst, int64_t a, int64_t b) {
#if defined(__x86_64__)
    static const bool avx = __builtin_cpu_supports("avx");
    if (avx) return narrow_check_avx(src, dst, a, b);
#endif
    bool ok = true;
    for (int64_t i = a; i < b; ++i) {
        dst[i] = (float)src[i];
        ok &= (double)dst[i] == src[i];
    }
    return ok;
}
template <typename T>
static void narrow_piece(const T* src, float* dst, int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) dst[i] = (float)src[i];
}
template <>
void narrow_piece<double>(const double* src, float* dst, int64_t a, int64_t b) {
#if defined(__x86_64__)
    static const bool avx = __builtin_cpu_supports("avx");
    if (avx) return narrow_piece_avx(src, dst, a, b);
#endif
    for (int64_t i = a; i < b; ++i) dst[i] = (float)src[i];
}

// host threads for the narrowing / staging copies: up to 16, and the cores are shared between the ranks of a node
static int host_copy_threads() {
    static const int forced = getenv("DPR_COPY_THREADS") ? atoi(getenv("DPR_COPY_THREADS")) : 0;   // tuning aid
    if (forced > 0) return forced;
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned per_rank = std::max(2u, hw / (unsigned)std::max(1, ctx().world));
    const unsigned t = std::min(16u, std::min(hw, per_rank));
    return (int)(t >= 8 ? t - 2 : t);   // (two cores left to the thread that ships the chunks and to the driver: 14 of 16 measured best, 36.7 ms against 38.4 for a 3.2 GB matrix)
}

// The host side of an upload as a pipeline: `total` elements in chunks of `chunk`, chunk c staged in ring buffer c % kStageBufs.  Worker
// threads (started once per call) draw pieces of chunks from one counter — fill(e0, cnt, a, b, buf) writes elements [a, b) of the chunk that
// starts at element e0 into pinned buffer `buf`, and returns false to stop everything (the caller then takes another route) — and the calling
// thread ships every chunk whose pieces are all there (ship(buf, e0, cnt): the asynchronous copy and whatever kernel follows it), records the
// buffer's event and hands buffers back to the workers as their copies complete.  No thread is created or joined per chunk, no worker waits
// for the slowest of its chunk, and the copy of the last chunk is all that is left when the workers are done.
template <class Fill, class Ship>
static int staged_pipeline(int64_t total, int64_t chunk, int threads, Fill fill, Ship ship, bool* stopped = nullptr) {
    if (stopped) *stopped = false;
    if (total <= 0) return DPR_OK;
    const int64_t nchunks = (total + chunk - 1) / chunk;
    const int P = chunk < (1 << 16) ? 1 : 2 * std::max(1, threads);   // pieces per chunk
    for (int b = 0; b < kStageBufs; ++b) DPR_CUDA(cudaEventSynchronize(g_stage.done[b]));   // (whatever used the ring before has left it)
    std::unique_ptr<std::atomic<int>[]> pieces_done(new std::atomic<int>[(size_t)nchunks]);
    for (int64_t c = 0; c < nchunks; ++c) pieces_done[c].store(0, std::memory_order_relaxed);
    std::atomic<int64_t> next{0}, free_upto{kStageBufs};
    std::atomic<int> stop{0};
    auto worker = [&] {
        for (;;) {
            const int64_t t = next.fetch_add(1, std::memory_order_relaxed);
            if (t >= nchunks * P || stop.load(std::memory_order_relaxed)) return;
            const int64_t c = t / P;
            const int piece = (int)(t % P);
            while (free_upto.load(std::memory_order_acquire) <= c) {
                if (stop.load(std::memory_order_relaxed)) return;
                std::this_thread::yield();
            }
            const int64_t e0 = c * chunk, cnt = std::min(chunk, total - e0);
            if (!fill(e0, cnt, cnt * piece / P, cnt * (piece + 1) / P, (int)(c % kStageBufs))) stop.store(1);
            pieces_done[c].fetch_add(1, std::memory_order_release);
        }
    };
    std::vector<std::thread> pool;
    const int n_workers = nchunks * P == 1 ? 0 : std::max(1, threads);
    for (int t = 0; t < n_workers; ++t) pool.emplace_back(worker);
    if (n_workers == 0) worker();
    int rc = DPR_OK;
    int64_t issued = 0, completed = 0;
    while (issued < nchunks && rc == DPR_OK && !stop.load(std::memory_order_relaxed)) {
        while (completed < issued && cudaEventQuery(g_stage.done[completed % kStageBufs]) == cudaSuccess) {
            ++completed;
            free_upto.store(completed + kStageBufs, std::memory_order_release);
        }
        if (pieces_done[issued].load(std::memory_order_acquire) == P) {
            const int b = (int)(issued % kStageBufs);
            const int64_t e0 = issued * chunk;
            rc = ship(b, e0, std::min(chunk, total - e0));
            if (rc == DPR_OK && cudaEventRecord(g_stage.done[b], ctx().stream) != cudaSuccess) rc = fail(DPR_ERR_CUDA, "upload: cudaEventRecord failed");
            ++issued;
            continue;
        }
        std::this_thread::yield();
    }
    if (rc != DPR_OK) stop.store(1);
    for (auto& th : pool) th.join();
    if (stopped) *stopped = stop.load() != 0 && rc == DPR_OK;
    return rc;
}
constexpr int64_t kUploadChunk = (int64_t)8 << 20;   // elements per ring buffer (32 MB of fp32)

template <typename T>
static int dataset_upload(dpr_dataset* ds, const T* X) {
    const int64_t total = ds->n * ds->d;
    const int64_t chunk = std::min<int64_t>(total, kUploadChunk);
    DPR_TRY(upload_stage(chunk));
    DPR_TRY(staged_pipeline(
        total, chunk, host_copy_threads(),
        [&](int64_t e0, int64_t, int64_t a, int64_t b, int buf) {
            narrow_piece<T>(X + e0, g_stage.pinned[buf], a, b);
            return true;
        },
        [&](int buf, int64_t e0, int64_t cnt) -> int {
            DPR_CUDA(cudaMemcpyAsync(g_stage.dev[buf], g_stage.pinned[buf], sizeof(float) * cnt, cudaMemcpyHostToDevice, ctx().stream));
            return launch_f32_to_cols(g_stage.dev[buf], e0, cnt, ds->n, ds->d, ds->x, ctx().stream);
        }));
    DPR_TRY(dataset_finish(ds));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}

// one Session over `ds` with a single problem
// The cell-major copy that the delta iterations of a Lloyd fit gather moved cells from (delta_sums_kernel): built once per
// dataset, on its first fit, when the matrix is large enough for the gathers to miss the L2 (the tile-major layout spreads
// a row over tile_groups(d) sectors) and the device has the memory to spare.  A plan decision, not a requirement: without
// it the gathers read the tile-major matrix.
static int dataset_rows(dpr_dataset* ds) {
    static const bool off = getenv("DPR_NO_ROWS_COPY") != nullptr;   // A/B aid
    if (ds->rows || ds->rows_tried || off) return DPR_OK;
    ds->rows_tried = true;
    const size_t bytes = sizeof(float) * (size_t)ds->n_pad * 4 * tile_groups(ds->d);
    if (bytes < ((size_t)64 << 20)) return DPR_OK;
    size_t free_b = 0, total_b = 0;
    if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < bytes + ((size_t)4 << 30)) {
        cudaGetLastError();
        return DPR_OK;
    }
    if (dev_alloc((void**)&ds->rows, bytes) != DPR_OK) {
        ds->rows = nullptr;
        cudaGetLastError();
        return DPR_OK;
    }
    return launch_rows_from_tiles(ds->x, ds->n_pad / kSlotCells, ds->d, ds->rows, ctx().stream);
}

static int single_session(dpr_dataset* ds, int k, double reg, int no_zero, bool want_sums, Session& s) {
    DPR_TRY(dataset_labels(ds));
    if (want_sums) DPR_TRY(dataset_rows(ds));
    std::vector<ProblemSpec> specs(1);
    specs[0] = ProblemSpec{ds->x, ds->tile_xx, ds->tile_xx + ds->n_pad / kSlotCells, ds->n, k, reg, no_zero, ds->labels, ds->rows};
    return s.init(specs, ds->d, want_sums);
}

// `base` is added to every label on the way out (R's clusters are 1-based: dAllocate.R:127-134 — adding it here saves the caller a
// pass over a vector with one entry per cell)
static int copy_labels_out(const int32_t* d_labels, int64_t n, int32_t* host, int32_t base = 0, const int32_t* lut = nullptr) {   // lut (host, nullable): label l goes out as lut[l]
    if (!host) return DPR_OK;
    if (n >= (1 << 19) && g_stage.cap == 0) DPR_TRY(upload_stage((int64_t)4 << 20));   // (a process that has not uploaded anything yet: the pinned buffers are made here)
    if (n >= (1 << 19) && g_stage.cap > 0) {
        // The caller's vector is pageable (and usually untouched: every page of it faults on its first write), the driver would stage
        // the copy at a few GB/s.  The upload pipeline in reverse: the labels arrive by DMA in the ring of pinned buffers, 4 MB at a
        // time; worker threads started once per call draw pieces of landed chunks from one counter and write them out (with the offset
        // or the lookup table applied), the calling thread issues a copy whenever a buffer has been emptied.
        const int64_t chunk = std::min<int64_t>(g_stage.cap, (int64_t)1 << 20);
        const int64_t nchunks = (n + chunk - 1) / chunk;
        const int threads = host_copy_threads();
        const int P = std::max(1, threads);
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));   // (the labels are final; whatever used the ring before has left it)
        std::unique_ptr<std::atomic<int>[]> pieces_done(new std::atomic<int>[(size_t)nchunks]);
        for (int64_t c = 0; c < nchunks; ++c) pieces_done[c].store(0, std::memory_order_relaxed);
        std::atomic<int64_t> next{0}, landed{0};
        std::atomic<int> stop{0};
        auto worker = [&] {
            for (;;) {
                const int64_t t = next.fetch_add(1, std::memory_order_relaxed);
                if (t >= nchunks * P) return;
                const int64_t c = t / P;
                const int piece = (int)(t % P);
                while (landed.load(std::memory_order_acquire) <= c) {
                    if (stop.load(std::memory_order_relaxed)) return;
                    std::this_thread::yield();
                }
                const int64_t e0 = c * chunk, cnt = std::min(chunk, n - e0);
                const int64_t a = cnt * piece / P, z = cnt * (piece + 1) / P;
                const int32_t* src = reinterpret_cast<const int32_t*>(g_stage.pinned[c % kStageBufs]);
                if (lut)
                    for (int64_t e = a; e < z; ++e) host[e0 + e] = lut[src[e]];
                else if (base == 0) std::memcpy(host + e0 + a, src + a, sizeof(int32_t) * (size_t)(z - a));
                else
                    for (int64_t e = a; e < z; ++e) host[e0 + e] = src[e] + base;
                pieces_done[c].fetch_add(1, std::memory_order_release);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 0; t < threads; ++t) pool.emplace_back(worker);
        int rc = DPR_OK;
        int64_t issued = 0, arrived = 0, freed = 0;
        while (arrived < nchunks && rc == DPR_OK) {
            bool progress = false;
            while (issued < nchunks && issued < freed + kStageBufs && rc == DPR_OK) {
                const int bb = (int)(issued % kStageBufs);
                const int64_t e0 = issued * chunk, cnt = std::min(chunk, n - e0);
                if (cudaMemcpyAsync(g_stage.pinned[bb], d_labels + e0, sizeof(int32_t) * cnt, cudaMemcpyDeviceToHost, ctx().stream) != cudaSuccess ||
                    cudaEventRecord(g_stage.done[bb], ctx().stream) != cudaSuccess)
                    rc = fail(DPR_ERR_CUDA, "labels to the host: copy failed");
                ++issued;
                progress = true;
            }
            while (arrived < issued && cudaEventQuery(g_stage.done[arrived % kStageBufs]) == cudaSuccess) {
                ++arrived;
                landed.store(arrived, std::memory_order_release);
                progress = true;
            }
            while (freed < arrived && pieces_done[freed].load(std::memory_order_acquire) == P) {
                ++freed;
                progress = true;
            }
            if (!progress) std::this_thread::yield();
        }
        if (rc != DPR_OK) stop.store(1);
        for (auto& th : pool) th.join();
        return rc;
    }
    DPR_CUDA(cudaMemcpyAsync(host, d_labels, sizeof(int32_t) * n, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    if (lut)
        for (int64_t e = 0; e < n; ++e) host[e] = lut[host[e]];
    else if (base != 0)
        for (int64_t e = 0; e < n; ++e) host[e] += base;
    return DPR_OK;
}

// cluster_norm (Clusterer.cpp:198-211) for problem p of a session
static int cluster_norm_dev(const Problem& pr, int d, double reg, double* out, bool row_sharded = false) {
    const int blocks = (int)std::min<int64_t>(148 * 4, (pr.n + 255) / 256);
    DevBuf part;
    DPR_TRY(part.alloc(sizeof(double) * blocks));
    DPR_TRY(launch_ssq(pr.x, pr.n, d, pr.mu, pr.k, pr.labels, part.as<double>(), blocks, ctx().stream));
    std::vector<double> hp(blocks), mu((size_t)pr.k * d);
    DPR_CUDA(cudaMemcpyAsync(hp.data(), part.p, sizeof(double) * blocks, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(mu.data(), pr.mu, sizeof(double) * mu.size(), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    double ssq = 0.0;
    for (double v : hp) ssq += v;
    if (row_sharded && ctx().world > 1) {            // rows sharded over ranks: the squared distances of all ranks' cells
        DPR_CUDA(cudaMemcpyAsync(part.p, &ssq, sizeof(double), cudaMemcpyHostToDevice, ctx().stream));
        DPR_TRY(comm_allreduce_sum_f64(part.as<double>(), 1));
        DPR_CUDA(cudaMemcpyAsync(&ssq, part.p, sizeof(double), cudaMemcpyDeviceToHost, ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    for (int j = 0; j < d; ++j)                      // column by column, like the reference (:204-208)
        for (int c = 0; c < pr.k; ++c) ssq += mu[(size_t)c * d + j] * reg;
    *out = ssq;
    return DPR_OK;
}

// k-means++ for a list of fits (data pointers may differ); mus[f] = device k x d row-major
struct SeedJob {
    const float* x;
    int64_t n;
    uint32_t fit_id;
    double* mu;
    long long* rows;
};
static int run_seeding(const std::vector<SeedJob>& jobs, int d, int k, uint64_t seed) {
    const int F = (int)jobs.size();
    int64_t n_max = 0, w_total = 0, b_total = 0;
    for (auto& j : jobs) {
        n_max = std::max(n_max, j.n);
        w_total += j.n;
        b_total += (j.n + kSeedBlock - 1) / kSeedBlock;
    }
    DevBuf w, bs, fits, done;
    DPR_TRY(w.alloc(sizeof(double) * w_total));
    DPR_TRY(bs.alloc(sizeof(double) * b_total));
    DPR_TRY(fits.alloc(sizeof(SeedFit) * F));
    DPR_TRY(done.alloc(sizeof(int) * F, true));
    std::vector<SeedFit> h(F);
    int64_t wo = 0, bo = 0;
    for (int f = 0; f < F; ++f) {
        h[f] = SeedFit{jobs[f].x, jobs[f].n, jobs[f].fit_id, 0, jobs[f].mu, w.as<double>() + wo, bs.as<double>() + bo, jobs[f].rows, done.as<int>() + f};
        wo += jobs[f].n;
        bo += (jobs[f].n + kSeedBlock - 1) / kSeedBlock;
    }
    DPR_CUDA(cudaMemcpyAsync(fits.p, h.data(), sizeof(SeedFit) * F, cudaMemcpyHostToDevice, ctx().stream));
    static const bool timing = getenv("DPR_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t0;
    if (timing) {
        cudaStreamSynchronize(ctx().stream);
        t0 = std::chrono::steady_clock::now();
    }
    DPR_TRY(launch_seeding(fits.as<SeedFit>(), F, n_max, d, k, seed, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    if (timing)
        fprintf(stderr, "[dpr] seeding: %d rounds of %d fits %.2f ms\n", k - 1, F, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return DPR_OK;
}

// ARI for pairs of device label arrays
struct PairJob {
    const int32_t* a;
    const int32_t* b;
    int64_t n;
    uint32_t pair_id;
};
static int run_ari(const std::vector<PairJob>& jobs, uint32_t k, uint64_t seed, double* out_host) {
    const int P = (int)jobs.size();
    const int kk = (int)k + 1;
    int64_t n_max = 0;
    std::vector<AriPair> h(P);
    for (int p = 0; p < P; ++p) {
        h[p] = AriPair{jobs[p].a, jobs[p].b, jobs[p].n, jobs[p].pair_id, 0};
        n_max = std::max(n_max, jobs[p].n);
    }
    DevBuf pairs, tables, stab, bad, out;
    DPR_TRY(pairs.alloc(sizeof(AriPair) * P));
    DPR_TRY(tables.alloc(sizeof(unsigned long long) * (size_t)P * kk * kk, true));
    DPR_TRY(stab.alloc(sizeof(unsigned int) * P, true));
    DPR_TRY(bad.alloc(sizeof(int32_t), true));
    DPR_TRY(out.alloc(sizeof(double) * P));
    DPR_CUDA(cudaMemcpyAsync(pairs.p, h.data(), sizeof(AriPair) * P, cudaMemcpyHostToDevice, ctx().stream));
    DPR_TRY(launch_ari(pairs.as<AriPair>(), P, n_max, kk, ctx().ari_mode, seed, tables.as<unsigned long long>(), stab.as<unsigned int>(),
                       bad.as<int32_t>(), out.as<double>(), ctx().stream));
    int32_t hbad = 0;
    DPR_CUDA(cudaMemcpyAsync(out_host, out.p, sizeof(double) * P, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    if (hbad) return fail(DPR_ERR_INVALID, "rand_index: a label lies outside [0, k]");
    return DPR_OK;
}

// ------------------------------------------------------------------------------------------------
// find_centers on a resident dataset (Clusterer.cpp:213-265), from k-means++ or from given centres
// ------------------------------------------------------------------------------------------------
static int find_centers_dev(dpr_dataset* ds, const double* init_mu_colmajor, uint32_t k, double reg, int no_zero, int max_iter,
                            uint64_t seed, uint32_t fit_id, int32_t* idx_out, double* centers_out, double* norm_out,
                            int32_t* n_iter_out, Session* ready = nullptr) {   // ready: a session already set up over `ds` (dpr_sparse_k_means on a host matrix)
    if (!ds) return fail(DPR_ERR_INVALID, "null dataset");
    if (k < 2) return fail(DPR_ERR_INVALID, "k must be >= 2 (initialize_mu writes two rows, Clusterer.cpp:157-162)");
    if (max_iter < 1) return fail(DPR_ERR_INVALID, "max_iter must be >= 1");
    static const bool timing = getenv("DPR_TIMING") != nullptr;   // diagnostics: host wall-clock per phase (each closed by a stream synchronize)
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(ctx().stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[dpr] find_centers: %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    // The caller's label vector is usually untouched memory: every page of it faults on its first write.  A helper thread writes it once
    // while the device seeds and iterates (the host has nothing else to do), so that the copy-out at the end finds the pages mapped.
    struct Joiner {
        std::thread t;
        ~Joiner() { if (t.joinable()) t.join(); }
    } prefault;
    if (idx_out && ds->n >= (1 << 19)) {
        int32_t* dst = idx_out;
        const int64_t cells = ds->n;
        prefault.t = std::thread([dst, cells] { std::memset(dst, 0, sizeof(int32_t) * (size_t)cells); });
    }
    Session own;
    Session& s = ready ? *ready : own;
    // several ranks: a fit from GIVEN centres is the row-sharded one (every rank holds a shard and the same initial centres, the tables
    // are summed over the ranks each iteration); a fit that seeds itself (sparse_k_means) draws its centres from the local cells and
    // stays local
    if (!ready) {
        s.row_sharded = init_mu_colmajor != nullptr;
        DPR_TRY(single_session(ds, (int)k, reg, no_zero, true, s));
    }
    if (init_mu_colmajor) {
        std::vector<double> rm;
        colmajor_to_rowmajor(init_mu_colmajor, (int)k, ds->d, rm);
        DPR_TRY(s.set_centres(0, rm.data()));
    } else {
        std::vector<SeedJob> jobs(1);
        jobs[0] = SeedJob{ds->x, ds->n, fit_id, s.h_problems[0].mu, nullptr};
        mark("session (+ cell-major copy)");
        DPR_TRY(run_seeding(jobs, ds->d, (int)k, seed));
        mark("k-means++");
    }
    DPR_TRY(s.prepare());
    DPR_TRY(s.zero_labels());
    DPR_TRY(s.lloyd(max_iter));
    DPR_TRY(s.fetch_state());
    mark("Lloyd loop");
    const Problem& pr = s.h_problems[0];
    if (n_iter_out) *n_iter_out = pr.n_assign;
    if (prefault.t.joinable()) prefault.t.join();
    DPR_TRY(copy_labels_out(pr.labels, ds->n, idx_out));
    mark("labels to the host");
    if (centers_out) {
        std::vector<double> rm((size_t)k * ds->d);
        DPR_TRY(s.get_centres(0, rm.data()));
        rowmajor_to_colmajor(rm, (int)k, ds->d, centers_out);
    }
    if (norm_out) DPR_TRY(cluster_norm_dev(pr, ds->d, reg, norm_out, s.row_sharded));
    return DPR_OK;
}

static std::unique_ptr<Session> g_bench_session;
static dpr_dataset* g_bench_ds = nullptr;
static int g_bench_k = 0, g_bench_mode = -1;
void bench_session_forget(dpr_dataset* ds) {
    if (ds == g_bench_ds) {
        g_bench_session.reset();
        g_bench_ds = nullptr;
    }
}

}  // namespace dpr

using namespace dpr;

extern "C" {

const char* dpr_last_error(void) { return g_error.c_str(); }
int dpr_version(void) { return 100; }
int dpr_device_count(int32_t* out) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) n = 0;
    if (out) *out = n;
    return DPR_OK;
}
int dpr_set_device(int32_t ordinal) {
    if (ctx().ready && ctx().device != ordinal) return fail(DPR_ERR_STATE, "device already initialised");
    ctx().device = ordinal;
    return DPR_OK;
}
int dpr_set_stream(void* s) {
    DPR_TRY(ensure_ready());
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));   // device memory is released in stream order: finish the old stream first
    ctx().stream = s ? (cudaStream_t)s : ctx().own_stream;
    return DPR_OK;
}
int dpr_synchronize(void) {
    DPR_TRY(ensure_ready());
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}
int dpr_recheck_count(int64_t* out, int32_t reset) {
    DPR_TRY(ensure_ready());
    if (!out) return fail(DPR_ERR_INVALID, "recheck_count: null output");
    unsigned long long v = 0;
    if (ctx().d_recheck) {
        DPR_CUDA(cudaMemcpyAsync(&v, ctx().d_recheck, sizeof(v), cudaMemcpyDeviceToHost, ctx().stream));
        if (reset) DPR_CUDA(cudaMemsetAsync(ctx().d_recheck, 0, sizeof(v), ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    *out = (int64_t)v;
    return DPR_OK;
}
int dpr_set_pass_kernel(int32_t which) {
    if (which != DPR_PASS_AUTO && which != DPR_PASS_FFMA) return fail(DPR_ERR_INVALID, "unknown pass kernel");
    ctx().force_ffma = which == DPR_PASS_FFMA;
    g_bench_session.reset();   // a cached session keeps the kernel it was planned for
    return DPR_OK;
}
int dpr_set_ari_mode(int32_t mode) {
    if (mode != DPR_ARI_EXACT && mode != DPR_ARI_MONTE_CARLO) return fail(DPR_ERR_INVALID, "unknown ARI mode");
    ctx().ari_mode = mode;
    return DPR_OK;
}
int dpr_pass_timing(int32_t enable, double* ms_sum_out, int64_t* launches_out) {
    DPR_TRY(ensure_ready());
    double sum = 0.0;
    int64_t cnt = 0;
    for (auto& ev : ctx().pass_events) {
        float ms = 0.f;
        if (cudaEventSynchronize(ev.second) == cudaSuccess && cudaEventElapsedTime(&ms, ev.first, ev.second) == cudaSuccess) {
            sum += ms;
            ++cnt;
        }
        cudaEventDestroy(ev.first);
        cudaEventDestroy(ev.second);
    }
    ctx().pass_events.clear();
    ctx().time_passes = enable != 0;
    if (ms_sum_out) *ms_sum_out = sum;
    if (launches_out) *launches_out = cnt;
    return DPR_OK;
}
int dpr_kernel_launches(int64_t* out, int32_t reset) {
    if (out) *out = ctx().launches;
    if (reset) ctx().launches = 0;
    return DPR_OK;
}

// ---- datasets -------------------------------------------------------------------------------------
int dpr_dataset_create(const double* X, int64_t n, int32_t d, dpr_dataset_t* out) {
    if (!X) return fail(DPR_ERR_INVALID, "null X");
    dpr_dataset* ds = nullptr;
    DPR_TRY(dataset_alloc(n, d, &ds));
    int rc = dataset_upload(ds, X);
    if (rc != DPR_OK) { dpr_dataset_destroy(ds); return rc; }
    *out = ds;
    return DPR_OK;
}
int dpr_dataset_create_f32(const float* X, int64_t n, int32_t d, dpr_dataset_t* out) {
    if (!X) return fail(DPR_ERR_INVALID, "null X");
    dpr_dataset* ds = nullptr;
    DPR_TRY(dataset_alloc(n, d, &ds));
    int rc = dataset_upload(ds, X);
    if (rc != DPR_OK) { dpr_dataset_destroy(ds); return rc; }
    *out = ds;
    return DPR_OK;
}
int dpr_dataset_create_synthetic(int64_t n, int32_t d, int32_t groups, double amplitude, uint64_t seed, int64_t row0, double scale,
                                 double* scale_out, dpr_dataset_t* out) {
    if (groups < 1) return fail(DPR_ERR_INVALID, "groups must be >= 1");
    dpr_dataset* ds = nullptr;
    DPR_TRY(dataset_alloc(n, d, &ds));
    // centre coordinates iid from {-a, 0, +a}; group g additionally has ceil(g*d/(2G)) zeroed markers
    std::vector<float> cen((size_t)groups * d);
    for (int g = 0; g < groups; ++g) {
        for (int j = 0; j < d; ++j) {
            const uint32_t r = dpr_rng_draw(seed, 10, (uint32_t)g, (uint32_t)j, 0).w[0] % 3u;
            cen[(size_t)g * d + j] = (float)((double)((int)r - 1) * amplitude);
        }
        const int zeros = (g * d + 2 * groups - 1) / (2 * groups);
        for (int z = 0; z < zeros; ++z) cen[(size_t)g * d + dpr_rng_draw(seed, 10, (uint32_t)g, (uint32_t)z, 1).w[0] % (uint32_t)d] = 0.f;
    }
    DevBuf dc, acc;
    int rc = dc.alloc(sizeof(float) * cen.size());
    if (rc == DPR_OK) rc = acc.alloc(sizeof(double) * 2, true);
    if (rc != DPR_OK) { dpr_dataset_destroy(ds); return rc; }
    cudaMemcpyAsync(dc.p, cen.data(), sizeof(float) * cen.size(), cudaMemcpyHostToDevice, ctx().stream);
    rc = launch_synth(ds->x, n, d, dc.as<float>(), groups, seed, row0, acc.as<double>(), acc.as<double>() + 1, ctx().stream);
    double h[2] = {0, 0};
    cudaMemcpyAsync(h, acc.p, sizeof(h), cudaMemcpyDeviceToHost, ctx().stream);
    cudaStreamSynchronize(ctx().stream);
    const double cnt = (double)n * d, mean = h[0] / cnt;
    const double var = (h[1] - cnt * mean * mean) / (cnt - 1);
    const double used = scale > 0 ? scale : std::sqrt(var);
    if (scale_out) *scale_out = used;
    if (rc == DPR_OK) rc = launch_scale(ds->x, ds->n_pad, d, (float)(1.0 / used), ctx().stream);
    if (rc == DPR_OK) rc = dataset_finish(ds);
    if (rc == DPR_OK && cudaStreamSynchronize(ctx().stream) != cudaSuccess) rc = fail(DPR_ERR_CUDA, "synthetic generation failed");
    if (rc != DPR_OK) { dpr_dataset_destroy(ds); return rc; }
    *out = ds;
    return DPR_OK;
}
// ---- depecheLogCenterSd on the device (R/depecheLogCenterSd.R:6-115) ----------------------------------------
namespace {
// R's pretty(c(lo, hi), n, min.n = 1) as hist.default calls it (R_pretty in src/appl/pretty.c + seq.int of the bounds)
std::vector<double> r_pretty_breaks(double lo, double hi, int ndiv) {
    const double h = 1.5, h5 = 0.5 + 1.5 * h, eps = 1e-10;
    const int min_n = 1;
    double dx = hi - lo, cell;
    bool i_small;
    if (dx == 0 && hi == 0) {
        cell = 1;
        i_small = true;
    } else {
        cell = std::max(std::fabs(lo), std::fabs(hi));
        double U = 1 + ((h5 >= 1.5 * h + .5) ? 1 / (1 + h) : 1.5 / (1 + h5));
        U *= std::max(1, ndiv) * 2.220446049250313e-16;
        i_small = dx < cell * U * 3;
    }
    if (i_small) {
        if (cell > 10) cell = 9 + cell / 10;
        cell *= 0.75;
    } else {
        cell = dx;
        if (ndiv > 1) cell /= ndiv;
    }
    cell = std::max(cell, 20 * 2.2250738585072014e-308);
    const double base = std::pow(10.0, std::floor(std::log10(cell)));
    double unit = base;
    if (2 * base - cell < h * (cell - unit)) {
        unit = 2 * base;
        if (5 * base - cell < h5 * (cell - unit)) {
            unit = 5 * base;
            if (10 * base - cell < h * (cell - unit)) unit = 10 * base;
        }
    }
    double ns = std::floor(lo / unit + eps), nu = std::ceil(hi / unit - eps);
    while (ns * unit > lo + eps * unit) ns--;
    while (nu * unit < hi - eps * unit) nu++;
    int k = (int)(0.5 + nu - ns);
    if (k < min_n) {
        k = min_n - k;
        if (ns >= 0.) { nu += k / 2; ns -= k / 2 + k % 2; }
        else { ns -= k / 2; nu += k / 2 + k % 2; }
        k = min_n;
    }
    // seq.int(l, u, length.out = k + 1): ends exact, interior l + i * by
    const double l = ns * unit, u = nu * unit;
    std::vector<double> b((size_t)k + 1);
    const double by = k > 0 ? (u - l) / k : 0.0;
    for (int i = 0; i <= k; ++i) b[i] = l + i * by;
    b[0] = l;
    b[k] = u;
    const double delta = k > 0 ? (u - l) / k : 0.0;
    for (double& v : b)
        if (std::fabs(v) < 1e-14 * delta) v = 0.0;
    return b;
}
double median_of(std::vector<double> v) {   // (selection, not a sort: hist() of 10M cells has 200 000 breaks per marker)
    const size_t m = v.size();
    std::nth_element(v.begin(), v.begin() + m / 2, v.end());
    const double hi = v[m / 2];
    if (m % 2) return hi;
    const double lo = *std::max_element(v.begin(), v.begin() + m / 2);
    return 0.5 * (lo + hi);
}
// median(diff(b)) — the same value median_of gives for the vector of differences, without building it: the breaks are l + i * by in
// floating point, so their differences take a handful of distinct values (by and its neighbours); they are tallied in one sweep and the
// middle element(s) read off the sorted tally.  More than 32 distinct values: the general route.
double median_of_diffs(const std::vector<double>& b) {
    const size_t m = b.size() - 1;
    struct Tally { double v; size_t c; } t[32];
    int nt = 0, last = 0;
    bool general = false;
    for (size_t i = 0; i < m && !general; ++i) {
        const double dv = b[i + 1] - b[i];
        if (nt && t[last].v == dv) { ++t[last].c; continue; }
        int at = -1;
        for (int q = 0; q < nt; ++q)
            if (t[q].v == dv) { at = q; break; }
        if (at < 0) {
            if (nt == 32) { general = true; break; }
            at = nt++;
            t[at] = Tally{dv, 0};
        }
        ++t[at].c;
        last = at;
    }
    if (general) {
        std::vector<double> diffs(m);
        for (size_t i = 0; i < m; ++i) diffs[i] = b[i + 1] - b[i];
        return median_of(std::move(diffs));
    }
    std::sort(t, t + nt, [](const Tally& a, const Tally& c) { return a.v < c.v; });
    auto at_rank = [&](size_t r) {   // the r-th smallest difference (0-based)
        size_t seen = 0;
        for (int q = 0; q < nt; ++q) {
            seen += t[q].c;
            if (r < seen) return t[q].v;
        }
        return t[nt - 1].v;
    };
    const double hi = at_rank(m / 2);
    return m % 2 ? hi : 0.5 * (at_rank(m / 2 - 1) + hi);
}
}  // namespace

int dpr_dataset_create_scaled(const double* X, int64_t n, int32_t d, int32_t log2_off, int32_t center_mode, double* centers_inout,
                              int32_t scale_mode, double* scale_inout, int32_t* log2_applied_out, dpr_dataset_t* out) {
    if (!X || n < 2 || d < 1 || !out) return fail(DPR_ERR_INVALID, "create_scaled: null argument or too few rows");
    if (center_mode < 0 || center_mode > 3 || scale_mode < 0 || scale_mode > 2) return fail(DPR_ERR_INVALID, "create_scaled: unknown mode");
    if ((center_mode == 3 && !centers_inout) || (scale_mode == 2 && !scale_inout)) return fail(DPR_ERR_INVALID, "create_scaled: missing values");
    static const bool timing = getenv("DPR_TIMING") != nullptr;   // diagnostics: host wall-clock per phase (each closed by a stream synchronize)
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char* what) {
        if (!timing) return;
        cudaStreamSynchronize(ctx().stream);
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "[dpr] create_scaled: %s %.2f ms\n", what, std::chrono::duration<double, std::milli>(now - t_prev).count());
        t_prev = now;
    };
    dpr_dataset* ds = nullptr;
    DPR_TRY(dataset_alloc(n, d, &ds));
    std::unique_ptr<dpr_dataset, int (*)(dpr_dataset*)> guard(ds, dpr_dataset_destroy);
    const int B = stat_blocks();
    const int64_t total = n * d;
    DevBuf raw, part, dvec;
    DPR_TRY(raw.alloc(sizeof(double) * total));
    DPR_TRY(part.alloc(sizeof(double) * 3 * B * d));
    DPR_TRY(dvec.alloc(sizeof(double) * 2 * d));
    // The raw values to the device.  Cytometry matrices are usually fp32 values stored as R doubles: host threads narrow each chunk and
    // check that every value survives the trip unchanged; while that holds, half the bytes cross PCIe and a kernel widens them back
    // (exactly).  The first value that does not survive sends the whole matrix as doubles instead.
    bool compact = true;
    {
        const int64_t chunk = std::min<int64_t>(total, kUploadChunk);
        DPR_TRY(upload_stage(chunk));
        bool stopped = false;
        DPR_TRY(staged_pipeline(
            total, chunk, host_copy_threads(),
            [&](int64_t e0, int64_t, int64_t a, int64_t b, int buf) { return narrow_check(X + e0, g_stage.pinned[buf], a, b); },
            [&](int buf, int64_t e0, int64_t cnt) -> int {
                DPR_CUDA(cudaMemcpyAsync(g_stage.dev[buf], g_stage.pinned[buf], sizeof(float) * cnt, cudaMemcpyHostToDevice, ctx().stream));
                return launch_widen(g_stage.dev[buf], raw.as<double>() + e0, cnt, ctx().stream);
            },
            &stopped));
        compact = !stopped;
        if (!compact) DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    // raw doubles to the device through the pinned staging area (as bytes)
    if (!compact) {
        const int64_t chunk = std::min<int64_t>(total, kUploadChunk / 2);
        DPR_TRY(upload_stage(2 * chunk));
        DPR_TRY(staged_pipeline(
            total, chunk, host_copy_threads(),
            [&](int64_t e0, int64_t, int64_t a, int64_t b, int buf) {
                copy_piece(X + e0 + a, reinterpret_cast<double*>(g_stage.pinned[buf]) + a, b - a);
                return true;
            },
            [&](int buf, int64_t e0, int64_t cnt) -> int {
                DPR_CUDA(cudaMemcpyAsync(raw.as<double>() + e0, g_stage.pinned[buf], sizeof(double) * cnt, cudaMemcpyHostToDevice, ctx().stream));
                return DPR_OK;
            }));
    }
    mark("allocations + upload");
    std::vector<double> hp((size_t)3 * B * d), colsum(d), colmin(d), colmax(d);
    auto col_stats = [&]() -> int {
        double* p = part.as<double>();
        DPR_TRY(launch_col_sum_minmax(raw.as<double>(), n, d, p, p + (size_t)B * d, p + (size_t)2 * B * d, ctx().stream));
        DPR_CUDA(cudaMemcpyAsync(hp.data(), p, sizeof(double) * 3 * B * d, cudaMemcpyDeviceToHost, ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));
        for (int j = 0; j < d; ++j) {
            double s = 0, lo = INFINITY, hi = -INFINITY;
            for (int q = 0; q < B; ++q) {
                s += hp[(size_t)j * B + q];
                lo = std::min(lo, hp[(size_t)B * d + (size_t)j * B + q]);
                hi = std::max(hi, hp[(size_t)2 * B * d + (size_t)j * B + q]);
            }
            colsum[j] = s;
            colmin[j] = lo;
            colmax[j] = hi;
        }
        return DPR_OK;
    };
    auto moments = [&](const std::vector<double>& shift, double* s2, double* s4) -> int {
        double* p = part.as<double>();
        DPR_CUDA(cudaMemcpyAsync(dvec.p, shift.data(), sizeof(double) * d, cudaMemcpyHostToDevice, ctx().stream));
        DPR_TRY(launch_col_moments(raw.as<double>(), n, d, dvec.as<double>(), p, p + (size_t)B * d, ctx().stream));
        DPR_CUDA(cudaMemcpyAsync(hp.data(), p, sizeof(double) * 2 * B * d, cudaMemcpyDeviceToHost, ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));
        double a = 0, b4 = 0;
        for (size_t e = 0; e < (size_t)B * d; ++e) {
            a += hp[e];
            b4 += hp[(size_t)B * d + e];
        }
        *s2 = a;
        *s4 = b4;
        return DPR_OK;
    };
    DPR_TRY(col_stats());
    // 1. log2 when the pooled kurtosis exceeds 100 (:11-45)
    int log2_applied = 0;
    if (!log2_off) {
        double tot = 0;
        for (int j = 0; j < d; ++j) tot += colsum[j];
        const double mean = tot / (double)total;
        double s2, s4;
        DPR_TRY(moments(std::vector<double>(d, mean), &s2, &s4));
        const double kurt = (double)total * s4 / (s2 * s2);
        if (kurt > 100) {
            double gmin = INFINITY;
            for (int j = 0; j < d; ++j) gmin = std::min(gmin, colmin[j]);
            DPR_CUDA(cudaMemcpyAsync(dvec.p, colmin.data(), sizeof(double) * d, cudaMemcpyHostToDevice, ctx().stream));
            DPR_TRY(launch_log2(raw.as<double>(), n, d, dvec.as<double>(), gmin >= 0 ? 0 : 1, ctx().stream));
            log2_applied = 1;
            DPR_TRY(col_stats());
        }
    }
    mark("column statistics + kurtosis");
    // 2. centres (:49-98)
    std::vector<double> centre(d, 0.0);
    if (center_mode == 1) {   // peak of hist(x, breaks = n / 50 | 10)   (dScaleCoFunction.R:74-88)
        const int n_breaks = n < 500 ? 10 : (int)((double)n / 50.0);
        // the breaks of every marker on the host (threads: one marker each; only their parameters go to the device, which recomputes a
        // break where it needs one), ONE histogram launch over all markers, one read of the d winners
        std::vector<std::vector<double>> br(d);
        std::vector<HistCol> hc(d);
        std::vector<int> bad(d, 0);
        {
            auto prep = [&](int j) {
                br[j] = r_pretty_breaks(colmin[j], colmax[j], n_breaks);
                const int nb = (int)br[j].size() - 1;
                if (nb < 1) {
                    bad[j] = 1;
                    return;
                }
                const double range = colmax[j] - colmin[j];
                HistCol c{};
                c.col = raw.as<double>() + (int64_t)j * n;
                c.nb = nb;
                c.diddle = 1e-7 * (nb + 1 > 5 ? median_of_diffs(br[j]) : (nb + 1 <= 3 ? range : std::max(range, median_of_diffs(br[j]))));
                c.inv_by = 1.0 / (br[j][nb / 2 + 1] - br[j][nb / 2]);
                // what r_pretty_breaks built the sequence from (its interior values are l + i * by)
                c.l = br[j][0];
                c.u = br[j][nb];
                c.by = (c.u - c.l) / nb;
                // (a break that was snapped to exactly zero has |l + i * by| < 1e-14 * by; the device applies the same rule.  The first and
                // last break are exact ends, never snapped: check that this holds, otherwise the table path would be needed)
                if (std::fabs(c.l) < 1e-14 * c.by && c.l != 0.0) bad[j] = 2;
                hc[j] = c;
            };
            const int threads = std::max(1, std::min(host_copy_threads(), (int)d));
            std::vector<std::thread> pool;
            for (int t = 0; t < threads; ++t)
                pool.emplace_back([&, t] {
                    for (int j = t; j < d; j += threads) prep(j);
                });
            for (auto& th : pool) th.join();
        }
        mark("  breaks on the host");
        size_t tot_bins = 0;
        for (int j = 0; j < d; ++j) {
            if (bad[j]) return fail(DPR_ERR_INVALID, "create_scaled: degenerate histogram");
            tot_bins += (size_t)hc[j].nb;
        }
        DevBuf cols_dev, counts, amax;
        DPR_TRY(cols_dev.alloc(sizeof(HistCol) * d));
        DPR_TRY(counts.alloc(sizeof(unsigned int) * tot_bins));
        DPR_TRY(amax.alloc(sizeof(int) * d));
        size_t at = 0;
        for (int j = 0; j < d; ++j) {
            hc[j].counts = counts.as<unsigned int>() + at;
            at += (size_t)hc[j].nb;
        }
        DPR_CUDA(cudaMemcpyAsync(cols_dev.p, hc.data(), sizeof(HistCol) * d, cudaMemcpyHostToDevice, ctx().stream));
        DPR_TRY(launch_hist_cols(cols_dev.as<HistCol>(), d, n, counts.as<unsigned int>(), tot_bins, amax.as<int>(), ctx().stream));
        std::vector<int> best(d, 0);
        DPR_CUDA(cudaMemcpyAsync(best.data(), amax.p, sizeof(int) * d, cudaMemcpyDeviceToHost, ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));   // (also keeps hc alive until its copy is done)
        for (int j = 0; j < d; ++j) centre[j] = 0.5 * (br[j][best[j]] + br[j][best[j] + 1]);
    } else if (center_mode == 2) {
        for (int j = 0; j < d; ++j) centre[j] = colsum[j] / (double)n;
    } else if (center_mode == 3) {
        for (int j = 0; j < d; ++j) centre[j] = centers_inout[j];
    }
    mark("centres");
    // 3. one sd over all centred values (:102-106)
    double sd = 1.0;
    if (scale_mode == 1) {
        double tot = 0;
        for (int j = 0; j < d; ++j) tot += colsum[j] - (double)n * centre[j];
        const double vmean = tot / (double)total;
        std::vector<double> shift(d);
        for (int j = 0; j < d; ++j) shift[j] = centre[j] + vmean;
        double s2, s4;
        DPR_TRY(moments(shift, &s2, &s4));
        sd = std::sqrt(s2 / (double)(total - 1));
    } else if (scale_mode == 2) {
        sd = *scale_inout;
    }
    if (!(sd > 0) || !std::isfinite(sd)) return fail(DPR_ERR_INVALID, "create_scaled: the standard deviation is not positive and finite");
    DPR_CUDA(cudaMemcpyAsync(dvec.p, centre.data(), sizeof(double) * d, cudaMemcpyHostToDevice, ctx().stream));
    DPR_TRY(launch_scale_to_tiles(raw.as<double>(), n, d, dvec.as<double>(), sd, ds->x, ctx().stream));
    DPR_TRY(dataset_finish(ds));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    mark("sd + scaling into tiles");
    if (centers_inout && center_mode != 0)
        for (int j = 0; j < d; ++j) centers_inout[j] = centre[j];
    if (scale_inout) *scale_inout = sd;
    if (log2_applied_out) *log2_applied_out = log2_applied;
    *out = guard.release();
    return DPR_OK;
}

int dpr_dataset_destroy(dpr_dataset_t ds) {
    if (!ds) return DPR_OK;
    bench_session_forget(ds);
    dev_free(ds->x);
    dev_free(ds->tile_xx);
    dev_free(ds->labels);
    dev_free(ds->rows);
    delete ds;
    return DPR_OK;
}
int dpr_dataset_shape(dpr_dataset_t ds, int64_t* n, int32_t* d) {
    if (!ds) return fail(DPR_ERR_INVALID, "null dataset");
    if (n) *n = ds->n;
    if (d) *d = ds->d;
    return DPR_OK;
}
int dpr_dataset_read(dpr_dataset_t ds, int64_t row0, int64_t rows, double* out) {
    if (!ds || !out || row0 < 0 || rows < 1 || row0 + rows > ds->n) return fail(DPR_ERR_INVALID, "bad row range");
    DevBuf tmp;
    DPR_TRY(tmp.alloc(sizeof(double) * rows * ds->d));
    DPR_TRY(launch_cols_to_f64(ds->x, row0, rows, ds->d, tmp.as<double>(), ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(out, tmp.p, sizeof(double) * rows * ds->d, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}
int dpr_dataset_gather(dpr_dataset_t ds, const int64_t* rows, int64_t m, dpr_dataset_t* out) {
    if (!ds || !rows || m < 1 || !out) return fail(DPR_ERR_INVALID, "bad gather arguments");
    for (int64_t i = 0; i < m; ++i)
        if (rows[i] < 0 || rows[i] >= ds->n) return fail(DPR_ERR_INVALID, "row index out of range");
    dpr_dataset* g = nullptr;
    DPR_TRY(dataset_alloc(m, ds->d, &g));
    DevBuf dr;
    int rc = dr.alloc(sizeof(int64_t) * m);
    if (rc == DPR_OK && cudaMemcpyAsync(dr.p, rows, sizeof(int64_t) * m, cudaMemcpyHostToDevice, ctx().stream) != cudaSuccess)
        rc = fail(DPR_ERR_CUDA, "copy of row indices failed");
    if (rc == DPR_OK) rc = launch_gather(ds->x, ds->n, ds->d, dr.as<int64_t>(), 0, 0, m, g->x, 1, 0, ctx().stream);
    if (rc == DPR_OK) rc = dataset_finish(g);
    if (rc == DPR_OK && cudaStreamSynchronize(ctx().stream) != cudaSuccess) rc = fail(DPR_ERR_CUDA, "gather failed");
    if (rc != DPR_OK) { dpr_dataset_destroy(g); return rc; }
    *out = g;
    return DPR_OK;
}

// ---- allocate_points -------------------------------------------------------------------------------
int dpr_dataset_allocate_points(dpr_dataset_t ds, const double* mu, int32_t k, int32_t no_zero, int32_t* idx_out) {
    return dpr_dataset_allocate_points_base(ds, mu, k, no_zero, 0, idx_out);
}
// depeche()'s final step (R/depeche.R:228-252): allocate, then renumber the clusters 1, 2, ... by decreasing size, ties in the order of
// the old numbers (`sort(table(.), decreasing = TRUE)`).  The sizes are counted on the device, the new numbers go through a lookup
// table while the labels are copied out: no pass over the host vector.  order_out (k entries): the old 0-based label of rank 1, 2, ...
// (-1 beyond the clusters that own cells).
int dpr_dataset_allocate_points_ranked(dpr_dataset_t ds, const double* mu, int32_t k, int32_t no_zero, int32_t* idx_out, int32_t* order_out) {
    if (!ds || !mu || k < 1 || !idx_out || !order_out) return fail(DPR_ERR_INVALID, "allocate_points_ranked: null argument or k < 1");
    Session s;
    DPR_TRY(single_session(ds, k, 0.0, no_zero, false, s));
    std::vector<double> rm;
    colmajor_to_rowmajor(mu, k, ds->d, rm);
    DPR_TRY(s.set_centres(0, rm.data()));
    DPR_TRY(s.prepare());
    DPR_TRY(s.pass(kPassAssign, false));
    DevBuf cnt;
    DPR_TRY(cnt.alloc(sizeof(unsigned long long) * k));
    DPR_TRY(launch_label_counts(ds->labels, ds->n, k, cnt.as<unsigned long long>(), ctx().stream));
    std::vector<unsigned long long> h((size_t)k);
    DPR_CUDA(cudaMemcpyAsync(h.data(), cnt.p, sizeof(unsigned long long) * k, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    std::vector<int32_t> order;
    for (int l = 0; l < k; ++l)
        if (h[l]) order.push_back(l);
    std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return h[a] > h[b]; });
    std::vector<int32_t> lut((size_t)k, 0);
    for (int l = 0; l < k; ++l) order_out[l] = -1;
    for (size_t r = 0; r < order.size(); ++r) {
        lut[order[r]] = (int32_t)r + 1;
        order_out[r] = order[r];
    }
    return copy_labels_out(ds->labels, ds->n, idx_out, 0, lut.data());
}
int dpr_dataset_allocate_points_base(dpr_dataset_t ds, const double* mu, int32_t k, int32_t no_zero, int32_t base, int32_t* idx_out) {
    if (!ds || !mu || k < 1) return fail(DPR_ERR_INVALID, "allocate_points: null argument or k < 1");
    Session s;
    DPR_TRY(single_session(ds, k, 0.0, no_zero, false, s));
    std::vector<double> rm;
    colmajor_to_rowmajor(mu, k, ds->d, rm);
    DPR_TRY(s.set_centres(0, rm.data()));
    DPR_TRY(s.prepare());
    DPR_TRY(s.pass(kPassAssign, false));
    return copy_labels_out(ds->labels, ds->n, idx_out, base);
}
int dpr_allocate_points(const double* X, int64_t n, int32_t d, const double* mu, int32_t k, int32_t no_zero, int32_t* idx_out) {
    if (!idx_out) return fail(DPR_ERR_INVALID, "null output");
    dpr_dataset_t ds = nullptr;
    DPR_TRY(dpr_dataset_create(X, n, d, &ds));
    const int rc = dpr_dataset_allocate_points(ds, mu, k, no_zero, idx_out);
    dpr_dataset_destroy(ds);
    return rc;
}

// ---- sparse_k_means ----------------------------------------------------------------------------------
int dpr_dataset_sparse_k_means(dpr_dataset_t ds, uint32_t k, double reg, int32_t no_zero, uint64_t seed, int32_t* idx_out,
                               double* centers_out, double* norm_out, int32_t* n_iter_out) {
    return find_centers_dev(ds, nullptr, k, reg, no_zero, 1000, seed, 0, idx_out, centers_out, norm_out, n_iter_out);
}
int dpr_sparse_k_means(const double* X, int64_t n, int32_t d, uint32_t k, double reg, int32_t no_zero, uint64_t seed,
                       int32_t* idx_out, double* centers_out, double* norm_out, int32_t* n_iter_out) {
    if (!X) return fail(DPR_ERR_INVALID, "null X");
    if (k < 2) return fail(DPR_ERR_INVALID, "k must be >= 2 (initialize_mu writes two rows, Clusterer.cpp:157-162)");
    dpr_dataset* ds = nullptr;
    DPR_TRY(dataset_alloc(n, d, &ds));
    std::unique_ptr<dpr_dataset, int (*)(dpr_dataset*)> guard(ds, dpr_dataset_destroy);
    // (a matrix that lives for this one fit: the cell-major copy the delta iterations would gather from costs more to make — 1.6 GB of pool
    // memory and a pass over the cells — than it saves in twenty-odd iterations: 0.062-0.070 -> 0.059 s for 10M x 40)
    ds->rows_tried = true;
    // the session's tables are allocated, cleared and described to the device BEFORE the upload: that work (and its synchronisation) then
    // overlaps with the first chunks the host threads narrow instead of standing between the upload and the seeding
    Session s;
    DPR_TRY(single_session(ds, (int)k, reg, no_zero, true, s));
    const auto t0 = std::chrono::steady_clock::now();
    DPR_TRY(dataset_upload(ds, X));
    if (getenv("DPR_TIMING")) fprintf(stderr, "[dpr] sparse_k_means: upload %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
    return find_centers_dev(ds, nullptr, k, reg, no_zero, 1000, seed, 0, idx_out, centers_out, norm_out, n_iter_out, &s);
}
int dpr_dataset_find_centers_from(dpr_dataset_t ds, const double* init_mu, uint32_t k, double reg, int32_t no_zero, int32_t max_iter,
                                  int32_t* idx_out, double* centers_out, double* norm_out, int32_t* n_iter_out) {
    if (!init_mu) return fail(DPR_ERR_INVALID, "null initial centres");
    return find_centers_dev(ds, init_mu, k, reg, no_zero, max_iter, 0, 0, idx_out, centers_out, norm_out, n_iter_out);
}

int dpr_dataset_initialize_mu(dpr_dataset_t ds, uint32_t k, uint64_t seed, uint32_t fit_id, double* mu_out, int64_t* rows_out) {
    if (!ds || !mu_out || k < 2) return fail(DPR_ERR_INVALID, "initialize_mu: null argument or k < 2");
    DPR_TRY(ensure_ready());
    DevBuf mu, rows;
    DPR_TRY(mu.alloc(sizeof(double) * k * ds->d, true));
    DPR_TRY(rows.alloc(sizeof(long long) * k, true));
    std::vector<SeedJob> jobs(1);
    jobs[0] = SeedJob{ds->x, ds->n, fit_id, mu.as<double>(), rows.as<long long>()};
    DPR_TRY(run_seeding(jobs, ds->d, (int)k, seed));
    std::vector<double> rm((size_t)k * ds->d);
    std::vector<long long> hr(k);
    DPR_CUDA(cudaMemcpyAsync(rm.data(), mu.p, sizeof(double) * rm.size(), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(hr.data(), rows.p, sizeof(long long) * k, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    rowmajor_to_colmajor(rm, (int)k, ds->d, mu_out);
    if (rows_out)
        for (uint32_t c = 0; c < k; ++c) rows_out[c] = hr[c];
    return DPR_OK;
}

int dpr_bootstrap_rows(int64_t n, uint32_t boot_n, uint64_t seed, uint32_t fit_id, int64_t* rows_out) {
    if (n < 1 || boot_n < 1 || !rows_out) return fail(DPR_ERR_INVALID, "bootstrap_rows: bad arguments");
    DPR_TRY(ensure_ready());
    DevBuf rows;
    DPR_TRY(rows.alloc(sizeof(int64_t) * boot_n));
    DPR_TRY(launch_bootstrap_rows(n, boot_n, seed, fit_id, rows.as<int64_t>(), ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(rows_out, rows.p, sizeof(int64_t) * boot_n, cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}

// reevaluate_centers for explicit labels (Clusterer.cpp:89-108)
int dpr_dataset_reevaluate_centers(dpr_dataset_t ds, const int32_t* idx, uint32_t k, double reg, double* centers_out,
                                   int64_t* counts_out) {
    if (!ds || !idx || !centers_out || k < 1) return fail(DPR_ERR_INVALID, "reevaluate_centers: bad arguments");
    for (int64_t i = 0; i < ds->n; ++i)
        if (idx[i] < 0 || (uint32_t)idx[i] >= k) return fail(DPR_ERR_INVALID, "label outside [0, k)");
    Session s;
    s.row_sharded = true;   // several ranks: one update from the labels of every rank's shard
    DPR_TRY(single_session(ds, (int)k, reg, 0, true, s));
    DPR_CUDA(cudaMemcpyAsync(ds->labels, idx, sizeof(int32_t) * ds->n, cudaMemcpyHostToDevice, ctx().stream));
    DPR_TRY(s.prepare());
    DPR_TRY(s.pass(kPassSums, false));
    // iter = 20 makes the ramp factor exactly 1 (reg_i = reg); max_iter large so no cap handling
    DPR_TRY(s.reduce_and_finalize(20, 1 << 30));
    std::vector<double> rm((size_t)k * ds->d);
    DPR_TRY(s.get_centres(0, rm.data()));
    rowmajor_to_colmajor(rm, (int)k, ds->d, centers_out);
    if (counts_out) {
        std::vector<long long> c(k);
        DPR_CUDA(cudaMemcpyAsync(c.data(), s.h_problems[0].counts, sizeof(long long) * k, cudaMemcpyDeviceToHost, ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));
        for (uint32_t i = 0; i < k; ++i) counts_out[i] = c[i];
    }
    return DPR_OK;
}

int dpr_dataset_cluster_norm(dpr_dataset_t ds, const double* mu, uint32_t k, const int32_t* idx, double reg, double* out) {
    if (!ds || !mu || !idx || !out || k < 1) return fail(DPR_ERR_INVALID, "cluster_norm: bad arguments");
    for (int64_t i = 0; i < ds->n; ++i)
        if (idx[i] < 0 || (uint32_t)idx[i] >= k) return fail(DPR_ERR_INVALID, "label outside [0, k)");
    DPR_TRY(dataset_labels(ds));
    DevBuf dmu;
    DPR_TRY(dmu.alloc(sizeof(double) * k * ds->d));
    std::vector<double> rm;
    colmajor_to_rowmajor(mu, (int)k, ds->d, rm);
    DPR_CUDA(cudaMemcpyAsync(dmu.p, rm.data(), sizeof(double) * rm.size(), cudaMemcpyHostToDevice, ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(ds->labels, idx, sizeof(int32_t) * ds->n, cudaMemcpyHostToDevice, ctx().stream));
    Problem pr{};
    pr.x = ds->x;
    pr.n = ds->n;
    pr.k = (int)k;
    pr.mu = dmu.as<double>();
    pr.labels = ds->labels;
    return cluster_norm_dev(pr, ds->d, reg, out);
}

int dpr_n_used_clusters(const int32_t* idx, int64_t n, uint32_t k, uint32_t* out) {
    if (!idx || !out || n < 1) return fail(DPR_ERR_INVALID, "n_used_clusters: bad arguments");
    DPR_TRY(ensure_ready());
    DevBuf lab, present, bad;
    DPR_TRY(lab.alloc(sizeof(int32_t) * n));
    DPR_TRY(present.alloc(sizeof(int32_t) * (k + 1), true));
    DPR_TRY(bad.alloc(sizeof(int32_t), true));
    DPR_CUDA(cudaMemcpyAsync(lab.p, idx, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx().stream));
    DPR_TRY(launch_label_presence(lab.as<int32_t>(), n, (int)k + 1, present.as<int32_t>(), bad.as<int32_t>(), ctx().stream));
    std::vector<int32_t> h(k + 1);
    int32_t hbad = 0;
    DPR_CUDA(cudaMemcpyAsync(h.data(), present.p, sizeof(int32_t) * (k + 1), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(&hbad, bad.p, sizeof(int32_t), cudaMemcpyDeviceToHost, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    if (hbad) return fail(DPR_ERR_INVALID, "label outside [0, k]");
    uint32_t c = 0;
    for (int32_t v : h) c += v != 0;
    *out = c;
    return DPR_OK;
}

// ---- rand_index ------------------------------------------------------------------------------------------
int dpr_rand_index(const int32_t* a, const int32_t* b, int64_t n, uint32_t k, double* out) {
    if (!a || !b || !out || n < 2) return fail(DPR_ERR_INVALID, "rand_index: null argument or n < 2");
    DPR_TRY(ensure_ready());
    DevBuf da, db;
    DPR_TRY(da.alloc(sizeof(int32_t) * n));
    DPR_TRY(db.alloc(sizeof(int32_t) * n));
    DPR_CUDA(cudaMemcpyAsync(da.p, a, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx().stream));
    DPR_CUDA(cudaMemcpyAsync(db.p, b, sizeof(int32_t) * n, cudaMemcpyHostToDevice, ctx().stream));
    std::vector<PairJob> jobs(1);
    jobs[0] = PairJob{da.as<int32_t>(), db.as<int32_t>(), n, 0};
    return run_ari(jobs, k, 0, out);
}

// ---- grid_search (optimize_param, Clusterer.cpp:334-394) ---------------------------------------------------
// All (iteration, k, penalty) cells are independent: the 2 * iterations * nk * nreg bootstrap fits run as ONE
// batch of problems (every launch serves all of them), then the 2 * ... scoring assignments of the full data
// run as one batch, then one batched ARI launch.  The two extra allocations of the reference whose results
// are discarded (:373-374) are not performed.
}  // extern "C"

// `cells`: the problems [c_lo, c_hi) of the grid in the reference's loop order c = (it * nk + j) * nreg + l; results are written
// compactly (index c - c_lo): ari / nclust one value per problem, centres ret1, ret2 per problem.
static int grid_search_range(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg, uint32_t iterations,
                             uint32_t boot_n, uint64_t seed, int c_lo, int c_hi, double* ari_out, double* nclust_out, double* centers_out) {
    if (!ds || !k || !reg || !ari_out || !nclust_out || nk < 1 || nreg < 1 || iterations < 1 || boot_n < 1)
        return fail(DPR_ERR_INVALID, "grid_search: bad arguments");
    for (int j = 0; j < nk; ++j)
        if (k[j] < 2) return fail(DPR_ERR_INVALID, "grid_search: every k must be >= 2");
    if (c_lo < 0 || c_hi > (int)iterations * nk * nreg || c_lo >= c_hi) return fail(DPR_ERR_INVALID, "grid_search: bad problem range");
    DPR_TRY(ensure_ready());
    static const bool timing = getenv("DPR_TIMING") != nullptr;   // diagnostics: host wall-clock per phase (each closed by a stream synchronize)
    auto now = [&]() {
        if (timing) cudaStreamSynchronize(ctx().stream);
        return std::chrono::steady_clock::now();
    };
    auto ms_since = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
        return std::chrono::duration<double, std::milli>(b - a).count();
    };
    const auto t_start = now();
    const int d = ds->d;
    const int cells = c_hi - c_lo;
    const int F = 2 * cells;
    const uint32_t fit0 = 2u * (uint32_t)c_lo;   // global number of the first fit: fit f = 2 * problem + which
    const int64_t ldb = ((int64_t)boot_n + kSlotCells - 1) / kSlotCells * kSlotCells;
    // 1. bootstrap samples (bootstrap_data, :354-355)
    DevBuf xb, lab_b;
    const int64_t tiles_b = ldb / kSlotCells;
    const int64_t fit_floats = tiles_b * (int64_t)tile_floats(d);   // floats of one bootstrap matrix
    DPR_TRY(xb.alloc(sizeof(float) * (size_t)F * fit_floats, true));
    DPR_TRY(lab_b.alloc(sizeof(int32_t) * (size_t)F * ldb, true));
    DPR_TRY(launch_gather(ds->x, ds->n, d, nullptr, seed, fit0, boot_n, xb.as<float>(), F, fit_floats, ctx().stream));
    DevBuf xb_xx;   // the F bootstrap matrices are back to back and each is a whole number of tiles
    DPR_TRY(xb_xx.alloc(sizeof(float) * (size_t)F * tiles_b));
    DPR_TRY(launch_tile_norms(xb.as<float>(), (int64_t)F * tiles_b, d, xb_xx.as<float>(), ctx().stream));
    // 2. the fits (find_centers with zero-centre removal, :361-362)
    std::vector<ProblemSpec> specs(F);
    auto cell_of = [&](int f, int& it, int& j, int& l) {
        const int c = c_lo + f / 2;
        l = c % nreg;
        j = (c / nreg) % nk;
        it = c / (nreg * nk);
    };
    for (int f = 0; f < F; ++f) {
        int it, j, l;
        cell_of(f, it, j, l);
        specs[f] = ProblemSpec{xb.as<float>() + (size_t)f * fit_floats, xb_xx.as<float>() + (size_t)f * tiles_b, ds->tile_xx + ds->n_pad / kSlotCells, (int64_t)boot_n, k[j], reg[l], 1,
                               lab_b.as<int32_t>() + (size_t)f * ldb};
    }
    const auto t_boot = now();
    Session fits;
    DPR_TRY(fits.init(specs, d, true));
    {
        std::vector<SeedJob> jobs(F);
        for (int f = 0; f < F; ++f) jobs[f] = SeedJob{specs[f].x, (int64_t)boot_n, fit0 + (uint32_t)f, fits.h_problems[f].mu, nullptr};
        // seeding rounds are batched per k (the number of rounds is k - 1)
        for (int j = 0; j < nk; ++j) {
            std::vector<SeedJob> sub;
            for (int f = 0; f < F; ++f) {
                int it, jj, l;
                cell_of(f, it, jj, l);
                if (jj == j) sub.push_back(jobs[f]);
            }
            if (!sub.empty()) DPR_TRY(run_seeding(sub, d, k[j], seed));
        }
    }
    const auto t_seed = now();
    DPR_TRY(fits.prepare());
    DPR_TRY(fits.lloyd(1000));
    DPR_TRY(fits.fetch_state());
    const auto t_fit = now();
    // distinct labels of each bootstrap fit (n_used_clusters, :379-380) = clusters with a non-zero count
    std::vector<long long> counts((size_t)F * fits.k_max());
    DPR_CUDA(cudaMemcpyAsync(counts.data(), fits.h_problems[0].counts, sizeof(long long) * counts.size(), cudaMemcpyDeviceToHost, ctx().stream));
    // 3. score on the full data, no_zero = false (:370-371), and 4. ARI per problem.  The label arrays of a batch of
    //    problems are bounded (4 GB), so a 100M-cell matrix is scored a few problems at a time instead of all at once.
    int kmax = 0;
    for (int j = 0; j < nk; ++j) kmax = std::max(kmax, (int)k[j]);
    std::vector<double> ari(cells);
    bool scored = false;
    if (ctx().ari_mode == DPR_ARI_EXACT) {
        // the scoring pass takes the K x K contingency counts of each (ret1, ret2) pair itself (score_pass.cu): every problem of the
        // range in ONE session, no label ever written to HBM, then one small launch turns the tables into ARI values
        std::vector<ProblemSpec> sspecs(2 * cells);
        for (int q = 0; q < 2 * cells; ++q)
            sspecs[q] = ProblemSpec{ds->x, ds->tile_xx, ds->tile_xx + ds->n_pad / kSlotCells, ds->n, specs[q].k, 0.0, 0, nullptr};
        Session score;
        const int kk = kmax + 1;
        DPR_TRY(score.init(sspecs, d, false, kk));
        if (score.fuse_pairs) {
            DevBuf tables, out;
            DPR_TRY(tables.alloc(sizeof(unsigned long long) * (size_t)cells * kk * kk, true));
            DPR_TRY(out.alloc(sizeof(double) * cells));
            score.pair_tables = tables.as<unsigned long long>();
            if (score.k_max() == fits.k_max())   // both sessions keep their centre matrices back to back with the same stride: one copy
                DPR_CUDA(cudaMemcpyAsync(score.h_problems[0].mu, fits.h_problems[0].mu, sizeof(double) * (size_t)fits.k_max() * d * (size_t)(2 * cells),
                                         cudaMemcpyDeviceToDevice, ctx().stream));
            else
                for (int q = 0; q < 2 * cells; ++q) DPR_TRY(score.set_centres_device(q, fits.h_problems[q].mu));
            DPR_TRY(score.prepare());
            DPR_TRY(score.pass(kPassAssign, false));
            DPR_TRY(launch_ari_from_tables(cells, ds->n, kk, tables.as<unsigned long long>(), out.as<double>(), ctx().stream));
            DPR_CUDA(cudaMemcpyAsync(ari.data(), out.p, sizeof(double) * cells, cudaMemcpyDeviceToHost, ctx().stream));
            DPR_CUDA(cudaStreamSynchronize(ctx().stream));
            scored = true;
        }
    }
    const size_t per_cell_bytes = 2 * sizeof(int32_t) * (size_t)ds->n_pad;
    const int batch_cells = (int)std::max<size_t>(1, std::min<size_t>((size_t)cells, ((size_t)4 << 30) / per_cell_bytes));
    DevBuf lab_full;
    if (!scored) DPR_TRY(lab_full.alloc(per_cell_bytes * (size_t)batch_cells));
    for (int c0 = 0; c0 < cells && !scored; c0 += batch_cells) {   // (Monte-Carlo ARI, or tables too wide for shared memory: labels to HBM, counted afterwards)
        const int nc = std::min(batch_cells, cells - c0);
        std::vector<ProblemSpec> sspecs(2 * nc);
        for (int q = 0; q < 2 * nc; ++q)
            sspecs[q] = ProblemSpec{ds->x, ds->tile_xx, ds->tile_xx + ds->n_pad / kSlotCells, ds->n, specs[2 * c0 + q].k, 0.0, 0, lab_full.as<int32_t>() + (size_t)q * ds->n_pad};
        Session score;
        DPR_TRY(score.init(sspecs, d, false));
        for (int q = 0; q < 2 * nc; ++q) DPR_TRY(score.set_centres_device(q, fits.h_problems[2 * c0 + q].mu));
        DPR_TRY(score.prepare());
        DPR_TRY(score.pass(kPassAssign, false));
        std::vector<PairJob> pairs(nc);
        for (int c = 0; c < nc; ++c) pairs[c] = PairJob{sspecs[2 * c].labels, sspecs[2 * c + 1].labels, ds->n, (uint32_t)(c_lo + c0 + c)};
        DPR_TRY(run_ari(pairs, (uint32_t)kmax, seed, ari.data() + c0));
    }
    const auto t_score = now();
    if (timing)
        fprintf(stderr, "[dpr] grid_search %d problems: bootstrap %.2f ms, seeding %.2f, fits %.2f, scoring + ARI %.2f\n", cells, ms_since(t_start, t_boot),
                ms_since(t_boot, t_seed), ms_since(t_seed, t_fit), ms_since(t_fit, t_score));
    // 5. per-problem results: its ARI and the mean number of used clusters of its two fits.  The centres of all fits come back in ONE copy
    //    (the session keeps them back to back, k_max x d doubles per fit): a copy and a synchronisation per fit cost 4 ms per 220-fit set
    std::vector<double> all_mu;
    const size_t mu_stride = (size_t)fits.k_max() * d;
    if (centers_out) {
        all_mu.resize(mu_stride * (size_t)F);
        DPR_CUDA(cudaMemcpyAsync(all_mu.data(), fits.h_problems[0].mu, sizeof(double) * all_mu.size(), cudaMemcpyDeviceToHost, ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    size_t coff = 0;
    for (int c = 0; c < cells; ++c) {
        const int j = ((c_lo + c) / nreg) % nk;
        ari_out[c] = ari[c];
        int used = 0;
        for (int w = 0; w < 2; ++w) {
            for (int q = 0; q < k[j]; ++q) used += counts[(size_t)(2 * c + w) * fits.k_max() + q] > 0;
            if (centers_out) {
                const double* src = all_mu.data() + mu_stride * (size_t)(2 * c + w);
                std::vector<double> rm(src, src + (size_t)k[j] * d);
                rowmajor_to_colmajor(rm, k[j], d, centers_out + coff);
                coff += (size_t)k[j] * d;
            }
        }
        nclust_out[c] = used / 2.0;
    }
    return DPR_OK;
}

// the whole grid; ari_each / nclust_each: `iterations` matrices nk x nreg col-major
static int grid_search_impl(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg, uint32_t iterations,
                            uint32_t boot_n, uint64_t seed, double* ari_each, double* nclust_each, double* centers_out) {
    if (!ari_each || !nclust_each || nk < 1 || nreg < 1 || iterations < 1) return fail(DPR_ERR_INVALID, "grid_search: bad arguments");
    const int cells = (int)iterations * nk * nreg;
    std::vector<double> a(cells), c(cells);
    DPR_TRY(grid_search_range(ds, k, nk, reg, nreg, iterations, boot_n, seed, 0, cells, a.data(), c.data(), centers_out));
    for (int q = 0; q < cells; ++q) {
        const int l = q % nreg, j = (q / nreg) % nk, it = q / (nreg * nk);
        const size_t o = (size_t)it * nk * nreg + (size_t)l * nk + j;
        ari_each[o] = a[q];
        nclust_each[o] = c[q];
    }
    return DPR_OK;
}

namespace dpr {
int grid_search_cells(dpr_dataset* ds, int32_t k, const double* reg, int32_t nreg, uint32_t iterations, uint32_t boot_n, uint64_t seed, int c_lo,
                      int c_hi, double* ari_out, double* nclust_out, double* centers_out) {
    return grid_search_range(ds, &k, 1, reg, nreg, iterations, boot_n, seed, c_lo, c_hi, ari_out, nclust_out, centers_out);
}
int comm_allreduce_host_f64(std::initializer_list<std::vector<double>*> bufs) {
    if (ctx().world <= 1) return DPR_OK;
    size_t total = 0;
    for (auto* b : bufs) total += b->size();
    DevBuf dev;
    DPR_TRY(dev.alloc(sizeof(double) * total));
    size_t off = 0;
    for (auto* b : bufs) {
        DPR_CUDA(cudaMemcpyAsync(dev.as<double>() + off, b->data(), sizeof(double) * b->size(), cudaMemcpyHostToDevice, ctx().stream));
        off += b->size();
    }
    DPR_TRY(comm_allreduce_sum_f64(dev.as<double>(), total));
    off = 0;
    for (auto* b : bufs) {
        DPR_CUDA(cudaMemcpyAsync(b->data(), dev.as<double>() + off, sizeof(double) * b->size(), cudaMemcpyDeviceToHost, ctx().stream));
        off += b->size();
    }
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}
}  // namespace dpr

extern "C" {

// reference-shaped result: the mean over iterations (distances/iterations, found_cluster/(iterations*2), :387-390)
int dpr_dataset_grid_search(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg, uint32_t iterations,
                            uint32_t boot_n, uint64_t seed, double* ari_out, double* nclust_out, double* centers_out) {
    if (!ari_out || !nclust_out || nk < 1 || nreg < 1 || iterations < 1) return fail(DPR_ERR_INVALID, "grid_search: bad arguments");
    const size_t m = (size_t)nk * nreg;
    std::vector<double> a(m * iterations), c(m * iterations);
    DPR_TRY(grid_search_impl(ds, k, nk, reg, nreg, iterations, boot_n, seed, a.data(), c.data(), centers_out));
    for (size_t e = 0; e < m; ++e) {
        double sa = 0.0, sc = 0.0;
        for (uint32_t it = 0; it < iterations; ++it) {
            sa += a[it * m + e];
            sc += 2.0 * c[it * m + e];
        }
        ari_out[e] = sa / iterations;
        nclust_out[e] = sc / (iterations * 2);
    }
    return DPR_OK;
}
// one result per iteration: what the reference obtains from `iterations` separate worker processes that each
// call grid_search(..., iterations = 1, ...) (R/depechePenaltyOpt.R:50-55), computed as one batch
int dpr_dataset_grid_search_each(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg, uint32_t iterations,
                                 uint32_t boot_n, uint64_t seed, double* ari_each, double* nclust_each, double* centers_out) {
    return grid_search_impl(ds, k, nk, reg, nreg, iterations, boot_n, seed, ari_each, nclust_each, centers_out);
}
// a contiguous range of the grid's problems, numbered c = (it * nk + j) * nreg + l like the reference's loops (:349-352)
int dpr_dataset_grid_search_range(dpr_dataset_t ds, const int32_t* k, int32_t nk, const double* reg, int32_t nreg, uint32_t iterations,
                                  uint32_t boot_n, uint64_t seed, int32_t first_problem, int32_t n_problems, double* ari_out, double* nclust_out,
                                  double* centers_out) {
    return grid_search_range(ds, k, nk, reg, nreg, iterations, boot_n, seed, first_problem, first_problem + n_problems, ari_out, nclust_out, centers_out);
}
int dpr_grid_search(const double* X, int64_t n, int32_t d, const int32_t* k, int32_t nk, const double* reg, int32_t nreg,
                    uint32_t iterations, uint32_t boot_n, uint64_t seed, double* ari_out, double* nclust_out, double* centers_out) {
    dpr_dataset_t ds = nullptr;
    DPR_TRY(dpr_dataset_create(X, n, d, &ds));
    const int rc = dpr_dataset_grid_search(ds, k, nk, reg, nreg, iterations, boot_n, seed, ari_out, nclust_out, centers_out);
    dpr_dataset_destroy(ds);
    return rc;
}

// ---- allocate to many centre sets + all-pairs ARI (depecheClusterCenterSelection.R:11-29) ---------------------
int dpr_dataset_allocate_many(dpr_dataset_t ds, const double* mus, int32_t m, int32_t k, int32_t no_zero, int32_t* labels_out,
                              double* ari_out) {
    if (!ds || !mus || m < 1 || k < 1) return fail(DPR_ERR_INVALID, "allocate_many: bad arguments");
    DPR_TRY(ensure_ready());
    const int d = ds->d;
    DevBuf lab;
    DPR_TRY(lab.alloc(sizeof(int32_t) * (size_t)m * ds->n_pad));
    std::vector<ProblemSpec> specs(m);
    for (int s = 0; s < m; ++s) specs[s] = ProblemSpec{ds->x, ds->tile_xx, ds->tile_xx + ds->n_pad / kSlotCells, ds->n, k, 0.0, no_zero, lab.as<int32_t>() + (size_t)s * ds->n_pad};
    Session sess;
    DPR_TRY(sess.init(specs, d, false));
    std::vector<double> rm;
    for (int s = 0; s < m; ++s) {
        colmajor_to_rowmajor(mus + (size_t)s * k * d, k, d, rm);
        DPR_CUDA(cudaMemcpyAsync(sess.h_problems[s].mu, rm.data(), sizeof(double) * rm.size(), cudaMemcpyHostToDevice, ctx().stream));
        DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    }
    DPR_TRY(sess.prepare());
    DPR_TRY(sess.pass(kPassAssign, false));
    if (labels_out)
        for (int s = 0; s < m; ++s) DPR_TRY(copy_labels_out(specs[s].labels, ds->n, labels_out + (size_t)s * ds->n));
    if (ari_out) {
        // dAllocate returns 1-based labels and rand_index gets k (R/dAllocate.R:134, depecheClusterCenterSelection.R:26-29);
        // the partition is the same with 0-based labels, and the (k+1)-wide table holds both.
        std::vector<PairJob> pairs;
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) pairs.push_back(PairJob{specs[j].labels, specs[i].labels, ds->n, (uint32_t)(i * m + j)});
        std::vector<double> vals(pairs.size());
        DPR_TRY(run_ari(pairs, (uint32_t)k, 0, vals.data()));
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) ari_out[(size_t)i * m + j] = vals[(size_t)i * m + j];
    }
    return DPR_OK;
}

// ---- device-resident iterations for the bench ----------------------------------------------------------------------

static int bench_session(dpr_dataset_t ds, const double* mu, uint32_t k, double reg, int no_zero, bool sums) {
    if (!g_bench_session || g_bench_ds != ds || g_bench_k != (int)k || g_bench_mode != (int)sums) {
        g_bench_session.reset(new Session());
        g_bench_session->row_sharded = true;   // several ranks: the iterations of one fit over all ranks' shards
        DPR_TRY(single_session(ds, (int)k, reg, no_zero, sums, *g_bench_session));
        g_bench_ds = ds;
        g_bench_k = (int)k;
        g_bench_mode = (int)sums;
    }
    std::vector<double> rm;
    colmajor_to_rowmajor(mu, (int)k, ds->d, rm);
    DPR_TRY(g_bench_session->set_centres(0, rm.data()));
    Problem& pr = g_bench_session->h_problems[0];
    pr.reg = reg;
    pr.no_zero = no_zero;
    pr.done = 0;
    pr.delta = 0;
    g_bench_session->mark_restart();
    DPR_CUDA(cudaMemcpyAsync(g_bench_session->d_problems, &pr, sizeof(Problem), cudaMemcpyHostToDevice, ctx().stream));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return g_bench_session->prepare();
}

int dpr_dataset_lloyd_iterations(dpr_dataset_t ds, double* mu_inout, uint32_t k, double reg, int32_t no_zero, int32_t ramp_i,
                                 int32_t iters, int64_t* changed_out) {
    if (!ds || !mu_inout || k < 2 || iters < 1) return fail(DPR_ERR_INVALID, "lloyd_iterations: bad arguments");
    DPR_TRY(bench_session(ds, mu_inout, k, reg, no_zero, true));
    Session& s = *g_bench_session;
    if (ramp_i < 0) DPR_TRY(s.zero_labels());   // a fresh fit: previous labels are all 0 (Clusterer.cpp:226)
    for (int i = 0; i < iters; ++i) {
        DPR_TRY(s.pass(kPassLloyd, true));
        // iteration index <= 20 keeps the convergence exit off; ramp_i < 0 walks the ramp 0, 1, 2, ... of a real fit
        DPR_TRY(s.reduce_and_finalize(std::min(ramp_i < 0 ? i : ramp_i, 20), 1 << 30));
    }
    DPR_TRY(s.fetch_state());
    if (changed_out) *changed_out = s.h_problems[0].changed;
    std::vector<double> rm((size_t)k * ds->d);
    DPR_TRY(s.get_centres(0, rm.data()));
    rowmajor_to_colmajor(rm, (int)k, ds->d, mu_inout);
    return DPR_OK;
}
int dpr_dataset_assign_iterations(dpr_dataset_t ds, const double* mu, uint32_t k, int32_t no_zero, int32_t iters) {
    if (!ds || !mu || k < 1 || iters < 1) return fail(DPR_ERR_INVALID, "assign_iterations: bad arguments");
    DPR_TRY(bench_session(ds, mu, k, 0.0, no_zero, false));
    for (int i = 0; i < iters; ++i) DPR_TRY(g_bench_session->pass(kPassAssign, false));
    DPR_CUDA(cudaStreamSynchronize(ctx().stream));
    return DPR_OK;
}

}  // extern "C"
