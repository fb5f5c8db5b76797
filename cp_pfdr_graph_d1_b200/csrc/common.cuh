// common.cuh - shared plumbing of libpfdr_b200: error handling, device buffers, launch
// accounting, the device-side control block of a PFDR solve and deterministic reductions.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>
#include <cfloat>
#include <cmath>
#include <atomic>
#include <mutex>
#include <thread>
#include "../../include/pfdr_b200.h"

namespace pfdr {

// ---------------------------------------------------------------------------------------------
// errors: internal code throws, the C ABI catches and turns into status + message
// ---------------------------------------------------------------------------------------------
struct Error : std::runtime_error {
    int status;
    Error(int s, const std::string &m) : std::runtime_error(m), status(s) {}
};

#define PFDR_CUDA(call)                                                                        \
    do {                                                                                       \
        cudaError_t err__ = (call);                                                            \
        if (err__ != cudaSuccess)                                                              \
            throw ::pfdr::Error(PFDR_B200_ERR_CUDA, std::string(#call) + " failed: " +         \
                                cudaGetErrorString(err__) + " (" __FILE__ ":" +                \
                                std::to_string(__LINE__) + ")");                               \
    } while (0)

#define PFDR_REQUIRE(cond, msg)                                                                \
    do { if (!(cond)) throw ::pfdr::Error(PFDR_B200_ERR_ARG, std::string(msg)); } while (0)

// ---------------------------------------------------------------------------------------------
// per-thread run-time context: device, stream, statistics of the call in flight
// ---------------------------------------------------------------------------------------------
struct Context {
    int device = -1;               // -1: keep the caller's current device
    cudaStream_t stream = 0;
    pfdr_b200_stats stats{};
    bool in_loop = false;
    bool user_stream = false;      // the stream was given by pfdr_b200_set_stream (else: library-owned, non-blocking)
    bool ordered = false;          // this call has already been ordered after the caller's default-stream work
    void count_launch() { stats.launches++; if (in_loop) stats.loop_launches++; }
};
Context &ctx();

// Pinned control-block mirrors and events are recycled per host thread (and device): creating them costs a
// large part of a millisecond per solve, and cut-pursuit calls the one-shot entry points once per outer iteration.
struct HostCache {
    int device = -1;
    std::vector<void *> pinned;        // blocks of PINNED_BLOCK bytes
    std::vector<cudaEvent_t> events;
};
constexpr size_t PINNED_BLOCK = 512;
inline HostCache &host_cache()
{
    static thread_local HostCache c;
    int dev = 0;
    cudaGetDevice(&dev);
    if (c.device != dev) {             // different device: drop what belonged to the other one
        for (void *q : c.pinned) cudaFreeHost(q);
        for (cudaEvent_t e : c.events) cudaEventDestroy(e);
        c.pinned.clear(); c.events.clear(); c.device = dev;
    }
    return c;
}
inline void *pinned_get()
{
    HostCache &c = host_cache();
    void *q = nullptr;
    if (!c.pinned.empty()) { q = c.pinned.back(); c.pinned.pop_back(); return q; }
    PFDR_CUDA(cudaMallocHost(&q, PINNED_BLOCK));
    return q;
}
inline void pinned_put(void *q) { if (q) host_cache().pinned.push_back(q); }
inline cudaEvent_t event_get()
{
    HostCache &c = host_cache();
    cudaEvent_t e = nullptr;
    if (!c.events.empty()) { e = c.events.back(); c.events.pop_back(); return e; }
    PFDR_CUDA(cudaEventCreate(&e));
    return e;
}
inline void event_put(cudaEvent_t e) { if (e) host_cache().events.push_back(e); }

#define PFDR_LAUNCH(kernel, grid, block, smem, ...)                                            \
    do {                                                                                       \
        kernel<<<(grid), (block), (smem), ::pfdr::ctx().stream>>>(__VA_ARGS__);                \
        ::pfdr::ctx().count_launch();                                                          \
        PFDR_CUDA(cudaGetLastError());                                                         \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device memory
// ---------------------------------------------------------------------------------------------
// Device memory comes from the stream-ordered allocator (cudaMallocAsync on the launch stream) with the pool's
// release threshold lifted, so the ~25 arrays of a solve cost microseconds after the first call instead of a
// cudaMalloc/cudaFree pair each (tens of milliseconds per solve: most of the one-shot entry points' overhead,
// and they are called once per cut-pursuit iteration by the reference's outer loop).
inline bool use_pool()
{
    static bool lifted[64] = {};                     // per device (several host threads may drive one GPU each)
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !lifted[dev]) {
        int ok = 0;
        cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev);
        cudaMemPool_t pool;
        if (ok && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long thr = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
            lifted[dev] = true;
        } else { cudaGetLastError(); return false; }
        cudaGetLastError();
    }
    return true;
}

template <typename T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    bool pooled = false, borrowed = false;
    cudaStream_t home = nullptr;       // stream the block was allocated on: freed on the same one, whoever destroys the buffer
    DevBuf() {}
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n), pooled(o.pooled), borrowed(o.borrowed), home(o.home) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; pooled = o.pooled; borrowed = o.borrowed; home = o.home; o.p = nullptr; o.n = 0; }
        return *this;
    }
    void borrow(T *q, size_t count) { release(); p = q; n = count; borrowed = true; }     // memory owned by someone else (a slab arena)
    ~DevBuf() { release(); }
    void alloc(size_t count) {
        release();
        n = count;
        if (!count) return;
        pooled = use_pool();
        home = ctx().stream;
        if (pooled) PFDR_CUDA(cudaMallocAsync((void **) &p, count * sizeof(T), home));
        else PFDR_CUDA(cudaMalloc((void **) &p, count * sizeof(T)));
    }
    void release() {
        if (p && !borrowed) { if (pooled) cudaFreeAsync(p, home); else cudaFree(p); }
        p = nullptr; n = 0; borrowed = false;
    }
    void zero() { if (p) PFDR_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx().stream)); }
    operator T *() const { return p; }
};

inline bool is_device_ptr(const void *q);

// A device pointer handed in by the caller may still be written by work the caller enqueued.  The library runs on its
// own non-blocking stream unless pfdr_b200_set_stream gave it the caller's: order that stream after everything
// enqueued so far on the legacy default stream (which itself waits for all blocking streams) - the stream a
// torch / CUDA runtime user produces data on unless they say otherwise.  (pfdr_b200.h: producers on other
// non-blocking streams must pass that stream with pfdr_b200_set_stream or synchronise it.)
inline void order_after_caller()
{
    Context &c = ctx();
    if (c.user_stream || c.ordered) return;
    c.ordered = true;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(c.stream, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) { cudaGetLastError(); return; }
    cudaEvent_t e = event_get();
    PFDR_CUDA(cudaEventRecord(e, cudaStreamLegacy));
    PFDR_CUDA(cudaStreamWaitEvent(c.stream, e, 0));
    event_put(e);
}

inline bool is_device_ptr(const void *q)
{
    if (!q) return false;
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, q);
    if (e != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

// ---------------------------------------------------------------------------------------------
// Large transfers from / to ORDINARY (pageable) host memory - what the MEX gateways and the cut-pursuit hand over.
// cudaMemcpyAsync on pageable memory stages through one driver buffer at 10-12 GB/s; here a few host threads copy
// chunks into / out of their own pinned blocks while the DMA engine moves the previous ones, which brings the C4 upload
// (2.6 GB) and download (1 GB) close to the pinned rate.  Pinned and registered memory take the plain path.
// ---------------------------------------------------------------------------------------------
constexpr size_t STAGE_CHUNK = 4u << 20;
constexpr int STAGE_THREADS = 4;

struct StagePool {
    std::mutex m;
    std::vector<void *> blocks;
    void *get()
    {
        { std::lock_guard<std::mutex> g(m); if (!blocks.empty()) { void *q = blocks.back(); blocks.pop_back(); return q; } }
        void *q = nullptr;
        return cudaMallocHost(&q, STAGE_CHUNK) == cudaSuccess ? q : nullptr;
    }
    void put(void *q)
    {
        if (!q) return;
        std::lock_guard<std::mutex> g(m);
        if (blocks.size() < 4 * STAGE_THREADS) blocks.push_back(q); else cudaFreeHost(q);
    }
};
inline StagePool &stage_pool() { static StagePool sp; return sp; }

inline size_t stage_min_bytes()
{
    const char *e = getenv("PFDR_B200_STAGE_MIN");           // debug switch (tests): stage transfers of at least this many bytes
    return e ? (size_t) strtoull(e, nullptr, 10) : ((size_t) 64 << 20);
}
inline bool is_pageable_host(const void *q)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, q) != cudaSuccess) { cudaGetLastError(); return true; }
    return at.type == cudaMemoryTypeUnregistered;
}

// dev <- host (to_device) or host <- dev, `bytes` bytes, enqueued on st; returns when every pinned block is free again
// (to_device: the data is on the device or in flight behind st; from device: the data is in `host`)
inline void staged_copy(void *dev, void *host, size_t bytes, bool to_device, cudaStream_t st)
{
    int device = 0;
    PFDR_CUDA(cudaGetDevice(&device));
    const size_t nch = (bytes + STAGE_CHUNK - 1) / STAGE_CHUNK;
    const int T = (int) std::min<size_t>(STAGE_THREADS, nch);
    std::atomic<int> failed{0};
    auto work = [&](int t) {
        if (cudaSetDevice(device) != cudaSuccess) { failed = 1; return; }
        void *buf[2] = {stage_pool().get(), stage_pool().get()};
        cudaEvent_t ev[2] = {nullptr, nullptr};
        bool ok = buf[0] && buf[1] && cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming) == cudaSuccess &&
                  cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming) == cudaSuccess;
        size_t pend_off[2] = {0, 0}, pend_n[2] = {0, 0};
        bool busy[2] = {false, false};
        auto settle = [&](int k) {             // the transfer using block k is over; from device: hand its bytes to the caller
            if (!busy[k]) return;
            if (cudaEventSynchronize(ev[k]) != cudaSuccess) ok = false;
            if (!to_device && ok) memcpy((char *) host + pend_off[k], buf[k], pend_n[k]);
            busy[k] = false;
        };
        int k = 0;
        for (size_t c = (size_t) t; c < nch && ok; c += (size_t) T, k ^= 1) {
            const size_t off = c * STAGE_CHUNK, n = std::min(STAGE_CHUNK, bytes - off);
            settle(k);
            if (!ok) break;
            cudaError_t rc;
            if (to_device) {
                memcpy(buf[k], (const char *) host + off, n);
                rc = cudaMemcpyAsync((char *) dev + off, buf[k], n, cudaMemcpyHostToDevice, st);
            } else {
                rc = cudaMemcpyAsync(buf[k], (const char *) dev + off, n, cudaMemcpyDeviceToHost, st);
            }
            if (rc != cudaSuccess || cudaEventRecord(ev[k], st) != cudaSuccess) { ok = false; break; }
            pend_off[k] = off; pend_n[k] = n; busy[k] = true;
        }
        settle(0); settle(1);
        if (!ok) failed = 1;
        for (int i = 0; i < 2; i++) { if (ev[i]) cudaEventDestroy(ev[i]); stage_pool().put(buf[i]); }
    };
    std::vector<std::thread> th;
    for (int t = 1; t < T; t++) th.emplace_back(work, t);
    work(0);
    for (auto &x : th) x.join();
    if (failed) { cudaGetLastError(); throw Error(PFDR_B200_ERR_CUDA, "staged host transfer failed"); }
}

// A read-only input array that lives on the device for the duration of a call: either the
// caller's own device pointer (used in place) or an uploaded copy of the caller's host array.
template <typename T>
struct DevIn {
    const T *p = nullptr;
    DevBuf<T> own;
    void bind(const T *src, size_t count) {
        if (!src || !count) { p = nullptr; return; }
        if (is_device_ptr(src)) { order_after_caller(); p = src; return; }
        own.alloc(count);
        if (count * sizeof(T) >= stage_min_bytes() && is_pageable_host(src))
            staged_copy(own.p, const_cast<T *>(src), count * sizeof(T), true, ctx().stream);
        else
            PFDR_CUDA(cudaMemcpyAsync(own.p, src, count * sizeof(T), cudaMemcpyHostToDevice, ctx().stream));
        ctx().stats.h2d_bytes += (long long) (count * sizeof(T));
        p = own.p;
    }
    operator const T *() const { return p; }
};

template <typename T>
inline void copy_out(T *dst, const T *dev_src, size_t count)
{
    if (!dst || !count) return;
    if (is_device_ptr(dst)) {
        PFDR_CUDA(cudaMemcpyAsync(dst, dev_src, count * sizeof(T), cudaMemcpyDeviceToDevice, ctx().stream));
    } else {
        if (count * sizeof(T) >= stage_min_bytes() && is_pageable_host(dst))
            staged_copy(const_cast<T *>(dev_src), dst, count * sizeof(T), false, ctx().stream);
        else
            PFDR_CUDA(cudaMemcpyAsync(dst, dev_src, count * sizeof(T), cudaMemcpyDeviceToHost, ctx().stream));
        ctx().stats.d2h_bytes += (long long) (count * sizeof(T));
    }
}

inline int num_sms()
{
    static thread_local int sms = 0, of_dev = -1;      // per host thread and device (a thread may drive one GPU of several)
    int dev; PFDR_CUDA(cudaGetDevice(&dev));
    if (dev != of_dev) {
        PFDR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        of_dev = dev;
    }
    return sms;
}

inline unsigned cdiv(size_t a, size_t b) { return (unsigned) ((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------------
// control block of one solve, resident in device memory.  Every kernel of an iteration reads it,
// so the host can enqueue whole batches of iterations without synchronising: once the stopping
// (or reconditioning) test fires on the device, the remaining launches of the batch are no-ops.
// ---------------------------------------------------------------------------------------------
enum : int { ST_RUN = 0, ST_FINAL = 1, ST_RECOND = 2 };

struct Ctl {
    int it;            // completed iterations == index of the trip in flight; parity selects buffers
    int state;         // ST_RUN | ST_FINAL (stop test fired: this trip only applies A / objective)
                       //        | ST_RECOND (this trip applies A / objective, then the host reconditions)
    int hold;          // set by the first iteration kernel that sees state != ST_RUN: freezes everything
    int itMax;
    int track;         // evolution is monitored (difTol > 0 || difRcd > 0 || Dif != NULL)
    unsigned ticket;   // last-block-done counter
    int bdone;         // slab-sharded solves: boundary chunks completed since reset() (slab.cuh)
    int pad1;
    double dif;        // current evolution measure (as the reference's `dif`)
    double tol;        // stop threshold in the units of dif  (quadratic: difTol^2, simplex: difTol)
    double rcd;        // reconditioning threshold, same units
    double eps;        // floor of the denominator (quadratic family)
    double amp;        // preconditioner: mean-amplitude constant `c` of the reference
    double scratch[4];
};

__device__ __forceinline__ int ld_flag(const int *p) { return *((const volatile int *) p); }

// One thread per CTA reads the control block and broadcasts it through shared memory: a million
// threads polling the same L2 sector cost ~20 us per launch (measured, profiles/r01_notes.md).
struct CtlView { int hold, state, it; };
__device__ __forceinline__ CtlView load_ctl(const Ctl *ctl)
{
    __shared__ int s_ctl[3];
    if (threadIdx.x == 0) { s_ctl[0] = ld_flag(&ctl->hold); s_ctl[1] = ld_flag(&ctl->state); s_ctl[2] = ld_flag(&ctl->it); }
    __syncthreads();
    CtlView v; v.hold = s_ctl[0]; v.state = s_ctl[1]; v.it = s_ctl[2];
    return v;
}

// ---------------------------------------------------------------------------------------------
// deterministic block reduction (fixed shuffle tree + fixed order over warps)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// all threads of the block must call; result valid in thread 0 (and broadcast to all)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* >= NV*32 doubles */)
{
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++) v[i] = warp_sum(v[i]);
    __syncthreads();
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) smem[i * 32 + wid] = v[i];
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; i++) {
        double t = (lane < nw) ? smem[i * 32 + lane] : 0.0;
        v[i] = warp_sum(t);
    }
}

// reciprocal for the recomputed edge constants: ~1 ulp (f32 MUFU.RCP; f64 MUFU.RCP64H + 2 Newton steps),
// instead of the ~10 / ~25 instruction IEEE division sequences
__device__ __forceinline__ float fast_rcp(float x)
{   // one MUFU.RCP (__fdividef(1, x) wraps it in five more instructions of range handling that the sums of
    // metric entries it is applied to never need)
    float r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ double fast_rcp(double x)
{
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    r = r * (2.0 - x * r);
    r = r * (2.0 - x * r);
    return r;
}

// Cheaper grid-wide deterministic sum for the iteration kernels: 32 values are added in `real` with one shuffle tree,
// the per-warp results go to shared memory as doubles and thread 0 alone adds them in warp order (no second
// shuffle tree in every warp).  The last CTA to arrive (ticket) adds the per-CTA partials in CTA order.
// smem: >= NV*32 doubles.  Returns true in every thread of the last CTA; the totals are valid in thread 0.
template <int NV, typename real>
__device__ __forceinline__ bool grid_sum_fast(Ctl *ctl, double *partial, real (&v)[NV], double (&tot)[NV], double *smem)
{
    __shared__ int s_last_fast;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; i++)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], o);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) smem[i * 32 + wid] = (double) v[i];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 0; i < NV; i++) {
            double a = 0.0;
            for (int w = 0; w < nw; w++) a += smem[i * 32 + w];
            partial[(size_t) blockIdx.x * NV + i] = a;
        }
        __threadfence();
        s_last_fast = (atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!s_last_fast) return false;
    __threadfence();
    double acc[NV];
#pragma unroll
    for (int i = 0; i < NV; i++) acc[i] = 0.0;
    const volatile double *vp = partial;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int i = 0; i < NV; i++) acc[i] += vp[(size_t) b * NV + i];
    }
    block_sum<NV>(acc, smem);
#pragma unroll
    for (int i = 0; i < NV; i++) tot[i] = acc[i];
    if (threadIdx.x == 0) ctl->ticket = 0;
    return true;
}

// `verbose` of the reference (src/PFDR_graph_quadratic_d1_l1.cpp:43-55, :425-428): a progress message every `verbose`
// iterations - iteration count, and the evolution measure with its thresholds when it is monitored.  The host sees the
// device's control block one batch of iterations behind, so the lines of a batch appear together; the evolution value
// of each reported iteration comes from the per-iteration record on the device (kept whenever verbose > 0).
struct ProgressPrinter {
    int every = 0, next = 0;
    cudaStream_t side = nullptr;
    ~ProgressPrinter() { if (side) cudaStreamDestroy(side); }
    void start(int verbose) { every = verbose > 0 ? verbose : 0; next = every; }
    template <typename real>
    void upto(int it_done, int itMax, const real *dDif, bool track, double tol, double rcd)
    {
        if (!every || it_done < next) return;
        const int first = next, count = (it_done - first) / every + 1;
        std::vector<real> vals;
        if (track && dDif) {          // completed iterations only: their record is final, whatever is still queued
            if (!side) PFDR_CUDA(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
            vals.resize((size_t) it_done - first + 1);
            PFDR_CUDA(cudaMemcpyAsync(vals.data(), dDif + (first - 1), vals.size() * sizeof(real), cudaMemcpyDeviceToHost, side));
            PFDR_CUDA(cudaStreamSynchronize(side));
        }
        for (int k = 0; k < count; k++) {
            const int it = first + k * every;
            printf("iteration %d (max. %d)\n", it, itMax);
            if (!vals.empty()) printf("iterate evolution %g (recond. %g; tol. %g)\n", (double) vals[(size_t) k * every], rcd, tol);
        }
        fflush(stdout);
        next = first + count * every;
    }
};

template <typename real> struct Pair;
template <> struct Pair<float> { typedef float2 type; };
template <> struct Pair<double> { typedef double2 type; };

template <typename real> __host__ __device__ inline real real_eps();
template <> __host__ __device__ inline float real_eps<float>() { return FLT_EPSILON; }
template <> __host__ __device__ inline double real_eps<double>() { return DBL_EPSILON; }

}  // namespace pfdr
