// common.cuh -- shared runtime plumbing of libgxb200: context, error handling,
// stream-ordered device buffers, launch accounting, small device helpers.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

#include "gxb200.h"

namespace gx {

// ----------------------------------------------------------------------------- errors
struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};

void set_last_error(const std::string &msg);

#define GX_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess)                                                                \
            throw ::gx::Error(e__ == cudaErrorMemoryAllocation ? GX_ERR_OOM : GX_ERR_CUDA,     \
                              std::string(#call) + ": " + cudaGetErrorString(e__) + " (" +     \
                                  __FILE__ + ":" + std::to_string(__LINE__) + ")");            \
    } while (0)

#define GX_REQUIRE(cond, msg)                                                                  \
    do {                                                                                       \
        if (!(cond)) throw ::gx::Error(GX_ERR_INVALID, std::string(msg));                      \
    } while (0)

// Every extern "C" entry point runs its body through this guard.
template <class F>
int guarded(F &&f)
{
    try {
        f();
        return GX_OK;
    } catch (const Error &e) {
        set_last_error(e.what());
        return e.code;
    } catch (const std::bad_alloc &) {
        set_last_error("host allocation failed");
        return GX_ERR_OOM;
    } catch (const std::exception &e) {
        set_last_error(e.what());
        return GX_ERR_CUDA;
    }
}

// ----------------------------------------------------------------------------- context
struct Context {
    bool ready = false;
    int device = -1;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_a = nullptr, ev_b = nullptr;      // per-phase timing inside calls
    cudaEvent_t sw_start = nullptr, sw_stop = nullptr; // gx_timer_*
    gx_timing timing{};
    void *flush_buf = nullptr;
    size_t flush_bytes = 0;
    // multi-GPU
    int rank = 0, nranks = 1;
    void *nccl_comm = nullptr;
    void *pinned_scratch = nullptr; // small pinned area for flag/count read-backs
};

Context &ctx();
void require_ready();

inline void count_launch(unsigned k = 1) { ctx().timing.kernel_launches += k; }

// Optional per-kernel timing (gx_profile): a CUDA event pair around every launch on the
// library stream, aggregated by kernel name.  Off by default (it costs two event records per
// launch); bench.py switches it on for a separate pass to get the dominant kernel's duration.
bool profiling();
void prof_begin(const char *name);
void prof_end();

// Launch wrapper: counts the launch and checks the launch status.
#define GX_LAUNCH(kernel, grid, block, smem, ...)                                              \
    do {                                                                                       \
        const bool prof__ = ::gx::profiling();                                                 \
        if (prof__) ::gx::prof_begin(#kernel);                                                 \
        kernel<<<(grid), (block), (smem), ::gx::ctx().stream>>>(__VA_ARGS__);                  \
        if (prof__) ::gx::prof_end();                                                          \
        ::gx::count_launch();                                                                  \
        GX_CUDA(cudaGetLastError());                                                           \
    } while (0)

// ----------------------------------------------------------------------------- device memory
// Stream-ordered allocation from the device's default pool (kept warm between
// calls by a high release threshold, see gx_init).
template <class T>
struct DevBuf {
    T *p = nullptr;
    size_t n = 0;
    DevBuf() = default;
    explicit DevBuf(size_t count) { alloc(count); }
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
    DevBuf &operator=(DevBuf &&o) noexcept
    {
        if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
        return *this;
    }
    ~DevBuf() { release(); }
    void alloc(size_t count)
    {
        release();
        n = count;
        if (count) GX_CUDA(cudaMallocAsync((void **)&p, count * sizeof(T), ctx().stream));
    }
    void release()
    {
        if (p) cudaFreeAsync(p, ctx().stream);
        p = nullptr;
        n = 0;
    }
    void zero() { if (n) GX_CUDA(cudaMemsetAsync(p, 0, n * sizeof(T), ctx().stream)); }
    void fill_byte(int b) { if (n) GX_CUDA(cudaMemsetAsync(p, b, n * sizeof(T), ctx().stream)); }
    size_t bytes() const { return n * sizeof(T); }
    operator T *() const { return p; }
};

// Timing of a phase with the context's event pair (synchronises the stream).
struct PhaseTimer {
    double *slot;
    explicit PhaseTimer(double *s) : slot(s) { GX_CUDA(cudaEventRecord(ctx().ev_a, ctx().stream)); }
    void stop()
    {
        if (!slot) return;
        GX_CUDA(cudaEventRecord(ctx().ev_b, ctx().stream));
        GX_CUDA(cudaEventSynchronize(ctx().ev_b));
        float ms = 0;
        GX_CUDA(cudaEventElapsedTime(&ms, ctx().ev_a, ctx().ev_b));
        *slot += ms;
        slot = nullptr;
    }
    ~PhaseTimer()
    {
        try { stop(); } catch (...) {}
    }
};

inline unsigned grid_for(uint64_t items, unsigned block, unsigned per_thread = 1)
{
    uint64_t per_block = (uint64_t)block * per_thread;
    uint64_t g = (items + per_block - 1) / per_block;
    if (g < 1) g = 1;
    if (g > 0x7FFFFFFFull) g = 0x7FFFFFFFull;
    return (unsigned)g;
}

// grid sized as a multiple of the SM count for grid-stride kernels
inline unsigned grid_persistent(unsigned ctas_per_sm) { return (unsigned)ctx().num_sms * ctas_per_sm; }

// read `count` 8-byte words back through the pinned scratch (synchronises)
void read_back(void *host_dst, const void *dev_src, size_t bytes);

// ----------------------------------------------------------------------------- device helpers
#ifdef __CUDACC__
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }

template <class T>
__device__ __forceinline__ T warp_sum(T v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
    return v;
}

__device__ __forceinline__ uint32_t warp_min_u32(uint32_t v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(FULL, v, o));
    return v;
}

// streaming (read-once) loads: keep the adjacency stream out of L1
__device__ __forceinline__ uint32_t ld_stream(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint4 ld_stream4(const uint4 *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
// 8 consecutive ids in one 256-bit load (sm_100), marked evict-first in L2 as well: a read-once
// stream must not push the gathered vectors out of L2.  p must be 32-byte aligned.
__device__ __forceinline__ void ld_stream8(const uint32_t *p, uint32_t (&v)[8])
{
    asm volatile("ld.global.nc.L1::no_allocate.L2::evict_first.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ double ld_stream_f64(const double *p)
{
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
#endif

} // namespace gx
