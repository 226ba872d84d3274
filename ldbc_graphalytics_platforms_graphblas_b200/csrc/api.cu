// api.cu -- context, error reporting, stopwatch, pinned memory, NCCL communicator.
#include <nccl.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "comm.cuh"

namespace gx {

static thread_local std::string g_last_error;
static Context g_ctx;

void set_last_error(const std::string &msg) { g_last_error = msg; }
Context &ctx() { return g_ctx; }

void require_ready()
{
    if (!g_ctx.ready)
        throw Error(GX_ERR_NO_DEVICE, "gx_init has not succeeded: no CUDA device bound (there is no CPU fallback)");
}

void read_back(void *host_dst, const void *dev_src, size_t bytes)
{
    Context &c = ctx();
    if (bytes > 4096) throw Error(GX_ERR_INVALID, "read_back is for small scalars only");
    GX_CUDA(cudaMemcpyAsync(c.pinned_scratch, dev_src, bytes, cudaMemcpyDeviceToHost, c.stream));
    GX_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(host_dst, c.pinned_scratch, bytes);
}

// ----------------------------------------------------------------------------- profiler
struct ProfAgg { double ms = 0; uint64_t count = 0; };
static bool g_prof_on = false;
static std::vector<cudaEvent_t> g_ev_pool;
static std::vector<std::pair<const char *, std::pair<cudaEvent_t, cudaEvent_t>>> g_pending;
static std::map<std::string, ProfAgg> g_agg;
static std::string g_prof_text;

bool profiling() { return g_prof_on; }

static cudaEvent_t prof_event()
{
    if (!g_ev_pool.empty()) { cudaEvent_t e = g_ev_pool.back(); g_ev_pool.pop_back(); return e; }
    cudaEvent_t e;
    GX_CUDA(cudaEventCreate(&e));
    return e;
}

void prof_begin(const char *name)
{
    cudaEvent_t a = prof_event(), b = prof_event();
    g_pending.push_back({name, {a, b}});
    GX_CUDA(cudaEventRecord(a, ctx().stream));
}

void prof_end() { GX_CUDA(cudaEventRecord(g_pending.back().second.second, ctx().stream)); }

static void prof_collect()
{
    if (g_pending.empty()) return;
    GX_CUDA(cudaStreamSynchronize(ctx().stream));
    for (auto &p : g_pending) {
        float ms = 0;
        GX_CUDA(cudaEventElapsedTime(&ms, p.second.first, p.second.second));
        ProfAgg &a = g_agg[p.first];
        a.ms += ms;
        a.count++;
        g_ev_pool.push_back(p.second.first);
        g_ev_pool.push_back(p.second.second);
    }
    g_pending.clear();
}

__global__ void k_flush(uint4 *__restrict__ buf, size_t n16, uint32_t tag)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; i < n16; i += stride) buf[i] = make_uint4(tag, tag, tag, tag);
}

#define GX_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            throw ::gx::Error(GX_ERR_CUDA, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

} // namespace gx

using namespace gx;

extern "C" const char *gx_last_error(void) { return g_last_error.c_str(); }

extern "C" int gx_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int gx_init(int device)
{
    return guarded([&] {
        Context &c = ctx();
        if (c.ready && c.device == device) return;
        if (c.ready) throw Error(GX_ERR_INVALID, "context already bound to another device");
        // one fresh process per job (GraphblasJob.java:70-97): load every kernel of the library now, with the context,
        // instead of lazily at its first launch inside an algorithm's timed window (has no effect once the CUDA
        // runtime is up, e.g. under a host that initialised it before; an explicit setting wins)
        setenv("CUDA_MODULE_LOADING", "EAGER", 0);
        int n = 0;
        cudaError_t e = cudaGetDeviceCount(&n);
        if (e != cudaSuccess || n == 0) {
            cudaGetLastError();
            throw Error(GX_ERR_NO_DEVICE, "no CUDA device available (there is no CPU fallback)");
        }
        GX_REQUIRE(device >= 0 && device < n, "device index out of range");
        GX_CUDA(cudaSetDevice(device));
        cudaDeviceProp prop;
        GX_CUDA(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            throw Error(GX_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major * 10 + prop.minor) +
                                              "; this library carries sm_100a code only");
        c.num_sms = prop.multiProcessorCount;
        GX_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
        GX_CUDA(cudaEventCreate(&c.ev_a));
        GX_CUDA(cudaEventCreate(&c.ev_b));
        GX_CUDA(cudaEventCreate(&c.sw_start));
        GX_CUDA(cudaEventCreate(&c.sw_stop));
        GX_CUDA(cudaMallocHost(&c.pinned_scratch, 4096));
        // keep freed blocks in the pool: repeated jobs reuse HBM instead of re-mapping it
        cudaMemPool_t pool;
        GX_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thresh = ~0ull;
        GX_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
        c.device = device;
        c.ready = true;
    });
}

// Grows the stream-ordered memory pool to `bytes` ahead of time (allocate + free; the release threshold keeps the
// memory in the pool), so that the algorithms' scratch buffers are served from the pool instead of being mapped on
// first use.  Resource allocation, not computation: the analogue of sizing a memory pool at start-up.
extern "C" int gx_reserve(uint64_t bytes)
{
    return guarded([&] {
        require_ready();
        Context &c = ctx();
        size_t free_b = 0, total_b = 0;
        GX_CUDA(cudaMemGetInfo(&free_b, &total_b));
        const uint64_t cap = (uint64_t)(free_b * 0.6);
        if (bytes > cap) bytes = cap;
        if (!bytes) return;
        void *p = nullptr;
        if (cudaMallocAsync(&p, bytes, c.stream) != cudaSuccess) { cudaGetLastError(); return; } // best effort
        GX_CUDA(cudaFreeAsync(p, c.stream));
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}

extern "C" int gx_finalize(void)
{
    return guarded([&] {
        Context &c = ctx();
        if (!c.ready) return;
        cudaStreamSynchronize(c.stream);
        peer_cache_clear();
        if (c.nccl_comm) { ncclCommDestroy((ncclComm_t)c.nccl_comm); c.nccl_comm = nullptr; }
        if (c.flush_buf) cudaFree(c.flush_buf);
        cudaFreeHost(c.pinned_scratch);
        cudaEventDestroy(c.ev_a);
        cudaEventDestroy(c.ev_b);
        cudaEventDestroy(c.sw_start);
        cudaEventDestroy(c.sw_stop);
        cudaStreamDestroy(c.stream);
        c = Context{};
    });
}

extern "C" int gx_last_timing(gx_timing *t)
{
    return guarded([&] {
        GX_REQUIRE(t != nullptr, "NULL argument");
        *t = ctx().timing;
    });
}

extern "C" int gx_timer_start(void)
{
    return guarded([&] {
        require_ready();
        GX_CUDA(cudaEventRecord(ctx().sw_start, ctx().stream));
    });
}

extern "C" int gx_timer_stop(double *elapsed_ms)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(elapsed_ms != nullptr, "NULL argument");
        GX_CUDA(cudaEventRecord(ctx().sw_stop, ctx().stream));
        GX_CUDA(cudaEventSynchronize(ctx().sw_stop));
        float ms = 0;
        GX_CUDA(cudaEventElapsedTime(&ms, ctx().sw_start, ctx().sw_stop));
        *elapsed_ms = ms;
    });
}

extern "C" int gx_sync(void)
{
    return guarded([&] {
        require_ready();
        GX_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}

extern "C" int gx_flush_l2(void)
{
    return guarded([&] {
        require_ready();
        Context &c = ctx();
        if (!c.flush_buf) {
            c.flush_bytes = 256ull << 20; // 2x the 126 MB L2
            GX_CUDA(cudaMalloc(&c.flush_buf, c.flush_bytes));
        }
        static uint32_t tag = 0;
        k_flush<<<grid_persistent(8), 256, 0, c.stream>>>((uint4 *)c.flush_buf, c.flush_bytes / 16, ++tag);
        GX_CUDA(cudaGetLastError());
    });
}

// gx_profile(1) starts a fresh per-kernel timing session, gx_profile(0) stops it.
extern "C" int gx_profile(int enable)
{
    return guarded([&] {
        require_ready();
        prof_collect();
        if (enable) g_agg.clear();
        g_prof_on = enable != 0;
    });
}

// "<kernel name>\t<launches>\t<total ms>\n" per kernel, sorted by total time (descending).
extern "C" const char *gx_profile_report(void)
{
    g_prof_text.clear();
    try {
        prof_collect();
    } catch (...) {
        return "";
    }
    std::vector<std::pair<std::string, ProfAgg>> v(g_agg.begin(), g_agg.end());
    std::sort(v.begin(), v.end(), [](const auto &x, const auto &y) { return x.second.ms > y.second.ms; });
    for (auto &e : v) g_prof_text += e.first + "\t" + std::to_string(e.second.count) + "\t" + std::to_string(e.second.ms) + "\n";
    return g_prof_text.c_str();
}

extern "C" int gx_host_alloc(void **p, uint64_t bytes)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(p != nullptr, "NULL argument");
        GX_CUDA(cudaMallocHost(p, bytes ? bytes : 1));
    });
}

extern "C" int gx_host_free(void *p)
{
    return guarded([&] {
        if (p) GX_CUDA(cudaFreeHost(p));
    });
}

extern "C" void gx_free_host(void *p) { free(p); }

// Page-locks a host array the caller already owns (e.g. the CSR arrays a loader filled), so that the upload that
// follows runs at PCIe speed without a pageable staging copy.  Best effort: a failure leaves the memory pageable.
extern "C" int gx_host_register(const void *p, uint64_t bytes)
{
    return guarded([&] {
        require_ready();
        if (!p || !bytes) return;
        if (cudaHostRegister(const_cast<void *>(p), bytes, cudaHostRegisterDefault) != cudaSuccess) cudaGetLastError();
    });
}

extern "C" int gx_host_unregister(const void *p)
{
    return guarded([&] {
        if (!p) return;
        if (cudaHostUnregister(const_cast<void *>(p)) != cudaSuccess) cudaGetLastError();
    });
}

// ------------------------------------------------------------------------- NCCL
extern "C" int gx_comm_unique_id(void *id128)
{
    return guarded([&] {
        GX_REQUIRE(id128 != nullptr, "NULL argument");
        static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
        ncclUniqueId id;
        GX_NCCL(ncclGetUniqueId(&id));
        memcpy(id128, &id, sizeof(id));
    });
}

extern "C" int gx_comm_init(int rank, int nranks, const void *id128)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(nranks >= 1 && rank >= 0 && rank < nranks, "bad rank / nranks");
        Context &c = ctx();
        if (c.nccl_comm) throw Error(GX_ERR_INVALID, "communicator already initialised");
        c.rank = rank;
        c.nranks = nranks;
        if (nranks == 1) return;
        GX_REQUIRE(id128 != nullptr, "NULL unique id");
        ncclUniqueId id;
        memcpy(&id, id128, sizeof(id));
        ncclComm_t comm;
        GX_NCCL(ncclCommInitRank(&comm, nranks, id, rank));
        c.nccl_comm = comm;
    });
}

extern "C" int gx_comm_destroy(void)
{
    return guarded([&] {
        Context &c = ctx();
        peer_cache_clear();
        if (c.nccl_comm) {
            cudaStreamSynchronize(c.stream);
            GX_NCCL(ncclCommDestroy((ncclComm_t)c.nccl_comm));
            c.nccl_comm = nullptr;
        }
        c.rank = 0;
        c.nranks = 1;
    });
}
