// comm.cu -- NCCL collectives on the library stream (see comm.cuh for the sharding model).
#include <nccl.h>

#include <chrono>
#include <cstring>
#include <vector>

#include "comm.cuh"

namespace gx {

#define GX_NCCL(call)                                                                          \
    do {                                                                                       \
        ncclResult_t r__ = (call);                                                             \
        if (r__ != ncclSuccess)                                                                \
            throw ::gx::Error(GX_ERR_CUDA, std::string(#call) + ": " + ncclGetErrorString(r__)); \
    } while (0)

// bounds[r] = first row whose entry offset reaches r * total / nranks, rounded down to 32
__global__ void k_partition(const uint64_t *__restrict__ rp0, const uint64_t *__restrict__ rp1, uint64_t n, int nranks,
                            uint64_t *__restrict__ bounds)
{
    const int r = threadIdx.x;
    if (r > nranks) return;
    if (r == 0) { bounds[0] = 0; return; }
    if (r == nranks) { bounds[r] = n; return; }
    const uint64_t total = rp0[n] + (rp1 ? rp1[n] : 0);
    const uint64_t target = (uint64_t)(((unsigned __int128)total * (unsigned)r) / (unsigned)nranks);
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = (lo + hi) >> 1;
        const uint64_t off = rp0[mid] + (rp1 ? rp1[mid] : 0);
        if (off < target) lo = mid + 1; else hi = mid;
    }
    bounds[r] = lo & ~31ull;
}

Partition make_partition(const uint64_t *rp0, const uint64_t *rp1, uint64_t n)
{
    Context &c = ctx();
    Partition p;
    p.b.assign(c.nranks + 1, 0);
    p.b[c.nranks] = n;
    if (c.nranks > 1) {
        GX_REQUIRE(c.nranks <= 255, "too many ranks");
        DevBuf<uint64_t> d(c.nranks + 1);
        GX_LAUNCH(k_partition, 1, 256, 0, rp0, rp1, n, c.nranks, d.p);
        read_back(p.b.data(), d.p, (c.nranks + 1) * sizeof(uint64_t));
        for (int r = 1; r <= c.nranks; r++)
            if (p.b[r] < p.b[r - 1]) p.b[r] = p.b[r - 1];
    }
    p.lo = p.b[c.rank];
    p.hi = p.b[c.rank + 1];
    return p;
}

Partition make_even_partition(uint64_t count, uint64_t align)
{
    Context &c = ctx();
    Partition p;
    p.b.assign(c.nranks + 1, 0);
    for (int r = 1; r < c.nranks; r++) {
        uint64_t x = (uint64_t)(((unsigned __int128)count * (unsigned)r) / (unsigned)c.nranks);
        p.b[r] = x / align * align;
    }
    p.b[c.nranks] = count;
    p.lo = p.b[c.rank];
    p.hi = p.b[c.rank + 1];
    return p;
}

// Mapped buffers released by a graph are parked here and handed to the next graph that asks for a
// similar size: cudaMalloc + cudaIpcOpenMemHandle on every rank and cudaIpcCloseMemHandle + cudaFree
// cost ~15 ms per graph.  Every rank allocates and frees in the same order with the same sizes, so
// the ranks' caches stay in lock-step and a reused buffer is mapped on all of them.
static std::vector<PeerBuf> g_parked;
constexpr size_t MAX_PARKED = 4;

static void peer_release(PeerBuf &b)
{
    Context &c = ctx();
    for (int r = 0; r < MAX_PEERS; r++)
        if (b.peer[r] && b.peer[r] != b.local) cudaIpcCloseMemHandle(b.peer[r]);
    if (b.local) { if (b.pooled) cudaFreeAsync(b.local, c.stream); else cudaFree(b.local); }
    b = PeerBuf{};
}

static PeerMail g_mail;
static bool g_mail_tried = false;

PeerMail &context_mail()
{
    if (!g_mail_tried) {
        g_mail_tried = true;
        peer_mail_open(g_mail);
    }
    return g_mail;
}

void peer_cache_clear()
{
    if (g_mail_tried) {
        if (ctx().ready) cudaStreamSynchronize(ctx().stream);
        peer_mail_close(g_mail);
        g_mail_tried = false;
    }
    if (g_parked.empty()) return;
    if (ctx().ready) cudaStreamSynchronize(ctx().stream);
    for (PeerBuf &p : g_parked) peer_release(p);
    g_parked.clear();
}

void peer_alloc(PeerBuf &b, size_t bytes)
{
    Context &c = ctx();
    b = PeerBuf{};
    for (size_t i = 0; i < g_parked.size(); i++) {
        if (g_parked[i].capacity >= bytes && g_parked[i].capacity <= 2 * bytes + 4096) {
            b = g_parked[i];
            b.bytes = bytes;
            g_parked.erase(g_parked.begin() + i);
            return;
        }
    }
    b.bytes = bytes;
    b.capacity = bytes ? bytes : 16;
    if (c.nranks <= 1 || c.nranks > MAX_PEERS) {
        // nothing to share: take it from the stream-ordered pool like every other buffer
        GX_CUDA(cudaMallocAsync(&b.local, bytes ? bytes : 16, c.stream));
        b.pooled = true;
        b.peer[0] = b.local;
        return;
    }
    GX_CUDA(cudaMalloc(&b.local, bytes ? bytes : 16)); // plain cudaMalloc: exportable through cudaIpcGetMemHandle
    b.peer[c.rank] = b.local;
    // exchange the 64-byte IPC handles with an all-gather, then map every peer's buffer
    cudaIpcMemHandle_t mine;
    GX_CUDA(cudaIpcGetMemHandle(&mine, b.local));
    DevBuf<cudaIpcMemHandle_t> all(c.nranks);
    GX_CUDA(cudaMemcpyAsync(all.p + c.rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, c.stream));
    GX_NCCL(ncclAllGather(all.p + c.rank, all.p, sizeof(mine), ncclUint8, (ncclComm_t)c.nccl_comm, c.stream));
    std::vector<cudaIpcMemHandle_t> h(c.nranks);
    GX_CUDA(cudaMemcpyAsync(h.data(), all.p, c.nranks * sizeof(mine), cudaMemcpyDeviceToHost, c.stream));
    GX_CUDA(cudaStreamSynchronize(c.stream));
    bool ok = true;
    for (int r = 0; r < c.nranks && ok; r++) {
        if (r == c.rank) continue;
        if (cudaIpcOpenMemHandle(&b.peer[r], h[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            b.peer[r] = nullptr;
            ok = false;
        }
    }
    // all ranks must agree on using the peer path
    DevBuf<int> flag(1);
    int hflag = ok ? 1 : 0;
    GX_CUDA(cudaMemcpyAsync(flag.p, &hflag, sizeof(int), cudaMemcpyHostToDevice, c.stream));
    GX_NCCL(ncclAllReduce(flag.p, flag.p, 1, ncclInt32, ncclMin, (ncclComm_t)c.nccl_comm, c.stream));
    read_back(&hflag, flag.p, sizeof(int));
    b.shared = hflag == 1;
}

void peer_free(PeerBuf &b)
{
    Context &c = ctx();
    if (b.shared && c.nccl_comm && g_parked.size() < MAX_PARKED) { // keep the mapping for the next graph
        g_parked.push_back(b);
        b = PeerBuf{};
        return;
    }
    if (c.ready) cudaStreamSynchronize(c.stream);
    peer_release(b);
}

void peer_mail_open(PeerMail &m)
{
    Context &c = ctx();
    m = PeerMail{};
    peer_alloc(m.buf, (size_t)2 * MAX_PEERS * MAIL_WORDS * sizeof(unsigned long long));
    if (!m.buf.shared) { peer_free(m.buf); return; }
    GX_CUDA(cudaMemsetAsync(m.buf.local, 0, m.buf.bytes, c.stream));
    for (int r = 0; r < c.nranks; r++) m.table.peer[r] = (unsigned long long *)m.buf.peer[r];
    m.table.nranks = c.nranks;
    m.table.rank = c.rank;
    GX_CUDA(cudaHostAlloc((void **)&m.host, 8 * sizeof(unsigned long long), cudaHostAllocMapped));
    memset(m.host, 0, 8 * sizeof(unsigned long long));
    GX_CUDA(cudaHostGetDevicePointer((void **)&m.host_dev, m.host, 0));
    // nobody may store into a mailbox before its owner has cleared it
    DevBuf<int> flag(1);
    flag.zero();
    allreduce(flag.p, 1, Dt::I32, Red::Max);
    GX_CUDA(cudaStreamSynchronize(c.stream));
    m.seq = 0;
    m.ok = true;
}

void peer_mail_close(PeerMail &m)
{
    if (m.host) cudaFreeHost(m.host);
    m.host = m.host_dev = nullptr;
    if (m.buf.local) peer_free(m.buf);
    m.ok = false;
}

void peer_mail_wait(PeerMail &m, unsigned long long seq, unsigned long long out[3])
{
    volatile unsigned long long *h = m.host + (size_t)(seq & 1ull) * 4;
    const auto t0 = std::chrono::steady_clock::now();
    unsigned spins = 0;
    for (;;) {
        const unsigned long long s = h[3];
        if (s == seq) break;
        if (s == ~0ull) throw Error(GX_ERR_CUDA, "peer mailbox exchange timed out on the device (a rank stopped responding)");
        if ((++spins & 0xFFFFu) == 0) {
            if (cudaStreamQuery(ctx().stream) != cudaErrorNotReady && h[3] != seq) {
                GX_CUDA(cudaGetLastError());
                if (h[3] == seq) break;
            }
            if (std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20))
                throw Error(GX_ERR_CUDA, "peer mailbox exchange timed out on the host");
        }
    }
    out[0] = h[0];
    out[1] = h[1];
    out[2] = h[2];
}

static ncclDataType_t nccl_dt(Dt dt)
{
    switch (dt) {
    case Dt::U32: return ncclUint32;
    case Dt::I32: return ncclInt32;
    case Dt::U64: return ncclUint64;
    case Dt::F64: return ncclFloat64;
    default: return ncclUint8;
    }
}

static size_t dt_size(Dt dt) { return dt == Dt::U8 ? 1 : (dt == Dt::U32 || dt == Dt::I32) ? 4 : 8; }

void allgatherv(void *buf, Dt dt, const Partition &p, uint64_t div, uint64_t total)
{
    Context &c = ctx();
    if (c.nranks <= 1) return;
    const bool prof = profiling();
    if (prof) prof_begin("nccl_allgatherv");
    // variable block sizes: one broadcast per owner, fused by the group into a single NCCL operation
    GX_NCCL(ncclGroupStart());
    for (int r = 0; r < c.nranks; r++) {
        uint64_t a = p.b[r] / div;
        uint64_t b = (r == c.nranks - 1 && total) ? total : p.b[r + 1] / div;
        if (b <= a) continue;
        char *ptr = (char *)buf + a * dt_size(dt);
        GX_NCCL(ncclBroadcast(ptr, ptr, b - a, nccl_dt(dt), r, (ncclComm_t)c.nccl_comm, c.stream));
    }
    GX_NCCL(ncclGroupEnd());
    if (prof) prof_end();
}

void allgather_equal(void *buf, Dt dt, uint64_t count_per_rank)
{
    Context &c = ctx();
    if (c.nranks <= 1 || count_per_rank == 0) return;
    const bool prof = profiling();
    if (prof) prof_begin("nccl_allgather");
    char *mine = (char *)buf + (uint64_t)c.rank * count_per_rank * dt_size(dt);
    GX_NCCL(ncclAllGather(mine, buf, count_per_rank, nccl_dt(dt), (ncclComm_t)c.nccl_comm, c.stream));
    if (prof) prof_end();
}

void allreduce(void *buf, uint64_t count, Dt dt, Red op)
{
    Context &c = ctx();
    if (c.nranks <= 1 || count == 0) return;
    const bool prof = profiling();
    if (prof) prof_begin("nccl_allreduce");
    const ncclRedOp_t o = op == Red::Sum ? ncclSum : op == Red::Min ? ncclMin : ncclMax;
    GX_NCCL(ncclAllReduce(buf, buf, count, nccl_dt(dt), o, (ncclComm_t)c.nccl_comm, c.stream));
    if (prof) prof_end();
}

} // namespace gx
