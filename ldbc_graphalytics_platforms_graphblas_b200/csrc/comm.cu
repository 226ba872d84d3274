// comm.cu -- NCCL collectives on the library stream (see comm.cuh for the sharding model).
#include <nccl.h>

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
