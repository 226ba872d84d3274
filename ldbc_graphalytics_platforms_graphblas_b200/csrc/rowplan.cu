// rowplan.cu -- long-row chunk plan (load balancing for power-law degree skew).
#include <algorithm>
#include <vector>

#include "graph.cuh"

namespace gx {

__global__ void k_collect_long(const uint64_t *__restrict__ rowptr, uint64_t v0, uint64_t v1, uint32_t thresh,
                               uint32_t *__restrict__ list, unsigned long long *__restrict__ count, uint64_t cap)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < v1; v += stride) {
        if (rowptr[v + 1] - rowptr[v] > thresh) {
            unsigned long long pos = atomicAdd(count, 1ull);
            if (pos < cap) list[pos] = (uint32_t)v;
        }
    }
}

__global__ void k_gather_pairs(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ rows, uint64_t count,
                               uint64_t *__restrict__ pairs)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { pairs[2 * i] = rowptr[rows[i]]; pairs[2 * i + 1] = rowptr[rows[i] + 1]; }
}

void ensure_plan(Adj &a, uint64_t n)
{
    RowPlan &p = a.plan;
    if (p.built) return;
    p.n_long = p.n_chunks = 0;
    p.part = make_partition(a.rowptr.p, nullptr, n);
    if (n == 0) { p.built = true; return; }
    const uint64_t v0 = p.part.lo, v1 = p.part.hi;
    // at most m / ROW_SPLIT rows can be long; size the list by a first counting pass
    DevBuf<unsigned long long> cnt(1);
    cnt.zero();
    DevBuf<uint32_t> dummy(1);
    GX_LAUNCH(k_collect_long, grid_persistent(8), 256, 0, a.rowptr.p, v0, v1, ROW_SPLIT, dummy.p, cnt.p, (uint64_t)0);
    unsigned long long nl = 0;
    read_back(&nl, cnt.p, sizeof(nl));
    p.n_long = nl;
    p.long_rows.alloc(nl ? nl : 1);
    p.long_first_chunk.alloc(nl + 1);
    if (nl == 0) {
        p.long_first_chunk.zero();
        p.chunk_row.alloc(1);
        p.chunk_begin.alloc(1);
        p.built = true;
        return;
    }
    cnt.zero();
    GX_LAUNCH(k_collect_long, grid_persistent(8), 256, 0, a.rowptr.p, v0, v1, ROW_SPLIT, p.long_rows.p, cnt.p, (uint64_t)nl);
    std::vector<uint32_t> rows(nl);
    GX_CUDA(cudaMemcpyAsync(rows.data(), p.long_rows.p, nl * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx().stream));
    GX_CUDA(cudaStreamSynchronize(ctx().stream));
    std::sort(rows.begin(), rows.end());
    // offsets of the long rows only: gathered on the device, one small copy back
    GX_CUDA(cudaMemcpyAsync(p.long_rows.p, rows.data(), nl * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx().stream));
    DevBuf<uint64_t> pairs(2 * nl);
    GX_LAUNCH(k_gather_pairs, grid_for(nl, 256), 256, 0, a.rowptr.p, p.long_rows.p, (uint64_t)nl, pairs.p);
    std::vector<uint64_t> rp_pairs(2 * nl);
    GX_CUDA(cudaMemcpyAsync(rp_pairs.data(), pairs.p, 2 * nl * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx().stream));
    GX_CUDA(cudaStreamSynchronize(ctx().stream));
    std::vector<uint32_t> first(nl + 1), crow;
    std::vector<uint64_t> cbeg;
    for (size_t i = 0; i < nl; i++) {
        first[i] = (uint32_t)crow.size();
        for (uint64_t b = rp_pairs[2 * i]; b < rp_pairs[2 * i + 1]; b += CHUNK) { crow.push_back(rows[i]); cbeg.push_back(b); }
    }
    first[nl] = (uint32_t)crow.size();
    p.n_chunks = crow.size();
    p.chunk_row.alloc(p.n_chunks);
    p.chunk_begin.alloc(p.n_chunks);
    cudaStream_t s = ctx().stream;
    GX_CUDA(cudaMemcpyAsync(p.long_first_chunk.p, first.data(), (nl + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    GX_CUDA(cudaMemcpyAsync(p.chunk_row.p, crow.data(), p.n_chunks * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
    GX_CUDA(cudaMemcpyAsync(p.chunk_begin.p, cbeg.data(), p.n_chunks * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    GX_CUDA(cudaStreamSynchronize(s)); // host vectors go out of scope
    p.built = true;
}

} // namespace gx
