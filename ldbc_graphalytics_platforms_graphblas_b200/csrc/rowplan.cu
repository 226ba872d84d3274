// rowplan.cu -- long-row chunk plan (load balancing for power-law degree skew).
#include <cub/device/device_scan.cuh>

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

// chunks per long row (one extra zero item so that the exclusive scan ends with the total)
__global__ void k_long_chunk_counts(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ rows, uint64_t count,
                                    uint32_t *__restrict__ nch)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > count) return;
    nch[i] = i < count ? (uint32_t)((rowptr[rows[i] + 1] - rowptr[rows[i]] + CHUNK - 1) / CHUNK) : 0u;
}

__global__ void k_fill_chunks(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ rows, uint64_t count,
                              const uint32_t *__restrict__ first, uint32_t *__restrict__ chunk_row, uint64_t *__restrict__ chunk_begin)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint32_t v = rows[i];
    const uint64_t a = rowptr[v];
    const uint32_t f = first[i], k_end = first[i + 1] - f;
    for (uint32_t k = 0; k < k_end; k++) { chunk_row[f + k] = v; chunk_begin[f + k] = a + (uint64_t)k * CHUNK; }
}

// Everything stays on the device: long rows are collected in one pass (at most m / ROW_SPLIT of them), sorted
// by id, their chunk counts scanned; two scalars come back to size the arrays.
void ensure_plan(Adj &a, uint64_t n)
{
    RowPlan &p = a.plan;
    if (p.built) return;
    p.n_long = p.n_chunks = 0;
    p.part = make_partition(a.rowptr.p, nullptr, n);
    if (n == 0) { p.built = true; return; }
    const uint64_t v0 = p.part.lo, v1 = p.part.hi;
    const uint64_t cap = a.col.n / ROW_SPLIT + 1; // a long row has more than ROW_SPLIT entries
    DevBuf<unsigned long long> cnt(1);
    cnt.zero();
    DevBuf<uint32_t> found(cap);
    GX_LAUNCH(k_collect_long, grid_persistent(8), 256, 0, a.rowptr.p, v0, v1, ROW_SPLIT, found.p, cnt.p, cap);
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
    GX_CUDA(cudaMemcpyAsync(p.long_rows.p, found.p, nl * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx().stream));
    sort_keys32(p.long_rows, nl, bits_for(n)); // ascending vertex ids whatever order the atomics appended them in
    GX_LAUNCH(k_long_chunk_counts, grid_for(nl + 1, 256), 256, 0, a.rowptr.p, p.long_rows.p, (uint64_t)nl, p.long_first_chunk.p);
    {
        size_t tb = 0;
        GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, p.long_first_chunk.p, p.long_first_chunk.p, (int64_t)(nl + 1), ctx().stream));
        DevBuf<char> tmp(tb);
        GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, p.long_first_chunk.p, p.long_first_chunk.p, (int64_t)(nl + 1), ctx().stream));
        count_launch();
    }
    uint32_t total = 0;
    read_back(&total, p.long_first_chunk.p + nl, sizeof(total));
    p.n_chunks = total;
    p.chunk_row.alloc(p.n_chunks ? p.n_chunks : 1);
    p.chunk_begin.alloc(p.n_chunks ? p.n_chunks : 1);
    GX_LAUNCH(k_fill_chunks, grid_for(nl, 256), 256, 0, a.rowptr.p, p.long_rows.p, (uint64_t)nl, p.long_first_chunk.p, p.chunk_row.p,
              p.chunk_begin.p);
    p.built = true;
}

} // namespace gx
