// graph.cu -- host CSR -> device graph, transposition (LAGraph_Cached_AT
// analogue), LCC union/orientation cache.  Replaces GxB_Matrix_export_CSR + the
// cudaMalloc/cudaMemcpy block of the reference's cdlp_gpu
// (cdlp_cuda.cu:181, cdlp_kernel.cu:1162-1196).
//
//   one GPU     gx_graph_create_csr32_cached(GX_CACHE_AT): upload in row-block chunks on a copy stream,
//               every chunk validated / sorted by column while the next one is on the bus, one merge
//               pass at the end (upload_transpose_pipelined)
//   several     every rank uploads 1/nranks of the arrays, all-gather over NVLink (upload_array);
//               transposition split by column range from 4 ranks on (transpose_partitioned)
//   LCC cache   degree-oriented union graph, membership tables of the longer rows, entries ordered by
//               the owner of the longer list (ensure_lcc_cache)
//
// Construction-time sorting / compaction uses CUB device primitives (CUDA
// toolkit headers); the per-algorithm hot loops in algo_*.cu are hand-written.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/tuple.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "graph.cuh"

namespace gx {

int bits_for(uint64_t n)
{
    int b = 1;
    while (b < 64 && (1ull << b) < n) b++;
    return b;
}

// ------------------------------------------------------------------------- small kernels
__global__ void k_narrow_u64(const uint64_t *__restrict__ in, uint32_t *__restrict__ out, uint64_t count,
                             uint64_t n, int *__restrict__ bad)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < count; i += stride) {
        uint64_t v = in[i];
        if (v >= n) *bad = 1;
        out[i] = (uint32_t)v;
    }
}

__global__ void k_validate(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t n,
                           uint64_t m, int *__restrict__ bad, int *__restrict__ unsorted)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t v = i; v < n; v += stride) {
        uint64_t a = rowptr[v], b = rowptr[v + 1];
        if (a > b || b > m) { *bad = 1; continue; }
        if (v == 0 && a != 0) *bad = 1;
        if (v == n - 1 && b != m) *bad = 1;
    }
    for (uint64_t e = i; e < m; e += stride)
        if (col[e] >= n) *bad = 1;
    // sortedness: entry e and e+1 in the same row must be non-decreasing.  Row membership of
    // e+1 is checked in k_check_sorted (needs row ids); here only the cheap global hint.
    (void)unsorted;
}

// Entry-parallel sortedness check: a decrease col[e] < col[e-1] is legal only where a row
// starts at e, which a binary search in rowptr decides (decreases are rare: ~one per row).
__global__ void k_check_sorted(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t n,
                               uint64_t m, int *__restrict__ unsorted)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) {
        if (col[e] >= col[e - 1]) continue;
        uint64_t lo = 0, hi = n; // is there a row r with rowptr[r] == e ?
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] < e) lo = mid + 1; else hi = mid;
        }
        if (lo >= n || rowptr[lo] != e) *unsorted = 1;
    }
}

// row id of every entry: every non-empty row writes its id at its first entry, an inclusive
// max-scan spreads it over the row (12 bytes of traffic per entry, no per-entry search and no
// per-row loop, so degree skew does not matter).
__global__ void k_row_heads(const uint64_t *__restrict__ rowptr, uint64_t n, uint32_t *__restrict__ row_of_edge)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) {
        const uint64_t a = rowptr[v];
        if (rowptr[v + 1] > a) row_of_edge[a] = (uint32_t)v;
    }
}

struct MaxU32 {
    __host__ __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

template <class K, int SHIFT>
__global__ void k_rowptr_from_sorted(const K *__restrict__ keys, uint64_t m, uint64_t n, uint64_t *__restrict__ rowptr)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v <= n; v += stride) {
        uint64_t lo = 0, hi = m; // first position with (key >> SHIFT) >= v
        while (lo < hi) {
            uint64_t mid = (lo + hi) >> 1;
            if ((uint64_t)(keys[mid] >> SHIFT) < v) lo = mid + 1; else hi = mid;
        }
        rowptr[v] = lo;
    }
}

void expand_row_ids(const uint64_t *rowptr, uint64_t n, uint64_t m, uint32_t *row_of_edge)
{
    if (!m) return;
    GX_CUDA(cudaMemsetAsync(row_of_edge, 0, m * sizeof(uint32_t), ctx().stream));
    GX_LAUNCH(k_row_heads, grid_persistent(8), 256, 0, rowptr, n, row_of_edge);
    size_t tb = 0;
    GX_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, row_of_edge, row_of_edge, MaxU32(), (int64_t)m, ctx().stream));
    DevBuf<char> tmp(tb);
    GX_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, row_of_edge, row_of_edge, MaxU32(), (int64_t)m, ctx().stream));
    count_launch();
}

void rowptr_from_sorted_rows(const uint32_t *sorted_rows, uint64_t m, uint64_t n, uint64_t *rowptr)
{
    GX_LAUNCH((k_rowptr_from_sorted<uint32_t, 0>), grid_for(n + 1, 256), 256, 0, sorted_rows, m, n, rowptr);
}

void rowptr_from_sorted_keys(const uint64_t *sorted_keys, uint64_t m, uint64_t n, uint64_t *rowptr)
{
    GX_LAUNCH((k_rowptr_from_sorted<uint64_t, 32>), grid_for(n + 1, 256), 256, 0, sorted_keys, m, n, rowptr);
}

void sort_keys64(DevBuf<uint64_t> &keys, uint64_t count, int end_bit)
{
    if (count < 2) return;
    DevBuf<uint64_t> alt(count);
    cub::DoubleBuffer<uint64_t> db(keys.p, alt.p);
    size_t tb = 0;
    GX_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, (int64_t)count, 0, end_bit, ctx().stream));
    DevBuf<char> tmp(tb);
    GX_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, db, (int64_t)count, 0, end_bit, ctx().stream));
    if (db.Current() != keys.p) std::swap(keys.p, alt.p); // sizes are equal; alt frees the other buffer
}

void sort_keys32(DevBuf<uint32_t> &keys, uint64_t count, int end_bit)
{
    if (count < 2) return;
    DevBuf<uint32_t> alt(keys.n);
    cub::DoubleBuffer<uint32_t> db(keys.p, alt.p);
    size_t tb = 0;
    GX_CUDA(cub::DeviceRadixSort::SortKeys(nullptr, tb, db, (int64_t)count, 0, end_bit, ctx().stream));
    DevBuf<char> tmp(tb);
    GX_CUDA(cub::DeviceRadixSort::SortKeys(tmp.p, tb, db, (int64_t)count, 0, end_bit, ctx().stream));
    if (db.Current() != keys.p) std::swap(keys.p, alt.p);
}

void sort_pairs64_f64(DevBuf<uint64_t> &keys, DevBuf<double> &vals, uint64_t count, int end_bit)
{
    if (count < 2) return;
    DevBuf<uint64_t> kalt(count);
    DevBuf<double> valt(count);
    cub::DoubleBuffer<uint64_t> dk(keys.p, kalt.p);
    cub::DoubleBuffer<double> dv(vals.p, valt.p);
    size_t tb = 0;
    GX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int64_t)count, 0, end_bit, ctx().stream));
    DevBuf<char> tmp(tb);
    GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (int64_t)count, 0, end_bit, ctx().stream));
    if (dk.Current() != keys.p) std::swap(keys.p, kalt.p);
    if (dv.Current() != vals.p) std::swap(vals.p, valt.p);
}

// ------------------------------------------------------------------------- upload + finish
__global__ void k_make_keys(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col, uint64_t m,
                            uint64_t *__restrict__ keys)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) keys[e] = ((uint64_t)row[e] << 32) | col[e];
}

__global__ void k_low32(const uint64_t *__restrict__ keys, uint64_t m, uint32_t *__restrict__ out)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) out[e] = (uint32_t)keys[e];
}

__global__ void k_high32(const uint64_t *__restrict__ keys, uint64_t m, uint32_t *__restrict__ out)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) out[e] = (uint32_t)(keys[e] >> 32);
}

static int read_flag(const int *dflag)
{
    int h = 0;
    read_back(&h, dflag, sizeof(int));
    return h;
}

void finish_graph(gx_graph *g)
{
    if (g->n == 0) return;
    DevBuf<int> flags(2);
    flags.zero();
    GX_LAUNCH(k_validate, grid_persistent(8), 256, 0, g->out.rowptr.p, g->out.col.p, g->n, g->m, flags.p, flags.p + 1);
    GX_LAUNCH(k_check_sorted, grid_persistent(8), 256, 0, g->out.rowptr.p, g->out.col.p, g->n, g->m, flags.p + 1);
    int h[2] = {0, 0};
    read_back(h, flags.p, sizeof(h));
    if (h[0]) throw Error(GX_ERR_INVALID, "CSR arrays are inconsistent (rowptr not monotone or column id >= n)");
    if (h[1] && g->m > 1) {
        // jumbled rows (GxB export may return them so): sort (row, col) keys once
        DevBuf<uint32_t> rows(g->m);
        expand_row_ids(g->out.rowptr.p, g->n, g->m, rows.p);
        DevBuf<uint64_t> keys(g->m);
        GX_LAUNCH(k_make_keys, grid_persistent(8), 256, 0, rows.p, g->out.col.p, g->m, keys.p);
        if (g->weighted) sort_pairs64_f64(keys, g->out.w, g->m, 32 + bits_for(g->n));
        else sort_keys64(keys, g->m, 32 + bits_for(g->n));
        GX_LAUNCH(k_low32, grid_persistent(8), 256, 0, keys.p, g->m, g->out.col.p);
    }
}

// Host -> device copy of one array.  On several GPUs every rank was handed the same host arrays
// (the adjacency is replicated, comm.cuh): a rank pushes only its 1/nranks slice over PCIe and the
// slices are all-gathered over NVLink, so the upload takes 1/nranks of the single-GPU time instead
// of nranks copies competing for the host's memory and PCIe bandwidth.
template <class T>
static void upload_array(T *dst, const T *src, uint64_t count, Dt dt)
{
    if (!count) return;
    cudaStream_t s = ctx().stream;
    if (!multi()) {
        GX_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(T), cudaMemcpyHostToDevice, s));
        return;
    }
    // equal slices so that the exchange is ONE ncclAllGather (4x the throughput of a group of broadcasts
    // with unequal sizes); the few elements beyond nranks * slice are uploaded by every rank
    const uint64_t nr = (uint64_t)ctx().nranks, per = count / nr / 64 * 64, main = per * nr;
    if (per) GX_CUDA(cudaMemcpyAsync(dst + ctx().rank * per, src + ctx().rank * per, per * sizeof(T), cudaMemcpyHostToDevice, s));
    if (count > main) GX_CUDA(cudaMemcpyAsync(dst + main, src + main, (count - main) * sizeof(T), cudaMemcpyHostToDevice, s));
    allgather_equal(dst, dt, per);
}

static void upload_common(gx_graph *g, uint64_t n, uint64_t nnz, const uint64_t *rowptr, const double *weights,
                          int directed)
{
    GX_REQUIRE(n < 0xFFFFFFFEull, "n must be < 2^32 - 2");
    GX_REQUIRE(rowptr != nullptr || n == 0, "rowptr is NULL");
    g->n = n;
    g->m = nnz;
    g->directed = directed != 0;
    g->weighted = weights != nullptr;
    g->out.rowptr.alloc(n + 1);
    if (n) upload_array(g->out.rowptr.p, rowptr, n + 1, Dt::U64);
    else g->out.rowptr.zero();
    g->out.col.alloc(nnz);
    if (weights) {
        g->out.w.alloc(nnz);
        upload_array(g->out.w.p, weights, nnz, Dt::F64);
    }
}

// Several GPUs: the transposition is split by COLUMN range.  A rank keeps the (column, row) pairs whose
// column lies in its 1/nranks slice of the vertices (stable selection), sorts only those, and the
// sorted slices -- consecutive pieces of the in-edge adjacency -- are all-gathered over NVLink.
__global__ void k_flag_col_range(const uint32_t *__restrict__ col, uint64_t m, uint32_t lo, uint32_t hi, uint8_t *__restrict__ flag)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) { const uint32_t c = ld_stream(col + e); flag[e] = (c >= lo && c < hi) ? 1 : 0; }
}

// rowptr[v] = base + first position in the rank's sorted keys with key >= v, for v in [v_lo, v_hi)
__global__ void k_rowptr_slice(const uint32_t *__restrict__ keys, uint64_t cnt, uint64_t base, uint64_t v_lo, uint64_t v_hi,
                               uint64_t *__restrict__ rowptr)
{
    uint64_t v = v_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < v_hi; v += stride) {
        uint64_t lo = 0, hi = cnt;
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (keys[mid] < v) lo = mid + 1; else hi = mid;
        }
        rowptr[v] = base + lo;
    }
}

// GX_TIMING_DEBUG=1: wall-clock per phase of the multi-GPU graph construction on stderr (rank 0)
struct PhaseLog {
    bool on;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseLog() : on(getenv("GX_TIMING_DEBUG") && ctx().rank == 0), t0(std::chrono::steady_clock::now()) {}
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(ctx().stream);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gx timing] %-28s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static void transpose_partitioned(gx_graph *g)
{
    Context &c = ctx();
    PhaseLog pl;
    const uint64_t n = g->n, m = g->m;
    const Partition pv = make_even_partition(n, 32);
    DevBuf<uint32_t> rows(m);
    DevBuf<uint8_t> flag(m);
    expand_row_ids(g->out.rowptr.p, n, m, rows.p);
    pl.mark("T row ids");
    GX_LAUNCH(k_flag_col_range, grid_persistent(8), 256, 0, g->out.col.p, m, (uint32_t)pv.lo, (uint32_t)pv.hi, flag.p);
    // the selection cannot hold more than m pairs; m / nranks on average -- sized by a first counting pass
    DevBuf<uint64_t> nsel(1);
    const uint64_t cap = m;
    DevBuf<uint32_t> keys(cap), vals(cap);
    size_t tb = 0, tb2 = 0;
    GX_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, g->out.col.p, flag.p, keys.p, nsel.p, (int64_t)m, c.stream));
    {
        DevBuf<char> tmp(tb);
        GX_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, g->out.col.p, flag.p, keys.p, nsel.p, (int64_t)m, c.stream));
        GX_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, rows.p, flag.p, vals.p, nsel.p, (int64_t)m, c.stream));
        count_launch(2);
    }
    uint64_t mine = 0;
    read_back(&mine, nsel.p, sizeof(mine));
    pl.mark("T select own columns");
    rows.release();
    flag.release();
    // sizes of all slices -> where this rank's slice starts in the in-edge adjacency
    DevBuf<uint64_t> sizes(c.nranks);
    GX_CUDA(cudaMemcpyAsync(sizes.p + c.rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, c.stream));
    allgather_equal(sizes.p, Dt::U64, 1);
    std::vector<uint64_t> hs(c.nranks);
    read_back(hs.data(), sizes.p, c.nranks * sizeof(uint64_t));
    Partition pe;
    pe.b.assign(c.nranks + 1, 0);
    for (int r = 0; r < c.nranks; r++) pe.b[r + 1] = pe.b[r] + hs[r];
    pe.lo = pe.b[c.rank];
    pe.hi = pe.b[c.rank + 1];
    GX_REQUIRE(pe.b[c.nranks] == m, "transposition slices do not add up");
    pl.mark("T slice sizes");
    if (mine > 1) {
        DevBuf<uint32_t> keys_alt(mine), vals_alt(mine);
        cub::DoubleBuffer<uint32_t> dk(keys.p, keys_alt.p), dv(vals.p, vals_alt.p);
        GX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, (int64_t)mine, 0, bits_for(n), c.stream));
        DevBuf<char> tmp(tb2);
        GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb2, dk, dv, (int64_t)mine, 0, bits_for(n), c.stream));
        count_launch();
        if (dk.Current() != keys.p) std::swap(keys.p, keys_alt.p);
        if (dv.Current() != vals.p) std::swap(vals.p, vals_alt.p);
        GX_CUDA(cudaStreamSynchronize(c.stream)); // keys_alt / vals_alt are released by scope
    }
    pl.mark("T sort own slice");
    if (pv.hi > pv.lo) GX_LAUNCH(k_rowptr_slice, grid_persistent(8), 256, 0, keys.p, mine, pe.lo, pv.lo, pv.hi, g->in.rowptr.p);
    GX_CUDA(cudaMemcpyAsync(g->in.rowptr.p + n, &m, sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
    const char *xe = getenv("GX_TRANSPOSE_XCHG"); // "bcast" / "gather" force one form (tests)
    const bool by_gather = xe ? xe[0] == 'g' : c.nranks <= 4;
    if (by_gather) {
        // the slices differ in size: one equal-size ncclAllGather into a padded scratch, then local copies
        // into place (a group of unequal broadcasts moves the same bytes at a quarter of the throughput).
        // Measured on 4 GPUs (RMAT-24): 1.4 ms for 1.1 GB.  On 8 GPUs (RMAT-25) the same exchange took 27 ms
        // and the ranks arrived 16 ms apart (profiles/r1c_dbg_n8_timing.txt) where the grouped broadcasts had
        // given 29 ms for the whole transposition, so from 5 ranks on the slices go by broadcast.
        uint64_t slot = 0;
        for (int r = 0; r < c.nranks; r++) slot = std::max(slot, hs[r]);
        slot = (slot + 63) / 64 * 64;
        DevBuf<uint32_t> scratch(slot * (uint64_t)c.nranks);
        if (mine) GX_CUDA(cudaMemcpyAsync(scratch.p + slot * c.rank, vals.p, mine * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
        pl.mark("T scratch alloc + copy");
        allgather_equal(scratch.p, Dt::U32, slot);
        pl.mark("T all-gather slices");
        for (int r = 0; r < c.nranks; r++)
            if (hs[r])
                GX_CUDA(cudaMemcpyAsync(g->in.col.p + pe.b[r], scratch.p + slot * r, hs[r] * sizeof(uint32_t), cudaMemcpyDeviceToDevice,
                                        c.stream));
    } else {
        if (mine) GX_CUDA(cudaMemcpyAsync(g->in.col.p + pe.lo, vals.p, mine * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
        allgatherv(g->in.col.p, Dt::U32, pe);
        pl.mark("T broadcast slices");
    }
    pl.mark("T copies into place");
    allgatherv(g->in.rowptr.p, Dt::U64, pv);
    GX_CUDA(cudaStreamSynchronize(c.stream)); // &m and the scoped buffers
    pl.mark("T all-gather offsets");
}

// Several GPUs, block-local form: a rank ends up with the in-edge rows of ITS row block only, and nothing of the
// adjacency crosses NVLink.  The in-degrees come first -- every rank counts the columns of 1/nranks of the entries, one
// all-reduce of n counters sums them -- so the offsets of all rows (and with them the nnz-balanced row blocks) are known
// on every rank before anything is sorted; a rank then selects the (column, row) pairs whose column lies in its block,
// sorts those by column (stable: rows stay ascending) and writes them at the block's offset.  The column-split form
// above moves the whole in-edge adjacency through an all-gather instead (2.1 GB on RMAT-25: 27 ms on 8 GPUs).
__global__ void k_col_histogram(const uint32_t *__restrict__ col, uint64_t count, uint32_t *__restrict__ deg)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < count; e += stride) atomicAdd(&deg[ld_stream(col + e)], 1u);
}

__global__ void k_widen_counts(const uint32_t *__restrict__ deg, uint64_t n, uint64_t *__restrict__ out)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v <= n; v += stride) out[v] = v < n ? deg[v] : 0;
}

struct ColumnInBlock {
    uint32_t lo, hi;
    __host__ __device__ bool operator()(const thrust::tuple<uint32_t, uint32_t> &p) const
    {
        return thrust::get<0>(p) >= lo && thrust::get<0>(p) < hi;
    }
};

static void transpose_block_local(gx_graph *g)
{
    Context &c = ctx();
    PhaseLog pl;
    const uint64_t n = g->n, m = g->m;
    {
        DevBuf<uint32_t> deg(n);
        deg.zero();
        const Partition ps = make_even_partition(m);
        if (ps.hi > ps.lo) GX_LAUNCH(k_col_histogram, grid_persistent(8), 256, 0, g->out.col.p + ps.lo, ps.hi - ps.lo, deg.p);
        allreduce(deg.p, n, Dt::U32, Red::Sum);
        GX_LAUNCH(k_widen_counts, grid_persistent(8), 256, 0, deg.p, n, g->in.rowptr.p);
        size_t tb = 0;
        GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, g->in.rowptr.p, g->in.rowptr.p, (int64_t)(n + 1), c.stream));
        DevBuf<char> tmp(tb);
        GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, g->in.rowptr.p, g->in.rowptr.p, (int64_t)(n + 1), c.stream));
        count_launch();
    }
    pl.mark("T in-degrees + offsets");
    // the row blocks ensure_plan(in) will use (same kernel, same input: the same boundaries on every rank)
    const Partition part = make_partition(g->in.rowptr.p, nullptr, n);
    g->in_block_entries.assign(c.nranks + 1, 0);
    for (int r = 1; r <= c.nranks; r++) read_back(&g->in_block_entries[r], g->in.rowptr.p + part.b[r], sizeof(uint64_t));
    const uint64_t e0 = g->in_block_entries[c.rank], mine = g->in_block_entries[c.rank + 1] - e0;
    if (mine) {
        DevBuf<uint32_t> keys(mine), vals(mine);
        {
            // one stable selection over (column, row) pairs: the columns are read once, the rows once
            DevBuf<uint32_t> rows(m);
            DevBuf<uint64_t> nsel(1);
            expand_row_ids(g->out.rowptr.p, n, m, rows.p);
            auto pairs = thrust::make_zip_iterator(thrust::make_tuple((const uint32_t *)g->out.col.p, (const uint32_t *)rows.p));
            auto picked = thrust::make_zip_iterator(thrust::make_tuple(keys.p, vals.p));
            const ColumnInBlock own{(uint32_t)part.lo, (uint32_t)part.hi};
            size_t tb = 0;
            GX_CUDA(cub::DeviceSelect::If(nullptr, tb, pairs, picked, nsel.p, (int64_t)m, own, c.stream));
            DevBuf<char> tmp(tb);
            GX_CUDA(cub::DeviceSelect::If(tmp.p, tb, pairs, picked, nsel.p, (int64_t)m, own, c.stream));
            count_launch();
            uint64_t got = 0;
            read_back(&got, nsel.p, sizeof(got));
            GX_REQUIRE(got == mine, "transposition: the selected pairs do not match the in-degree offsets");
        }
        pl.mark("T select own columns");
        DevBuf<uint32_t> keys_alt(mine);
        cub::DoubleBuffer<uint32_t> dk(keys.p, keys_alt.p), dv(vals.p, g->in.col.p + e0); // the block's place is the alternate buffer
        size_t tb2 = 0;
        GX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, (int64_t)mine, 0, bits_for(n), c.stream));
        DevBuf<char> tmp(tb2);
        GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb2, dk, dv, (int64_t)mine, 0, bits_for(n), c.stream));
        count_launch();
        if (dv.Current() != g->in.col.p + e0)
            GX_CUDA(cudaMemcpyAsync(g->in.col.p + e0, dv.Current(), mine * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
        GX_CUDA(cudaStreamSynchronize(c.stream)); // the scoped buffers
        pl.mark("T sort own block");
    }
    g->in_block_only = true;
}

void ensure_in_full(gx_graph *g)
{
    if (!g->in_block_only) return;
    PhaseTimer t(&ctx().timing.build_ms);
    Partition pe;
    pe.b = g->in_block_entries;
    pe.lo = pe.b[ctx().rank];
    pe.hi = pe.b[ctx().rank + 1];
    allgatherv(g->in.col.p, Dt::U32, pe);
    g->in_block_only = false;
}

void ensure_in_adj(gx_graph *g)
{
    if (!g->directed || g->have_in) return;
    PhaseTimer t(&ctx().timing.build_ms);
    const uint64_t n = g->n, m = g->m;
    g->in.rowptr.alloc(n + 1);
    g->in.col.alloc(m);
    if (m == 0) {
        g->in.rowptr.zero();
        g->have_in = true;
        return;
    }
    // splitting pays from 4 ranks on (at 2 the selection passes cost what halving the sort saves);
    // GX_TRANSPOSE_SPLIT=0 / 1 forces it off / on
    // several GPUs: block-local by default (GX_TRANSPOSE_LOCAL=0: the older forms -- column-split with an all-gather of
    // the sorted slices from 4 ranks on or with GX_TRANSPOSE_SPLIT=1, the whole sort replicated otherwise)
    const char *le = getenv("GX_TRANSPOSE_LOCAL");
    if (multi() && m >= (1ull << 16) && !(le && le[0] == '0')) {
        transpose_block_local(g);
        g->have_in = true;
        return;
    }
    const char *pe = getenv("GX_TRANSPOSE_SPLIT");
    const bool split = pe ? pe[0] != '0' : ctx().nranks >= 4;
    if (multi() && m >= (1ull << 16) && split) {
        transpose_partitioned(g);
        g->have_in = true;
        return;
    }
    // stable LSD radix sort of (key = column, value = row): rows stay ascending inside a column
    DevBuf<uint32_t> keys(m), keys_alt(m), rows(m);
    GX_CUDA(cudaMemcpyAsync(keys.p, g->out.col.p, m * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx().stream));
    expand_row_ids(g->out.rowptr.p, n, m, rows.p);
    cub::DoubleBuffer<uint32_t> dk(keys.p, keys_alt.p);
    cub::DoubleBuffer<uint32_t> dv(rows.p, g->in.col.p);
    size_t tb = 0;
    GX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int64_t)m, 0, bits_for(n), ctx().stream));
    DevBuf<char> tmp(tb);
    GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (int64_t)m, 0, bits_for(n), ctx().stream));
    if (dv.Current() != g->in.col.p) std::swap(rows.p, g->in.col.p);
    rowptr_from_sorted_rows(dk.Current(), m, n, g->in.rowptr.p);
    g->have_in = true;
}

// ------------------------------------------------------------------------- pipelined upload + transposition
// gx_graph_create_csr32_cached(..., GX_CACHE_AT) on one GPU: the column ids go up in row-block chunks
// on a copy stream; while chunk k+1 is on the PCIe bus, chunk k is validated, its entries get their row
// ids, a column histogram and a chunk-local stable radix sort by column ("run" k).  Rows ascend
// from run to run, so the in-edge adjacency is, per column, run 0's segment followed by run 1's, ...
// -- after the last chunk only its sort, one prefix pass over the histograms and ONE merge pass
// (8 B read, 4 B written per entry) remain, instead of a full 3-pass sort of all pairs after the
// upload has finished.
constexpr int UP_MAX_CHUNKS = 8;
struct ChunkBounds { uint64_t e[UP_MAX_CHUNKS + 1]; int k; };

__global__ void k_check_rowptr(const uint64_t *__restrict__ rowptr, uint64_t n, uint64_t m, int *__restrict__ bad)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) {
        const uint64_t a = rowptr[v], b = rowptr[v + 1];
        if (a > b || b > m || (v == 0 && a != 0) || (v == n - 1 && b != m)) *bad = 1;
    }
}

// column ids of [e_lo, e_hi) in range; a decrease inside the chunk is legal only at a row start
__global__ void k_check_chunk(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t n, uint64_t e_lo,
                              uint64_t e_hi, int *__restrict__ bad, int *__restrict__ unsorted)
{
    uint64_t e = e_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < e_hi; e += stride) {
        const uint32_t c = col[e];
        if (c >= n) { *bad = 1; continue; }
        if (e == e_lo || c >= col[e - 1]) continue;
        uint64_t lo = 0, hi = n; // is there a row r with rowptr[r] == e ?
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] < e) lo = mid + 1; else hi = mid;
        }
        if (lo >= n || rowptr[lo] != e) *unsorted = 1;
    }
}

__global__ void k_row_heads_range(const uint64_t *__restrict__ rowptr, uint64_t r_lo, uint64_t r_hi, uint64_t e_lo,
                                  uint64_t e_hi, uint32_t *__restrict__ row_of_edge)
{
    uint64_t v = r_lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < r_hi; v += stride) {
        const uint64_t a = rowptr[v];
        if (rowptr[v + 1] > a && a >= e_lo && a < e_hi) row_of_edge[a - e_lo] = (uint32_t)v; // (offsets not yet known to be sane)
    }
}

// rs[k][c] = first position of column c inside run k (n + 1 offsets per run)
// (all three merge kernels leave at once when the validation kernels before them on the stream flagged the input:
// with a column id >= n the runs are not sorted in the full key and the offsets below would be garbage)
__global__ void k_sum_runs(const uint64_t *__restrict__ rs, int K, uint64_t n, uint64_t *__restrict__ total, const int *__restrict__ bad)
{
    if (*bad) return;
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; c <= n; c += stride) {
        uint64_t t = 0;
        if (c < n)
            for (int k = 0; k < K; k++) t += rs[(uint64_t)k * (n + 1) + c + 1] - rs[(uint64_t)k * (n + 1) + c];
        total[c] = t;
    }
}

// delta[k][c] + (position inside run k) = destination of a run entry with column c
__global__ void k_run_delta(const uint64_t *__restrict__ in_rowptr, const uint64_t *__restrict__ rs, int K, uint64_t n,
                            uint64_t *__restrict__ delta, const int *__restrict__ bad)
{
    if (*bad) return;
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; c < n; c += stride) {
        uint64_t acc = in_rowptr[c];
        for (int k = 0; k < K; k++) {
            const uint64_t a = rs[(uint64_t)k * (n + 1) + c], b = rs[(uint64_t)k * (n + 1) + c + 1];
            delta[(uint64_t)k * n + c] = acc - a;
            acc += b - a;
        }
    }
}

__global__ void k_merge_runs(const uint32_t *__restrict__ run_keys, const uint32_t *__restrict__ run_vals, const ChunkBounds cb,
                             const uint64_t *__restrict__ delta, uint64_t n, uint32_t *__restrict__ in_col, const int *__restrict__ bad)
{
    if (*bad) return;
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t m = cb.e[cb.k];
    for (; e < m; e += stride) {
        int k = 0;
#pragma unroll
        for (int j = 1; j < UP_MAX_CHUNKS; j++) k += (j < cb.k && e >= cb.e[j]) ? 1 : 0;
        const uint32_t c = run_keys[e];
        if (c < n) in_col[delta[(uint64_t)k * n + c] + (e - cb.e[k])] = run_vals[e]; // c >= n: rejected below
    }
}

// returns false (nothing uploaded yet) when the shortcut does not apply; throws on invalid input
static bool upload_transpose_pipelined(gx_graph *g, uint64_t n, uint64_t nnz, const uint64_t *rowptr, const uint32_t *colidx,
                                       const double *weights)
{
    Context &c = ctx();
    if (multi() || n == 0 || nnz < (1ull << 21) || nnz >= 0xFFFFFFFFull) return false;
    if (const char *e = getenv("GX_UPLOAD_PIPELINE")) if (e[0] == '0') return false;
    // row-block chunks cut by entry count, read off the caller's (host) offsets
    ChunkBounds cb;
    uint64_t rows[UP_MAX_CHUNKS + 1];
    int K = (int)std::min<uint64_t>(UP_MAX_CHUNKS, nnz >> 20);
    if (rowptr[0] != 0 || rowptr[n] != nnz) return false; // the plain path reports it
    rows[0] = 0;
    cb.e[0] = 0;
    int kk = 0;
    for (int k = 1; k < K; k++) {
        // chunks shrink towards the end: what is left to do after the last byte has arrived is the last chunk's sort
        const uint64_t target = (uint64_t)((double)nnz * (1.0 - std::pow(1.0 - (double)k / K, 1.6)));
        uint64_t r = (uint64_t)(std::lower_bound(rowptr, rowptr + n + 1, target) - rowptr);
        if (r > n) r = n;
        const uint64_t e = rowptr[r];
        if (e > nnz || e < cb.e[kk] || r < rows[kk]) return false; // offsets not monotone: the plain path reports it
        if (e == cb.e[kk]) continue;
        kk++;
        rows[kk] = r;
        cb.e[kk] = e;
    }
    if (cb.e[kk] < nnz) { kk++; rows[kk] = n; cb.e[kk] = nnz; } else rows[kk] = n;
    K = kk;
    cb.k = K;
    for (int k = K + 1; k <= UP_MAX_CHUNKS; k++) cb.e[k] = nnz;
    uint64_t max_chunk = 0;
    for (int k = 0; k < K; k++) max_chunk = std::max(max_chunk, cb.e[k + 1] - cb.e[k]);

    static cudaStream_t copy_stream = nullptr;
    static cudaEvent_t ev_ready = nullptr, ev_chunk[UP_MAX_CHUNKS + 1], ev_t[3]; // ev_t: start / copies done / all done
    if (!copy_stream) {
        GX_CUDA(cudaStreamCreateWithFlags(&copy_stream, cudaStreamNonBlocking));
        GX_CUDA(cudaEventCreateWithFlags(&ev_ready, cudaEventDisableTiming));
        for (auto &e : ev_t) GX_CUDA(cudaEventCreate(&e));
        for (auto &e : ev_chunk) GX_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    }
    cudaStream_t s = c.stream;
    const int bits = bits_for(n);
    g->out.rowptr.alloc(n + 1);
    g->out.col.alloc(nnz);
    if (weights) g->out.w.alloc(nnz);
    g->in.rowptr.alloc(n + 1);
    g->in.col.alloc(nnz);
    DevBuf<uint32_t> run_keys(nnz), run_vals(nnz), chunk_rows(max_chunk);
    DevBuf<uint64_t> runstart((uint64_t)K * (n + 1)), delta((uint64_t)K * n), total(n + 1);
    DevBuf<int> flags(2);
    size_t tb_sort = 0, tb_scan = 0, tb_scan64 = 0;
    GX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb_sort, (const uint32_t *)nullptr, (uint32_t *)nullptr, (const uint32_t *)nullptr,
                                            (uint32_t *)nullptr, (int64_t)max_chunk, 0, bits, s));
    GX_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb_scan, (uint32_t *)nullptr, (uint32_t *)nullptr, MaxU32(), (int64_t)max_chunk, s));
    GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb_scan64, (uint64_t *)nullptr, (uint64_t *)nullptr, (int64_t)(n + 1), s));
    DevBuf<char> tmp(std::max(std::max(tb_sort, tb_scan), tb_scan64));
    int h[2] = {0, 0};
    try {
    flags.zero();
    GX_CUDA(cudaEventRecord(ev_t[0], s));
    GX_CUDA(cudaEventRecord(ev_ready, s)); // the copy stream must not touch the buffers before they exist
    GX_CUDA(cudaStreamWaitEvent(copy_stream, ev_ready, 0));
    GX_CUDA(cudaMemcpyAsync(g->out.rowptr.p, rowptr, (n + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, copy_stream));
    GX_CUDA(cudaEventRecord(ev_chunk[UP_MAX_CHUNKS], copy_stream));
    for (int k = 0; k < K; k++) {
        GX_CUDA(cudaMemcpyAsync(g->out.col.p + cb.e[k], colidx + cb.e[k], (cb.e[k + 1] - cb.e[k]) * sizeof(uint32_t),
                                cudaMemcpyHostToDevice, copy_stream));
        GX_CUDA(cudaEventRecord(ev_chunk[k], copy_stream));
    }
    if (weights) GX_CUDA(cudaMemcpyAsync(g->out.w.p, weights, nnz * sizeof(double), cudaMemcpyHostToDevice, copy_stream));
    GX_CUDA(cudaEventRecord(ev_t[1], copy_stream));
    GX_CUDA(cudaEventRecord(ev_ready, copy_stream)); // reused: "all copies done"
    GX_CUDA(cudaStreamWaitEvent(s, ev_chunk[UP_MAX_CHUNKS], 0));
    GX_LAUNCH(k_check_rowptr, grid_persistent(4), 256, 0, g->out.rowptr.p, n, nnz, flags.p);
    for (int k = 0; k < K; k++) {
        const uint64_t e0 = cb.e[k], e1 = cb.e[k + 1], cnt = e1 - e0;
        GX_CUDA(cudaStreamWaitEvent(s, ev_chunk[k], 0));
        GX_LAUNCH(k_check_chunk, grid_persistent(4), 256, 0, g->out.rowptr.p, g->out.col.p, n, e0, e1, flags.p, flags.p + 1);
        GX_CUDA(cudaMemsetAsync(chunk_rows.p, 0, cnt * sizeof(uint32_t), s));
        GX_LAUNCH(k_row_heads_range, grid_persistent(4), 256, 0, g->out.rowptr.p, rows[k], rows[k + 1], e0, e1, chunk_rows.p);
        GX_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb_scan, chunk_rows.p, chunk_rows.p, MaxU32(), (int64_t)cnt, s));
        GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb_sort, (const uint32_t *)(g->out.col.p + e0), run_keys.p + e0,
                                                (const uint32_t *)chunk_rows.p, run_vals.p + e0, (int64_t)cnt, 0, bits, s));
        rowptr_from_sorted_rows(run_keys.p + e0, cnt, n, runstart.p + (uint64_t)k * (n + 1)); // binary search per column
        count_launch(2);
    }
    GX_LAUNCH(k_sum_runs, grid_persistent(8), 256, 0, runstart.p, K, n, total.p, flags.p);
    GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb_scan64, total.p, g->in.rowptr.p, (int64_t)(n + 1), s));
    GX_LAUNCH(k_run_delta, grid_persistent(8), 256, 0, g->in.rowptr.p, runstart.p, K, n, delta.p, flags.p);
    GX_LAUNCH(k_merge_runs, grid_persistent(8), 256, 0, run_keys.p, run_vals.p, cb, delta.p, n, g->in.col.p, flags.p);
    GX_CUDA(cudaStreamWaitEvent(s, ev_ready, 0)); // weights (and everything else on the copy stream)
    GX_CUDA(cudaEventRecord(ev_t[2], s));
    read_back(h, flags.p, sizeof(h));
    } catch (...) {
        // the scoped buffers are released on `s`: nothing may still be writing into them from the copy stream
        cudaStreamSynchronize(copy_stream);
        cudaStreamSynchronize(s);
        throw;
    }
    {
        // h2d_ms: until the last byte arrived; build_ms: what validation + transposition add after it
        float up = 0, all = 0;
        GX_CUDA(cudaEventElapsedTime(&up, ev_t[0], ev_t[1]));
        GX_CUDA(cudaEventElapsedTime(&all, ev_t[0], ev_t[2]));
        c.timing.h2d_ms += up;
        c.timing.build_ms += all > up ? all - up : 0.0;
    }
    if (h[0]) throw Error(GX_ERR_INVALID, "CSR arrays are inconsistent (rowptr not monotone or column id >= n)");
    g->have_in = true;
    if (h[1]) {
        // jumbled rows: sort them the usual way and redo the transposition from the sorted rows
        g->have_in = false;
        g->in.rowptr.release();
        g->in.col.release();
        finish_graph(g);
        ensure_in_adj(g);
    }
    return true;
}

// ------------------------------------------------------------------------- LCC cache
// key = (a << 33) | (b << 1) | rev : rev = 0 for (a,b) in A, 1 for the mirrored copy of (b,a)
__global__ void k_lcc_keys(const uint32_t *__restrict__ row, const uint32_t *__restrict__ col, uint64_t m,
                           int directed, uint64_t *__restrict__ keys)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) {
        uint64_t a = row[e], b = col[e];
        uint64_t sentinel = ~0ull;
        if (directed) {
            keys[2 * e] = (a == b) ? sentinel : ((a << 33) | (b << 1));
            keys[2 * e + 1] = (a == b) ? sentinel : ((b << 33) | (a << 1) | 1ull);
        } else {
            keys[e] = (a == b) ? sentinel : ((a << 33) | (b << 1));
        }
    }
}

// group head of equal (a,b): emits (a << 32) | mult31 | b, mult31 set when both directions exist
__global__ void k_lcc_heads(const uint64_t *__restrict__ keys, uint64_t cnt, int directed,
                            uint64_t *__restrict__ ukeys, uint8_t *__restrict__ head)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) {
        uint64_t k = keys[i];
        bool is_head = (k != ~0ull) && (i == 0 || (keys[i - 1] >> 1) != (k >> 1));
        uint64_t out = 0;
        if (is_head) {
            bool fwd = (k & 1) == 0, rev = (k & 1) != 0;
            for (uint64_t j = i + 1; j < cnt && (keys[j] >> 1) == (k >> 1); j++) {
                if (keys[j] & 1) rev = true; else fwd = true;
            }
            uint64_t a = k >> 33, b = (k >> 1) & 0xFFFFFFFFull;
            bool both = directed ? (fwd && rev) : true;
            out = (a << 32) | (both ? (uint64_t)LCC_MULT_BIT : 0ull) | b;
        }
        ukeys[i] = out;
        head[i] = is_head ? 1 : 0;
    }
}

__global__ void k_degree_from_rowptr(const uint64_t *__restrict__ rowptr, uint64_t n, uint32_t *__restrict__ deg)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) deg[v] = (uint32_t)(rowptr[v + 1] - rowptr[v]);
}

__global__ void k_lcc_orient_flags(const uint64_t *__restrict__ ukeys, uint64_t um, const uint32_t *__restrict__ udeg,
                                   uint8_t *__restrict__ keep)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < um; i += stride) {
        uint64_t k = ukeys[i];
        uint32_t a = (uint32_t)(k >> 32), b = (uint32_t)k & ~LCC_MULT_BIT;
        uint32_t da = udeg[a], db = udeg[b];
        keep[i] = (da < db || (da == db && a < b)) ? 1 : 0;
    }
}

__global__ void k_lcc_list_bytes(const uint64_t *__restrict__ orowptr, const uint32_t *__restrict__ orow,
                                 const uint32_t *__restrict__ ocol, uint64_t om, unsigned long long *__restrict__ total)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long acc = 0;
    for (; i < om; i += stride) {
        uint32_t u = orow[i], v = ocol[i] & ~LCC_MULT_BIT;
        acc += (orowptr[u + 1] - orowptr[u]) + (orowptr[v + 1] - orowptr[v]);
    }
    acc = warp_sum(acc);
    if (lane_id() == 0 && acc) atomicAdd(total, acc);
}

uint64_t select_flagged(const uint64_t *in, const uint8_t *flags, uint64_t count, DevBuf<uint64_t> &out)
{
    DevBuf<uint64_t> nsel(1);
    size_t tb = 0;
    GX_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, in, flags, out.p, nsel.p, (int64_t)count, ctx().stream));
    DevBuf<char> tmp(tb);
    GX_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, in, flags, out.p, nsel.p, (int64_t)count, ctx().stream));
    uint64_t h = 0;
    read_back(&h, nsel.p, sizeof(h));
    return h;
}

// table size of an oriented row: 0 below LCC_TAB_MIN entries, else the power of two >= 4 * entries
__global__ void k_lcc_tab_sizes(const uint64_t *__restrict__ orowptr, uint64_t n, uint64_t per_entry, uint64_t *__restrict__ size)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v <= n; v += stride) {
        uint64_t sz = 0;
        if (v < n) {
            const uint64_t d = orowptr[v + 1] - orowptr[v];
            if (d >= LCC_TAB_MIN) { sz = 32; while (sz < per_entry * d) sz <<= 1; }
        }
        size[v] = sz;
    }
}

__global__ void k_lcc_tab_fill(const uint64_t *__restrict__ tab_off, const uint32_t *__restrict__ orow,
                               const uint32_t *__restrict__ ocol, uint64_t om, uint32_t *__restrict__ tab)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < om; e += stride) {
        const uint32_t u = orow[e];
        const uint64_t t0 = tab_off[u], tsz = tab_off[u + 1] - t0;
        if (!tsz) continue;
        const uint32_t c = ocol[e];
        uint64_t s = lcc_tab_hash(c & ~LCC_MULT_BIT, tsz);
        while (atomicCAS(&tab[t0 + s], 0xFFFFFFFFu, c) != 0xFFFFFFFFu) s = (s + 1) & (tsz - 1);
    }
}

// sort key of an oriented entry u -> v: the vertex whose list is the longer one of the intersection
__global__ void k_lcc_owner_keys(const uint64_t *__restrict__ orowptr, const uint32_t *__restrict__ orow,
                                 const uint32_t *__restrict__ ocol, uint64_t om, uint32_t *__restrict__ key, uint32_t *__restrict__ idx)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < om; e += stride) {
        const uint32_t u = orow[e], v = ocol[e] & ~LCC_MULT_BIT;
        const uint64_t du = orowptr[u + 1] - orowptr[u], dv = orowptr[v + 1] - orowptr[v];
        key[e] = du <= dv ? v : u;
        idx[e] = (uint32_t)e;
    }
}

// first-level order inside an owner's segment: by the length of the SHORTER list, longest first -- the 8-lane groups
// of a warp (and the warps of a CTA) then walk lists of similar length side by side instead of waiting for the longest
__global__ void k_lcc_len_keys(const uint64_t *__restrict__ orowptr, const uint32_t *__restrict__ orow,
                               const uint32_t *__restrict__ ocol, uint64_t om, uint32_t *__restrict__ key, uint32_t *__restrict__ idx)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < om; e += stride) {
        const uint32_t u = orow[e], v = ocol[e] & ~LCC_MULT_BIT;
        const uint64_t du = orowptr[u + 1] - orowptr[u], dv = orowptr[v + 1] - orowptr[v];
        const uint64_t sh = du <= dv ? du : dv;
        key[e] = 0xFFFFu - (uint32_t)(sh > 0xFFFFull ? 0xFFFFull : sh);
        idx[e] = (uint32_t)e;
    }
}

// owner (the vertex whose list is the longer one) of the entries in the order idx[]
__global__ void k_lcc_owner_of(const uint64_t *__restrict__ orowptr, const uint32_t *__restrict__ orow,
                               const uint32_t *__restrict__ ocol, const uint32_t *__restrict__ idx, uint64_t om, uint32_t *__restrict__ key)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < om; i += stride) {
        const uint32_t e = idx[i];
        const uint32_t u = orow[e], v = ocol[e] & ~LCC_MULT_BIT;
        const uint64_t du = orowptr[u + 1] - orowptr[u], dv = orowptr[v + 1] - orowptr[v];
        key[i] = du <= dv ? v : u;
    }
}

__global__ void k_lcc_permute(const uint32_t *__restrict__ idx, const uint32_t *__restrict__ orow, const uint32_t *__restrict__ ocol,
                              uint64_t om, uint32_t *__restrict__ eu, uint32_t *__restrict__ ev)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < om; e += stride) { const uint32_t s = idx[e]; eu[e] = orow[s]; ev[e] = ocol[s]; }
}

void ensure_lcc_cache(gx_graph *g)
{
    if (g->have_lcc) return;
    PhaseTimer t(&ctx().timing.build_ms);
    const uint64_t n = g->n, m = g->m;
    GX_REQUIRE(n < 0x7FFFFFFFull, "LCC needs n < 2^31");
    g->udeg.alloc(n);
    g->orowptr.alloc(n + 1);
    const uint64_t cnt = g->directed ? 2 * m : m;
    if (cnt == 0) {
        g->udeg.zero();
        g->orowptr.zero();
        g->om = 0;
        g->have_lcc = true;
        return;
    }
    DevBuf<uint64_t> ukeys_sel;
    uint64_t um = 0;
    {
        DevBuf<uint64_t> keys(cnt);
        {
            DevBuf<uint32_t> rows(m);
            expand_row_ids(g->out.rowptr.p, n, m, rows.p);
            GX_LAUNCH(k_lcc_keys, grid_persistent(8), 256, 0, rows.p, g->out.col.p, m, g->directed ? 1 : 0, keys.p);
        }
        // an undirected store is already sorted by (row, col); a directed one needs the merge sort
        if (g->directed) sort_keys64(keys, cnt, 64);
        DevBuf<uint64_t> ukeys(cnt);
        DevBuf<uint8_t> head(cnt);
        GX_LAUNCH(k_lcc_heads, grid_persistent(8), 256, 0, keys.p, cnt, g->directed ? 1 : 0, ukeys.p, head.p);
        keys.release();
        ukeys_sel.alloc(cnt);
        um = select_flagged(ukeys.p, head.p, cnt, ukeys_sel);
    }
    // degrees in U, then keep only low -> high (degree, id) entries
    DevBuf<uint64_t> urowptr(n + 1);
    rowptr_from_sorted_keys(ukeys_sel.p, um, n, urowptr.p);
    GX_LAUNCH(k_degree_from_rowptr, grid_for(n, 256), 256, 0, urowptr.p, n, g->udeg.p);
    DevBuf<uint8_t> keep(um ? um : 1);
    DevBuf<uint64_t> okeys(um ? um : 1);
    uint64_t om = 0;
    if (um) {
        GX_LAUNCH(k_lcc_orient_flags, grid_persistent(8), 256, 0, ukeys_sel.p, um, g->udeg.p, keep.p);
        om = select_flagged(ukeys_sel.p, keep.p, um, okeys);
    }
    g->om = om;
    g->ocol.alloc(om ? om : 1);
    g->orow.alloc(om ? om : 1);
    rowptr_from_sorted_keys(okeys.p, om, n, g->orowptr.p);
    if (om) {
        GX_LAUNCH(k_low32, grid_persistent(8), 256, 0, okeys.p, om, g->ocol.p);
        GX_LAUNCH(k_high32, grid_persistent(8), 256, 0, okeys.p, om, g->orow.p);
        DevBuf<unsigned long long> total(1);
        total.zero();
        GX_LAUNCH(k_lcc_list_bytes, grid_persistent(8), 256, 0, g->orowptr.p, g->orow.p, g->ocol.p, om, total.p);
        unsigned long long h = 0;
        read_back(&h, total.p, sizeof(h));
        g->lcc_list_bytes = 4ull * h;
        // membership tables of the longer rows
        g->ltab_off.alloc(n + 1);
        uint64_t per_entry = 4; // slots per entry: load <= 1/4
        if (const char *e = getenv("GX_LCC_SLOTS")) per_entry = atoi(e) >= 2 ? (uint64_t)atoi(e) : 2; // tuning knob
        GX_LAUNCH(k_lcc_tab_sizes, grid_persistent(8), 256, 0, g->orowptr.p, n, per_entry, g->ltab_off.p);
        size_t tb = 0;
        GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, g->ltab_off.p, g->ltab_off.p, (int64_t)(n + 1), ctx().stream));
        DevBuf<char> tmp(tb);
        GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, g->ltab_off.p, g->ltab_off.p, (int64_t)(n + 1), ctx().stream));
        uint64_t slots = 0;
        read_back(&slots, g->ltab_off.p + n, sizeof(slots));
        g->ltab.alloc(slots ? slots : 1);
        g->ltab.fill_byte(0xFF);
        if (slots) GX_LAUNCH(k_lcc_tab_fill, grid_persistent(8), 256, 0, g->ltab_off.p, g->orow.p, g->ocol.p, om, g->ltab.p);
        // entries ordered by the owner of the longer list (stable: ties stay in (u, v) order)
        GX_REQUIRE(om < 0xFFFFFFFFull, "LCC needs fewer than 2^32 oriented entries");
        g->lcc_eu.alloc(om);
        g->lcc_ev.alloc(om);
        DevBuf<uint32_t> key(om), key_alt(om), idx(om), idx_alt(om);
        cub::DoubleBuffer<uint32_t> dk(key.p, key_alt.p), dv(idx.p, idx_alt.p);
        size_t tb2 = 0;
        GX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb2, dk, dv, (int64_t)om, 0, bits_for(n), ctx().stream));
        DevBuf<char> tmp2(tb2);
        // two stable sorts: by the shorter list's length (16 bits), then by owner -- inside an owner's segment the
        // entries stay ordered by length
        GX_LAUNCH(k_lcc_len_keys, grid_persistent(8), 256, 0, g->orowptr.p, g->orow.p, g->ocol.p, om, dk.Current(), dv.Current());
        GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, tb2, dk, dv, (int64_t)om, 0, 16, ctx().stream));
        GX_LAUNCH(k_lcc_owner_of, grid_persistent(8), 256, 0, g->orowptr.p, g->orow.p, g->ocol.p, dv.Current(), om, dk.Current());
        GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp2.p, tb2, dk, dv, (int64_t)om, 0, bits_for(n), ctx().stream));
        GX_LAUNCH(k_lcc_permute, grid_persistent(8), 256, 0, dv.Current(), g->orow.p, g->ocol.p, om, g->lcc_eu.p, g->lcc_ev.p);
        // the sorted keys are kept: the counting kernel finds the owner segments of a run by comparing them
        g->lcc_owner.alloc(om);
        GX_CUDA(cudaMemcpyAsync(g->lcc_owner.p, dk.Current(), om * sizeof(uint32_t), cudaMemcpyDeviceToDevice, ctx().stream));
        GX_CUDA(cudaStreamSynchronize(ctx().stream)); // scoped sort buffers
    } else {
        g->lcc_eu.alloc(1);
        g->lcc_ev.alloc(1);
        g->lcc_owner.alloc(1);
        g->ltab_off.alloc(n + 1);
        g->ltab_off.zero();
        g->ltab.alloc(1);
    }
    g->have_lcc = true;
}

// ------------------------------------------------------------------------- max degree vertex
__global__ void k_max_degree(const uint64_t *__restrict__ rowptr, uint64_t n, unsigned long long *__restrict__ best)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long b = 0;
    for (; v < n; v += stride) {
        // (degree, ~id) packed so that max picks the largest degree, then the smallest id
        unsigned long long d = rowptr[v + 1] - rowptr[v];
        unsigned long long key = (d << 32) | (0xFFFFFFFFull - v);
        b = key > b ? key : b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long x = __shfl_xor_sync(FULL, b, o);
        b = x > b ? x : b;
    }
    if (lane_id() == 0) atomicMax(best, b);
}

} // namespace gx

using namespace gx;

void gx_cdlp_plan_free(void *p);
void gx_pr_cache_free(void *p);
void gx_sssp_cache_free(void *p);
gx_graph::~gx_graph()
{
    if (sssp_cache) gx_sssp_cache_free(sssp_cache);
    if (cdlp_plan) gx_cdlp_plan_free(cdlp_plan);
    if (pr_cache) gx_pr_cache_free(pr_cache);
}

// ------------------------------------------------------------------------- C ABI
extern "C" int gx_graph_create_csr(gx_graph **out, uint64_t n, uint64_t nnz, const uint64_t *rowptr,
                                   const uint64_t *colidx, const double *weights, int directed)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(out != nullptr, "graph handle is NULL");
        GX_REQUIRE(colidx != nullptr || nnz == 0, "colidx is NULL");
        ctx().timing = gx_timing{};
        gx_graph *g = new gx_graph();
        try {
            {
                PhaseTimer t(&ctx().timing.h2d_ms);
                upload_common(g, n, nnz, rowptr, weights, directed);
                // GrB_Index (uint64) column ids: staged in chunks and narrowed to 4 bytes on the device
                // (several GPUs: a rank stages and narrows its slice only, the 4-byte ids are all-gathered)
                const uint64_t CH = 1ull << 26;
                const Partition part = make_even_partition(nnz, 64);
                DevBuf<uint64_t> stage(nnz < CH ? nnz : CH);
                DevBuf<int> bad(1);
                bad.zero();
                for (uint64_t o = part.lo; o < part.hi; o += CH) {
                    uint64_t c = part.hi - o < CH ? part.hi - o : CH;
                    GX_CUDA(cudaMemcpyAsync(stage.p, colidx + o, c * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx().stream));
                    GX_LAUNCH(k_narrow_u64, grid_persistent(8), 256, 0, stage.p, g->out.col.p + o, c, n, bad.p);
                }
                if (multi() && nnz) {
                    allgatherv(g->out.col.p, Dt::U32, part);
                    allreduce(bad.p, 1, Dt::I32, Red::Max);
                }
                if (nnz && read_flag(bad.p)) throw Error(GX_ERR_INVALID, "column id >= n");
            }
            {
                PhaseTimer t(&ctx().timing.build_ms);
                finish_graph(g);
            }
        } catch (...) {
            delete g;
            throw;
        }
        *out = g;
    });
}

extern "C" int gx_graph_create_csr32(gx_graph **out, uint64_t n, uint64_t nnz, const uint64_t *rowptr,
                                     const uint32_t *colidx, const double *weights, int directed)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(out != nullptr, "graph handle is NULL");
        GX_REQUIRE(colidx != nullptr || nnz == 0, "colidx is NULL");
        ctx().timing = gx_timing{};
        gx_graph *g = new gx_graph();
        try {
            {
                PhaseTimer t(&ctx().timing.h2d_ms);
                upload_common(g, n, nnz, rowptr, weights, directed);
                upload_array(g->out.col.p, colidx, nnz, Dt::U32);
            }
            {
                PhaseTimer t(&ctx().timing.build_ms);
                finish_graph(g);
            }
        } catch (...) {
            delete g;
            throw;
        }
        *out = g;
    });
}

extern "C" int gx_graph_create_csr32_cached(gx_graph **out, uint64_t n, uint64_t nnz, const uint64_t *rowptr,
                                            const uint32_t *colidx, const double *weights, int directed, unsigned cache)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(out != nullptr, "graph handle is NULL");
        GX_REQUIRE(colidx != nullptr || nnz == 0, "colidx is NULL");
        GX_REQUIRE(rowptr != nullptr || n == 0, "rowptr is NULL");
        GX_REQUIRE(n < 0xFFFFFFFEull, "n must be < 2^32 - 2");
        ctx().timing = gx_timing{};
        gx_graph *g = new gx_graph();
        try {
            g->n = n;
            g->m = nnz;
            g->directed = directed != 0;
            g->weighted = weights != nullptr;
            bool done = false;
            if (g->directed && (cache & GX_CACHE_AT)) // upload with the transposition riding along
                done = upload_transpose_pipelined(g, n, nnz, rowptr, colidx, weights);
            if (!done) {
                PhaseLog pl;
                {
                    PhaseTimer t(&ctx().timing.h2d_ms);
                    upload_common(g, n, nnz, rowptr, weights, directed);
                    upload_array(g->out.col.p, colidx, nnz, Dt::U32);
                }
                pl.mark("upload + all-gather");
                {
                    PhaseTimer t(&ctx().timing.build_ms);
                    finish_graph(g);
                }
                pl.mark("validation");
                if (cache & GX_CACHE_AT) ensure_in_adj(g); // (times itself into build_ms)
                pl.mark("transposition");
            }
            if (cache & GX_CACHE_LCC) {
                PhaseTimer t(&ctx().timing.build_ms);
                ensure_lcc_cache(g);
            }
            GX_CUDA(cudaStreamSynchronize(ctx().stream));
        } catch (...) {
            delete g;
            throw;
        }
        *out = g;
    });
}

extern "C" int gx_graph_free(gx_graph *g)
{
    return guarded([&] {
        if (!g) return;
        delete g;
        if (ctx().ready) GX_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}

extern "C" int gx_graph_info(const gx_graph *g, uint64_t *n, uint64_t *nnz, int *directed, int *weighted)
{
    return guarded([&] {
        GX_REQUIRE(g != nullptr, "graph is NULL");
        if (n) *n = g->n;
        if (nnz) *nnz = g->m;
        if (directed) *directed = g->directed;
        if (weighted) *weighted = g->weighted;
    });
}

extern "C" int gx_graph_cache(gx_graph *g, unsigned what)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        ctx().timing = gx_timing{};
        if (what & GX_CACHE_AT) ensure_in_adj(g);
        if (what & GX_CACHE_LCC) ensure_lcc_cache(g);
        GX_CUDA(cudaStreamSynchronize(ctx().stream));
    });
}

extern "C" int gx_graph_download(const gx_graph *g, uint64_t *rowptr, uint32_t *colidx, double *weights)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        cudaStream_t s = ctx().stream;
        if (rowptr) GX_CUDA(cudaMemcpyAsync(rowptr, g->out.rowptr.p, (g->n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        if (colidx && g->m) GX_CUDA(cudaMemcpyAsync(colidx, g->out.col.p, g->m * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        if (weights && g->weighted && g->m)
            GX_CUDA(cudaMemcpyAsync(weights, g->out.w.p, g->m * sizeof(double), cudaMemcpyDeviceToHost, s));
        GX_CUDA(cudaStreamSynchronize(s));
    });
}

extern "C" int gx_graph_max_degree_vertex(const gx_graph *g, uint64_t *v)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr && v != nullptr, "NULL argument");
        GX_REQUIRE(g->n > 0, "empty graph");
        DevBuf<unsigned long long> best(1);
        best.zero();
        GX_LAUNCH(k_max_degree, grid_persistent(4), 256, 0, g->out.rowptr.p, g->n, best.p);
        unsigned long long h = 0;
        read_back(&h, best.p, sizeof(h));
        *v = 0xFFFFFFFFull - (h & 0xFFFFFFFFull);
    });
}
