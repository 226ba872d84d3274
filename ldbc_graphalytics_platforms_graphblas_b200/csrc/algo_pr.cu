// algo_pr.cu -- PageRank as a pull plus.second SpMV over the in-edge adjacency
// with dangling-node correction.  Replaces LA_PR (pr.cpp:47-66) ->
// LAGr_PageRankGX(&r, &iters, G, (float)damping, itermax): exactly `iters`
// iterations, FP64, r0 = 1/n, d = outdeg / damping,
//   teleport' = (1-damping)/n + damping/n * sum_{sinks} r,   w = r ./ d,
//   r = teleport' + A' (plus.second) w.
//
// Kernels per iteration (all HBM/L2-gather bound, no tensor-core work):
//   k_pr_short  sub-warp group per row (<= ROW_SPLIT in-edges): streams col ids,
//               gathers w[col], fused epilogue r -> w' = r/d and the sink sum
//   k_pr_chunk  one CTA per CHUNK entries of a long row -> partial sums
//   k_pr_long   one warp per long row: ordered sum of its partials + epilogue
//   k_pr_tele   folds the per-CTA sink partials into next iteration's teleport
// Sums are combined in a fixed order, so results are bit-reproducible.
// Algorithmic bytes per iteration: 4m (col) + 8(n+1) (rowptr) + 8n (d) + 8n (w'),
// gathers of w (8 B each, served by the 126 MB L2 while 8n fits).
#include "graph.cuh"

namespace gx {

constexpr int PR_G = 8; // lanes per short row

__global__ void k_pr_init(const uint64_t *__restrict__ out_rowptr, uint64_t n, uint64_t v0, uint64_t v1, double damping,
                          double *__restrict__ d, double *__restrict__ w, double *__restrict__ sink_part)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const double r0 = 1.0 / (double)n;
    double sink = 0.0;
    for (; v < n; v += stride) {
        uint64_t od = out_rowptr[v + 1] - out_rowptr[v];
        // d and w0 are replicated on every rank; the sink mass is counted for owned rows only
        if (od == 0) { d[v] = 0.0; w[v] = 0.0; if (v >= v0 && v < v1) sink += r0; }
        else { double dv = (double)od / damping; d[v] = dv; w[v] = r0 / dv; }
    }
    __shared__ double red[32];
    sink = warp_sum(sink);
    if (lane_id() == 0) red[threadIdx.x >> 5] = sink;
    __syncthreads();
    if (threadIdx.x < 32) {
        double x = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        x = warp_sum(x);
        if (threadIdx.x == 0) sink_part[blockIdx.x] = x;
    }
}

// sum of the per-CTA sink partials; one CTA, fixed order (all-reduced across ranks afterwards)
__global__ void k_pr_tele(const double *__restrict__ sink_part, unsigned nparts, double *__restrict__ sink_sum)
{
    __shared__ double red[32];
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < nparts; i += blockDim.x) s += sink_part[i];
    s = warp_sum(s);
    if (lane_id() == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double x = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        x = warp_sum(x);
        if (threadIdx.x == 0) *sink_sum = x;
    }
}

__device__ __forceinline__ void pr_epilogue(uint64_t v, double s, double tele, const double *__restrict__ d,
                                            double *__restrict__ w_new, double *__restrict__ rank, double &sink)
{
    double r = tele + s;
    double dv = d[v];
    if (dv == 0.0) { sink += r; w_new[v] = 0.0; }
    else w_new[v] = r / dv;
    if (rank) rank[v] = r;
}

struct PrScalars { double teleport, damping, n; };

__global__ void __launch_bounds__(256)
k_pr_short(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t v0, uint64_t v1,
           const double *__restrict__ w, const double *__restrict__ d, const double *__restrict__ sink_sum, PrScalars sc,
           double *__restrict__ w_new, double *__restrict__ rank, double *__restrict__ sink_part)
{
    const double tele = sc.teleport + sc.damping * *sink_sum / sc.n;
    const unsigned sub = threadIdx.x & (PR_G - 1);
    uint64_t grp = v0 + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / PR_G;
    const uint64_t ngrp = ((uint64_t)gridDim.x * blockDim.x) / PR_G;
    double sink = 0.0;
    // every lane of a warp runs the same number of trips, so the shuffles are convergent
    const uint64_t trips = (v1 - v0 + ngrp - 1) / ngrp;
    for (uint64_t t = 0; t < trips; t++, grp += ngrp) {
        const bool live = grp < v1;
        uint64_t a = 0, b = 0;
        if (live) { a = rowptr[grp]; b = rowptr[grp + 1]; }
        const bool is_short = live && (b - a) <= ROW_SPLIT;
        double s = 0.0;
        if (is_short) {
#pragma unroll 4
            for (uint64_t e = a + sub; e < b; e += PR_G) s += w[ld_stream(col + e)];
        }
#pragma unroll
        for (int o = PR_G / 2; o > 0; o >>= 1) s += __shfl_xor_sync(FULL, s, o);
        if (is_short && sub == 0) pr_epilogue(grp, s, tele, d, w_new, rank, sink);
    }
    __shared__ double red[32];
    sink = warp_sum(sink);
    if (lane_id() == 0) red[threadIdx.x >> 5] = sink;
    __syncthreads();
    if (threadIdx.x < 32) {
        double x = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        x = warp_sum(x);
        if (threadIdx.x == 0) sink_part[blockIdx.x] = x;
    }
}

// one CTA (256 threads) per chunk of CHUNK entries
__global__ void __launch_bounds__(256)
k_pr_chunk(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const uint32_t *__restrict__ chunk_row,
           const uint64_t *__restrict__ chunk_begin, const double *__restrict__ w, double *__restrict__ partial)
{
    const uint32_t c = blockIdx.x;
    const uint64_t b = chunk_begin[c];
    const uint64_t row_end = rowptr[chunk_row[c] + 1];
    const uint64_t e_end = (b + CHUNK < row_end) ? b + CHUNK : row_end;
    double s = 0.0;
#pragma unroll 8
    for (uint64_t e = b + threadIdx.x; e < e_end; e += 256) s += w[ld_stream(col + e)];
    __shared__ double red[8];
    s = warp_sum(s);
    if (lane_id() == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0;
#pragma unroll
        for (int i = 0; i < 8; i++) x += red[i];
        partial[c] = x;
    }
}

// one warp per long row
__global__ void __launch_bounds__(256)
k_pr_long(const uint32_t *__restrict__ long_rows, const uint32_t *__restrict__ first_chunk, uint64_t n_long,
          const double *__restrict__ partial, const double *__restrict__ d, const double *__restrict__ sink_sum,
          PrScalars sc, double *__restrict__ w_new, double *__restrict__ rank, double *__restrict__ sink_part)
{
    const double tele = sc.teleport + sc.damping * *sink_sum / sc.n;
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    double sink = 0.0;
    if (wid < n_long) {
        uint32_t c0 = first_chunk[wid], c1 = first_chunk[wid + 1];
        double s = 0.0;
        for (uint32_t c = c0 + lane_id(); c < c1; c += 32) s += partial[c];
        s = warp_sum(s);
        if (lane_id() == 0) pr_epilogue(long_rows[wid], s, tele, d, w_new, rank, sink);
    }
    __shared__ double red[8];
    if (lane_id() == 0) red[threadIdx.x >> 5] = sink;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0;
        for (unsigned i = 0; i < (blockDim.x >> 5); i++) x += red[i];
        sink_part[blockIdx.x] = x;
    }
}

__global__ void k_fill_f64(double *__restrict__ p, uint64_t n, double v)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}

} // namespace gx

using namespace gx;

extern "C" int gx_pagerank(gx_graph *g, double damping_in, int iters, double *rank_host)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        GX_REQUIRE(iters >= 0, "negative iteration count");
        Context &c = ctx();
        c.timing = gx_timing{};
        const uint64_t n = g->n, m = g->m;
        if (n == 0) return;
        const double damping = (double)(float)damping_in; // LAGr_PageRankGX takes `float damping`
        ensure_in_adj(g);
        Adj &in = g->in_adj();
        {
            PhaseTimer tb(&c.timing.build_ms);
            ensure_plan(in, n);
        }
        const RowPlan &plan = in.plan;
        const uint64_t v0 = plan.part.lo, v1 = plan.part.hi;
        g->res_f64.alloc(n);
        DevBuf<double> d(n), w0(n), w1(n), sink_sum(1);
        const unsigned g_short = grid_persistent(8);
        const unsigned g_long = plan.n_long ? grid_for(plan.n_long * 32, 256) : 0;
        const unsigned g_init = grid_persistent(4);
        const unsigned nparts = (g_short + g_long > g_init) ? g_short + g_long : g_init;
        DevBuf<double> sink_part(nparts), partial(plan.n_chunks ? plan.n_chunks : 1);
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            // teleport' = (1-damping)/n + (damping/n) * sum_{sinks} r   (LAGr_PageRankGX)
            const PrScalars sc{(1.0 - damping) / (double)n, damping, (double)n};
            sink_part.zero();
            GX_LAUNCH(k_pr_init, g_init, 256, 0, g->out.rowptr.p, n, v0, v1, damping, d.p, w0.p, sink_part.p);
            if (iters == 0) GX_LAUNCH(k_fill_f64, grid_persistent(4), 256, 0, g->res_f64.p, n, 1.0 / (double)n);
            double *w_old = w0.p, *w_new = w1.p;
            for (int it = 0; it < iters; it++) {
                GX_LAUNCH(k_pr_tele, 1, 256, 0, sink_part.p, nparts, sink_sum.p);
                allreduce(sink_sum.p, 1, Dt::F64, Red::Sum);
                double *rank = (it == iters - 1) ? g->res_f64.p : nullptr;
                if (plan.n_chunks)
                    GX_LAUNCH(k_pr_chunk, (unsigned)plan.n_chunks, 256, 0, in.rowptr.p, in.col.p, plan.chunk_row.p,
                              plan.chunk_begin.p, w_old, partial.p);
                GX_LAUNCH(k_pr_short, g_short, 256, 0, in.rowptr.p, in.col.p, v0, v1, w_old, d.p, sink_sum.p, sc, w_new, rank,
                          sink_part.p);
                if (g_long)
                    GX_LAUNCH(k_pr_long, g_long, 256, 0, plan.long_rows.p, plan.long_first_chunk.p, plan.n_long, partial.p,
                              d.p, sink_sum.p, sc, w_new, rank, sink_part.p + g_short);
                // the ranks exchange the owned slices of the new w (and of r after the last iteration)
                if (it + 1 < iters) allgatherv(w_new, Dt::F64, plan.part);
                else allgatherv(rank, Dt::F64, plan.part);
                double *t = w_old; w_old = w_new; w_new = t;
            }
        }
        c.timing.iterations = (uint32_t)iters;
        c.timing.edges_inspected = m * (uint64_t)iters;
        c.timing.algorithmic_bytes = (uint64_t)iters * (4 * m + 8 * (n + 1) + 28 * n);
        if (rank_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(rank_host, g->res_f64.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
