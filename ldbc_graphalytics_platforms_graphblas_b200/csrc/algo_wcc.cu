// algo_wcc.cu -- weakly connected components with FastSV.  Replaces
// WeaklyConnectedComponents (wcc.cpp:39-66): A = A v A' for directed input
// (the in-edge adjacency is streamed as a second pass instead of materialising
// the union), then LAGr_ConnectedComponents (FastSV): parent f, grandparent gp,
//   mngp(u) = min_{v in N(u)} gp(v);  f[f[u]] min= mngp(u);  f[u] min= mngp(u), gp(u);
//   gp = f[f];  stop when an iteration changes nothing.
// All updates are monotone minima inside one component, so the fix-point is the
// smallest dense id of every component -- exactly what LAGraph returns
// (wcc.cpp:31-34 prints it unmapped), independent of the update order.
//
//   k_wcc_hook        sub-warp group per row (<= ROW_SPLIT entries)
//   k_wcc_hook_chunk  one CTA per CHUNK entries of a long row
//   k_wcc_shortcut    gp = f[f], path halving, change detection
// Algorithmic bytes per iteration: 4 m_sym + 8(n+1) + 6*4n.
#include "graph.cuh"

namespace gx {

constexpr int WCC_G = 8;

__global__ void k_wcc_init(uint32_t *__restrict__ f, uint32_t *__restrict__ gp, uint64_t n)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) { f[v] = (uint32_t)v; gp[v] = (uint32_t)v; }
}

__device__ __forceinline__ bool wcc_apply(uint32_t *f, uint32_t u, uint32_t mn)
{
    bool ch = false;
    const uint32_t p = f[u];
    if (mn < p) {
        // stochastic hooking on the parent, aggressive hooking on u itself
        if (atomicMin(&f[p], mn) > mn) ch = true;
        if (atomicMin(&f[u], mn) > mn) ch = true;
    } else if (mn < f[p]) {
        if (atomicMin(&f[p], mn) > mn) ch = true;
    }
    return ch;
}

__global__ void __launch_bounds__(256)
k_wcc_hook(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t v0, uint64_t v1,
           const uint32_t *__restrict__ gp, uint32_t *__restrict__ f, int *__restrict__ changed)
{
    const unsigned sub = threadIdx.x & (WCC_G - 1);
    uint64_t grp = v0 + ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / WCC_G;
    const uint64_t ngrp = ((uint64_t)gridDim.x * blockDim.x) / WCC_G;
    const uint64_t trips = (v1 - v0 + ngrp - 1) / ngrp;
    bool ch = false;
    for (uint64_t t = 0; t < trips; t++, grp += ngrp) {
        const bool live = grp < v1;
        uint64_t a = 0, b = 0;
        if (live) { a = rowptr[grp]; b = rowptr[grp + 1]; }
        const bool is_short = live && b > a && (b - a) <= ROW_SPLIT;
        uint32_t mn = 0xFFFFFFFFu;
        if (is_short) {
#pragma unroll 4
            for (uint64_t e = a + sub; e < b; e += WCC_G) mn = min(mn, gp[ld_stream(col + e)]);
        }
#pragma unroll
        for (int o = WCC_G / 2; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(FULL, mn, o));
        if (is_short && sub == 0) ch |= wcc_apply(f, (uint32_t)grp, mn);
    }
    if (ch) *changed = 1;
}

__global__ void __launch_bounds__(256)
k_wcc_hook_chunk(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
                 const uint32_t *__restrict__ chunk_row, const uint64_t *__restrict__ chunk_begin,
                 const uint32_t *__restrict__ gp, uint32_t *__restrict__ f, int *__restrict__ changed)
{
    const uint32_t c = blockIdx.x;
    const uint32_t u = chunk_row[c];
    const uint64_t b0 = chunk_begin[c];
    const uint64_t row_end = rowptr[u + 1];
    const uint64_t e_end = (b0 + CHUNK < row_end) ? b0 + CHUNK : row_end;
    uint32_t mn = 0xFFFFFFFFu;
#pragma unroll 8
    for (uint64_t e = b0 + threadIdx.x; e < e_end; e += 256) mn = min(mn, gp[ld_stream(col + e)]);
    __shared__ uint32_t red[8];
    mn = warp_min_u32(mn);
    if (lane_id() == 0) red[threadIdx.x >> 5] = mn;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; i++) mn = min(mn, red[i]);
        if (wcc_apply(f, u, mn)) *changed = 1;
    }
}

// f_prev (multi-GPU only): parents before this iteration's hooking, so that a change brought in by
// the min-reduction over the ranks' replicas also counts as a change.
__global__ void k_wcc_shortcut(uint32_t *__restrict__ f, uint32_t *__restrict__ gp, const uint32_t *__restrict__ f_prev,
                               uint64_t n, int *__restrict__ changed)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool ch = false;
    for (; v < n; v += stride) {
        const uint32_t fu = f[v];
        if (f_prev && fu != f_prev[v]) ch = true;
        const uint32_t g = f[fu]; // new grandparent
        if (g != gp[v]) { gp[v] = g; ch = true; }
        if (g < fu) { atomicMin(&f[v], g); ch = true; }
    }
    if (ch) *changed = 1;
}

__global__ void k_widen_u32(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) out[v] = in[v];
}

static void wcc_hook_pass(Adj &a, uint64_t n, const uint32_t *gp, uint32_t *f, int *changed)
{
    const RowPlan &p = a.plan;
    if (p.n_chunks)
        GX_LAUNCH(k_wcc_hook_chunk, (unsigned)p.n_chunks, 256, 0, a.rowptr.p, a.col.p, p.chunk_row.p, p.chunk_begin.p, gp, f, changed);
    GX_LAUNCH(k_wcc_hook, grid_persistent(8), 256, 0, a.rowptr.p, a.col.p, p.part.lo, p.part.hi, gp, f, changed);
    (void)n;
}

} // namespace gx

using namespace gx;

extern "C" int gx_wcc(gx_graph *g, uint64_t *comp_host)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        Context &c = ctx();
        c.timing = gx_timing{};
        const uint64_t n = g->n;
        if (n == 0) return;
        ensure_in_adj(g);
        {
            PhaseTimer tb(&c.timing.build_ms);
            ensure_plan(g->out, n);
            if (g->directed) ensure_plan(g->in, n);
        }
        const uint64_t m_sym = g->directed ? 2 * g->m : g->m;
        g->res_u64.alloc(n);
        DevBuf<uint32_t> f(n), gp(n), f_prev(multi() ? n : 0);
        DevBuf<int> changed(1);
        uint32_t iters = 0;
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            GX_LAUNCH(k_wcc_init, grid_persistent(8), 256, 0, f.p, gp.p, n);
            for (;;) {
                changed.zero();
                if (multi()) GX_CUDA(cudaMemcpyAsync(f_prev.p, f.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
                wcc_hook_pass(g->out, n, gp.p, f.p, changed.p);
                if (g->directed) wcc_hook_pass(g->in, n, gp.p, f.p, changed.p);
                if (multi()) {
                    // every rank hooked with the rows of its block on its own replica: combine by min
                    allreduce(f.p, n, Dt::U32, Red::Min);
                }
                GX_LAUNCH(k_wcc_shortcut, grid_persistent(8), 256, 0, f.p, gp.p, multi() ? f_prev.p : nullptr, n, changed.p);
                // replicas may shortcut in different orders; the ranks stop together, once nobody changed anything
                allreduce(changed.p, 1, Dt::I32, Red::Max);
                iters++;
                int h = 0;
                read_back(&h, changed.p, sizeof(h));
                if (!h) break;
            }
            GX_LAUNCH(k_widen_u32, grid_persistent(8), 256, 0, f.p, n, g->res_u64.p);
        }
        c.timing.iterations = iters;
        c.timing.edges_inspected = m_sym * iters;
        c.timing.algorithmic_bytes = (uint64_t)iters * (4 * m_sym + 8 * (n + 1) + 24 * n);
        if (comp_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(comp_host, g->res_u64.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
