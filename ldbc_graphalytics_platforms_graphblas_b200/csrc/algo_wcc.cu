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
// Like LAGraph's LG_CC_FastSV6 the run starts on a SAMPLE of the edges and then skips the
// giant component:
//   1. two rounds: k_wcc_sample_link unites every vertex with the r-th stored entry of its row
//      (lock-free root hooking under the smaller id), k_wcc_compress turns the forest into stars;
//   2. k_wcc_mode takes the most frequent root L among 1024 sampled vertices;
//   3. S = {v : root(v) = L} is frozen.  Edges inside S are settled (one tree for good), so only
//      the rows of the vertices outside S are hooked from now on -- both ways per entry, because
//      the S side of an edge never looks at it again -- with a shortcut over all vertices per
//      iteration, until an iteration changes nothing.
// On RMAT the sample already connects > 99 % of the vertices, so the full adjacency is read
// for a few thousand rows instead of 4-5 times for all of them.  If more than a quarter of the
// vertices stay outside S (no giant component) every row is hooked as in plain FastSV:
//   k_wcc_hook        sub-warp group per row (<= ROW_SPLIT entries)
//   k_wcc_hook_chunk  one CTA per CHUNK entries of a long row
//   k_wcc_shortcut    gp = f[f], path halving, change detection
// Algorithmic bytes per FULL pass: 4 m_sym + 8(n+1) + 6*4n; the sampled run
// reports the entries it actually read.
#include <cstdlib>

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

// BOTH: every entry (u, v) also hooks v with gp(u) -- a directed graph whose in-edge adjacency is not cached is
// covered by its out-entries alone (an edge of A v A' is stored in at least one of the two rows)
template <bool BOTH>
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
            const uint32_t gu = BOTH ? gp[grp] : 0u;
#pragma unroll 4
            for (uint64_t e = a + sub; e < b; e += WCC_G) {
                const uint32_t v = ld_stream(col + e);
                const uint32_t gv = gp[v];
                mn = min(mn, gv);
                if (BOTH && gu < gv) ch |= wcc_apply(f, v, gu);
            }
        }
#pragma unroll
        for (int o = WCC_G / 2; o > 0; o >>= 1) mn = min(mn, __shfl_xor_sync(FULL, mn, o));
        if (is_short && sub == 0) ch |= wcc_apply(f, (uint32_t)grp, mn);
    }
    if (ch) *changed = 1;
}

template <bool BOTH>
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
    const uint32_t gu = BOTH ? gp[u] : 0u;
    bool chb = false;
#pragma unroll 8
    for (uint64_t e = b0 + threadIdx.x; e < e_end; e += 256) {
        const uint32_t v = ld_stream(col + e);
        const uint32_t gv = gp[v];
        mn = min(mn, gv);
        if (BOTH && gu < gv) chb |= wcc_apply(f, v, gu);
    }
    if (chb) *changed = 1;
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


// ---- sampled start (LG_CC_FastSV6) ----------------------------------------------------------
// Lock-free union of the trees of u and v (the hooking step of Afforest, Sutton et al. 2018):
// only roots are hooked, always under a smaller id, so parents keep decreasing towards the
// smallest id of the tree and the FastSV iterations that follow start from a valid forest.
// Parents are read through L1 (ld.global.ca): a stale parent is still an ancestor -- ancestors stay
// ancestors for good -- so walking stale values is valid, only the hook itself must be atomic, and a
// failed CAS hands back the fresh parent.  Reading the few hot roots from L1 instead of from one
// L2 slice is what keeps the round off a single-address bottleneck.
__device__ __forceinline__ uint32_t ld_parent(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

__device__ __forceinline__ void wcc_link(uint32_t *f, uint32_t u, uint32_t v)
{
    uint32_t p1 = ld_parent(f + u), p2 = ld_parent(f + v);
    while (p1 != p2) {
        const uint32_t hi = p1 > p2 ? p1 : p2, lo = p1 > p2 ? p2 : p1;
        uint32_t ph = ld_parent(f + hi);
        if (ph == lo) break;
        if (ph == hi) { // looks like a root: hook it
            ph = atomicCAS(&f[hi], hi, lo);
            if (ph == hi || ph == lo) break;
        }
        p1 = ph; // hi has a parent by now: climb
        p2 = lo;
    }
}

// round r: every vertex is united with the r-th stored entry of its row (the out-adjacency alone
// serves directed graphs: a union is symmetric).  The forest is compressed between the rounds so
// that the chains wcc_link walks stay short.
constexpr uint32_t WCC_SAMPLE_K = 2;
__global__ void __launch_bounds__(256)
k_wcc_sample_link(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t n, uint32_t r,
                  uint32_t *__restrict__ f)
{
    uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; u < n; u += stride) {
        const uint64_t a = rowptr[u], b = rowptr[u + 1];
        if (a + r < b) wcc_link(f, (uint32_t)u, col[a + r]);
    }
}

// every vertex points at the root of its tree (parents only ever decrease towards the root, so
// chasing while others compress is safe)
__global__ void k_wcc_compress(uint32_t *__restrict__ f, uint32_t *__restrict__ gp, uint64_t n)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) {
        uint32_t r = f[v];
        for (uint32_t p = f[r]; p != r; p = f[r]) r = p;
        f[v] = r;
        gp[v] = r;
    }
}

// most frequent root among WCC_NSAMPLE evenly spread vertices (ties: the smaller id); one CTA
constexpr int WCC_NSAMPLE = 1024;
__global__ void __launch_bounds__(WCC_NSAMPLE) k_wcc_mode(const uint32_t *__restrict__ f, uint64_t n, uint32_t *__restrict__ giant)
{
    __shared__ uint32_t s_lab[WCC_NSAMPLE];
    __shared__ unsigned long long s_best;
    const uint64_t v = (uint64_t)threadIdx.x * n / WCC_NSAMPLE; // < n
    const uint32_t mine = f[v];
    s_lab[threadIdx.x] = mine;
    if (threadIdx.x == 0) s_best = 0;
    __syncthreads();
    uint32_t c = 0;
    for (int i = 0; i < WCC_NSAMPLE; i++) c += s_lab[i] == mine ? 1u : 0u;
    atomicMax(&s_best, ((unsigned long long)c << 32) | (uint32_t)~mine);
    __syncthreads();
    if (threadIdx.x == 0) *giant = ~(uint32_t)s_best;
}

// vertices outside the giant tree, in any order; count[0] = all of them, count[1] = the non-empty rows among them
__global__ void __launch_bounds__(256)
k_wcc_collect_rest(const uint32_t *__restrict__ f, const uint32_t *__restrict__ giant, uint64_t n,
                   uint32_t *__restrict__ rest, unsigned long long *__restrict__ count)
{
    const uint32_t L = *giant;
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nround = (n + 31) & ~31ull;
    for (; v < nround; v += stride) {
        const bool out = v < n && f[v] != L;
        const unsigned mask = __ballot_sync(FULL, out);
        if (mask == 0) continue;
        unsigned long long base = 0;
        if (lane_id() == 0) base = atomicAdd(count, (unsigned long long)__popc(mask));
        base = __shfl_sync(FULL, base, 0);
        if (out) rest[base + __popc(mask & ((1u << lane_id()) - 1u))] = (uint32_t)v;
    }
}

// rows of the vertices outside S: warp per row, every entry hooks both ways
__global__ void __launch_bounds__(256)
k_wcc_hook_rest(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const uint32_t *__restrict__ rest,
                uint64_t count, uint64_t v0, uint64_t v1, const uint32_t *__restrict__ gp, uint32_t *__restrict__ f,
                int *__restrict__ changed, unsigned long long *__restrict__ inspected)
{
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    const unsigned lane = lane_id();
    bool ch = false;
    unsigned long long seen = 0;
    for (; wid < count; wid += nw) {
        const uint32_t u = rest[wid];
        if (u < v0 || u >= v1) continue; // several GPUs: the owner of the row hooks it
        const uint64_t a = rowptr[u], b = rowptr[u + 1];
        const uint32_t gu = gp[u];
        uint32_t mn = 0xFFFFFFFFu;
        for (uint64_t e = a + lane; e < b; e += 32) {
            const uint32_t v = ld_stream(col + e);
            const uint32_t gv = gp[v];
            mn = min(mn, gv);
            if (gu < gv) ch |= wcc_apply(f, v, gu);
        }
        mn = warp_min_u32(mn);
        if (lane == 0) { if (b > a) ch |= wcc_apply(f, u, mn); seen += b - a; }
    }
    if (ch) *changed = 1;
    if (seen) atomicAdd(inspected, seen);
}

__global__ void k_widen_u32(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) out[v] = in[v];
}

static void wcc_hook_pass(Adj &a, bool both, const uint32_t *gp, uint32_t *f, int *changed)
{
    const RowPlan &p = a.plan;
    if (both) {
        if (p.n_chunks)
            GX_LAUNCH(k_wcc_hook_chunk<true>, (unsigned)p.n_chunks, 256, 0, a.rowptr.p, a.col.p, p.chunk_row.p, p.chunk_begin.p, gp, f,
                      changed);
        GX_LAUNCH(k_wcc_hook<true>, grid_persistent(8), 256, 0, a.rowptr.p, a.col.p, p.part.lo, p.part.hi, gp, f, changed);
        return;
    }
    if (p.n_chunks)
        GX_LAUNCH(k_wcc_hook_chunk<false>, (unsigned)p.n_chunks, 256, 0, a.rowptr.p, a.col.p, p.chunk_row.p, p.chunk_begin.p, gp, f,
                  changed);
    GX_LAUNCH(k_wcc_hook<false>, grid_persistent(8), 256, 0, a.rowptr.p, a.col.p, p.part.lo, p.part.hi, gp, f, changed);
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
        // A directed graph whose in-edge adjacency is not cached is NOT transposed for this (the reference pays
        // for A v A' inside its window, wcc.cpp:53-55; a transposition costs more than the whole run here):
        // every out-entry then hooks both of its ends, which covers A v A' with one pass over A.
        const bool out_only = g->directed && !g->have_in;
        {
            PhaseTimer tb(&c.timing.build_ms);
            ensure_plan(g->out, n);
            if (g->directed && !out_only) { ensure_in_full(g); ensure_plan(g->in, n); } // the rest lists walk in-rows of any block
        }
        const uint64_t m_sym = g->directed ? 2 * g->m : g->m;
        g->res_u64.alloc(n);
        DevBuf<uint32_t> f(n), gp(n), f_prev(multi() ? n : 0), rest(n), giant(1);
        DevBuf<int> changed(1);
        DevBuf<unsigned long long> counts(2); // [0] vertices outside S, [1] entries read by k_wcc_hook_rest
        uint32_t iters = 0;
        uint64_t inspected = 0;
        const char *se = getenv("GX_WCC_SAMPLE"); // GX_WCC_SAMPLE=0: plain FastSV over all rows
        const bool sample = !(se && se[0] == '0');
        bool all_rows = !sample;
        uint64_t n_rest = 0;
        const Partition &part = g->out.plan.part;
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            GX_LAUNCH(k_wcc_init, grid_persistent(8), 256, 0, f.p, gp.p, n);
            if (sample) {
                // every rank samples all vertices on its replica: the forests may differ in shape, but each
                // tree's root is the smallest id of its sampled component, so the compressed stars agree
                for (uint32_t r = 0; r < WCC_SAMPLE_K; r++) {
                    GX_LAUNCH(k_wcc_sample_link, grid_persistent(8), 256, 0, g->out.rowptr.p, g->out.col.p, n, r, f.p);
                    GX_LAUNCH(k_wcc_compress, grid_persistent(8), 256, 0, f.p, gp.p, n);
                }
                GX_LAUNCH(k_wcc_mode, 1, WCC_NSAMPLE, 0, f.p, n, giant.p);
                counts.zero();
                GX_LAUNCH(k_wcc_collect_rest, grid_persistent(8), 256, 0, f.p, giant.p, n, rest.p, counts.p);
                unsigned long long h = 0;
                read_back(&h, counts.p, sizeof(h));
                n_rest = h;
                inspected += 2 * n;
                all_rows = n_rest > n / 4;
                if (out_only && n_rest) all_rows = true; // the rest lists need the in-edges of the rest vertices
            }
            for (;;) {
                if (!all_rows && n_rest == 0) break; // the sample connected everything
                changed.zero();
                if (multi()) GX_CUDA(cudaMemcpyAsync(f_prev.p, f.p, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
                if (all_rows) {
                    wcc_hook_pass(g->out, out_only, gp.p, f.p, changed.p);
                    if (g->directed && !out_only) wcc_hook_pass(g->in, false, gp.p, f.p, changed.p);
                    inspected += out_only ? g->m : m_sym;
                } else {
                    const unsigned grid = grid_for(n_rest * 32, 256) < grid_persistent(8) ? grid_for(n_rest * 32, 256) : grid_persistent(8);
                    GX_LAUNCH(k_wcc_hook_rest, grid, 256, 0, g->out.rowptr.p, g->out.col.p, rest.p, n_rest, part.lo, part.hi, gp.p, f.p,
                              changed.p, counts.p + 1);
                    if (g->directed)
                        GX_LAUNCH(k_wcc_hook_rest, grid, 256, 0, g->in.rowptr.p, g->in.col.p, rest.p, n_rest, part.lo, part.hi, gp.p,
                                  f.p, changed.p, counts.p + 1);
                }
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
            if (!all_rows && n_rest) {
                unsigned long long h = 0;
                read_back(&h, counts.p + 1, sizeof(h));
                inspected += h;
            }
            GX_LAUNCH(k_widen_u32, grid_persistent(8), 256, 0, f.p, n, g->res_u64.p);
        }
        c.timing.iterations = iters + (sample ? 1 : 0);
        c.timing.edges_inspected = inspected;
        // one full pass over the symmetric adjacency is the compulsory figure; a sampled
        // run reads less than that
        c.timing.algorithmic_bytes = (uint64_t)(all_rows ? iters : 1) * (4 * m_sym + 8 * (n + 1) + 24 * n);
        if (comp_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(comp_host, g->res_u64.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
