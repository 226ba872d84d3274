// algo_cdlp.cu -- community detection by label propagation as a per-row mode
// reduction.  Replaces MY_CDLP_GPU / LA_CDLP_CPU (cdlp.cpp:54-81) with the
// semantics of LAGraph_cdlp (LAGraph_cdlp.c:241-333): L0(v) = v; every
// iteration each stored entry (v,u) contributes label(u) to row v -- for
// directed graphs both the out- and the in-adjacency contribute (:272-283), so
// a reciprocal pair counts twice -- and the smallest most frequent label wins
// (:293-323); rows without entries keep their label; stop at itermax or at a
// fix-point (:328-332).  Unlike the reference's CUDA path (cdlp_kernel.cu) the
// directed case counts in-neighbours too, isolated vertices are defined, and
// the scratch is 8 B per slot only for the few rows that spill out of shared
// memory (the reference allocates 48 B per edge, cdlp_kernel.cu:1169).
//
// Rows are binned by entry count d once per graph:
//   S  d <= 64    8-lane group per row,  128-slot open-addressing table in smem
//   M  d <= 512   warp per row,          1024-slot table in smem
//   L  d  > 512   CHUNK-entry pieces inserted by whole CTAs into a global table
//                 of 2d slots, slot-parallel arg-max, per-row finalize
// The arg-max key is (count << 32) | ~label, so max() picks the highest count
// and, among equals, the smallest label -- bit-exact with the sorted-run scan.
// Algorithmic bytes per iteration: 4 m' + 8(n+1) [x2 directed] + 4n + 4n.
#include <algorithm>
#include <vector>

#include "graph.cuh"

namespace gx {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr uint32_t CDLP_S_MAX = 64, CDLP_M_MAX = 512;
constexpr uint32_t SCAN_CHUNK = 4096; // global table slots per CTA in the arg-max pass

struct CdlpPlan {
    bool built = false;
    Partition part; // row blocks balanced by entries (out + in)
    uint64_t nS = 0, nM = 0, nL = 0, n_ins = 0, n_scan = 0, slots = 0;
    DevBuf<uint32_t> listS, listM, listL;
    DevBuf<uint64_t> tab_off;      // nL + 1: first slot of each L row's table
    DevBuf<uint32_t> ins_row;      // insert chunks: index into listL
    DevBuf<uint8_t> ins_side;      // 0 = out adjacency, 1 = in adjacency
    DevBuf<uint64_t> ins_begin;    // first entry of the chunk
    DevBuf<uint32_t> scan_row;     // scan chunks: index into listL
    DevBuf<uint64_t> scan_begin;   // first slot of the chunk
    DevBuf<uint32_t> gkeys, gcnt;  // global tables
    DevBuf<unsigned long long> best; // nL arg-max accumulators
};

__device__ __forceinline__ uint32_t hash32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

__global__ void k_cdlp_bin(const uint64_t *__restrict__ rp0, const uint64_t *__restrict__ rp1, uint64_t v0, uint64_t v1,
                           uint32_t *__restrict__ listS, uint32_t *__restrict__ listM, uint32_t *__restrict__ listL,
                           unsigned long long *__restrict__ counts, int write)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < v1; v += stride) {
        uint64_t d = rp0[v + 1] - rp0[v];
        if (rp1) d += rp1[v + 1] - rp1[v];
        if (d == 0) continue;
        int b = d <= CDLP_S_MAX ? 0 : d <= CDLP_M_MAX ? 1 : 2;
        unsigned long long pos = atomicAdd(&counts[b], 1ull);
        if (write) (b == 0 ? listS : b == 1 ? listM : listL)[pos] = (uint32_t)v;
    }
}

__global__ void k_cdlp_init(uint32_t *__restrict__ a, uint32_t *__restrict__ b, uint64_t n)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) { a[v] = (uint32_t)v; b[v] = (uint32_t)v; }
}

// G lanes per row, T-slot table per group in shared memory
template <int G, int T>
__global__ void __launch_bounds__(256)
k_cdlp_rows(const uint32_t *__restrict__ list, uint64_t count, const uint64_t *__restrict__ rp0,
            const uint32_t *__restrict__ col0, const uint64_t *__restrict__ rp1, const uint32_t *__restrict__ col1,
            const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt, int *__restrict__ changed)
{
    constexpr int GROUPS = 256 / G;
    extern __shared__ uint32_t s_tab[];
    uint32_t *s_key = s_tab;
    uint32_t *s_cnt = s_tab + GROUPS * T;
    for (int i = threadIdx.x; i < GROUPS * T; i += 256) { s_key[i] = EMPTY; s_cnt[i] = 0; }
    __syncthreads();
    const unsigned sub = threadIdx.x & (G - 1);
    uint32_t *key = s_key + (threadIdx.x / G) * T;
    uint32_t *cnt = s_cnt + (threadIdx.x / G) * T;
    uint64_t gi = ((uint64_t)blockIdx.x * 256 + threadIdx.x) / G;
    const uint64_t ngrp = ((uint64_t)gridDim.x * 256) / G;
    const uint64_t trips = (count + ngrp - 1) / ngrp;
    bool ch = false;
    for (uint64_t t = 0; t < trips; t++, gi += ngrp) {
        const bool live = gi < count;
        uint32_t v = 0;
        uint64_t a0 = 0, d0 = 0, a1 = 0, d = 0;
        if (live) {
            v = list[gi];
            a0 = rp0[v]; d0 = rp0[v + 1] - a0;
            d = d0;
            if (rp1) { a1 = rp1[v]; d += rp1[v + 1] - a1; }
        }
        // table size: smallest power of two >= 2d (>= 2), at most T
        uint32_t teff = 2;
        while (teff < 2 * d && teff < (uint32_t)T) teff <<= 1;
        const uint32_t mask = teff - 1;
        for (uint64_t k = sub; k < d; k += G) {
            const uint32_t u = k < d0 ? ld_stream(col0 + a0 + k) : ld_stream(col1 + a1 + (k - d0));
            const uint32_t lab = cur[u];
            uint32_t s = hash32(lab) & mask;
            for (;;) {
                const uint32_t old = atomicCAS(&key[s], EMPTY, lab);
                if (old == EMPTY || old == lab) { atomicAdd(&cnt[s], 1u); break; }
                s = (s + 1) & mask;
            }
        }
        __syncwarp();
        unsigned long long best = 0;
        if (live) {
            for (uint32_t s = sub; s < teff; s += G) {
                const uint32_t c = cnt[s];
                if (c) {
                    const unsigned long long kk = ((unsigned long long)c << 32) | (uint32_t)~key[s];
                    best = kk > best ? kk : best;
                    key[s] = EMPTY;
                    cnt[s] = 0;
                }
            }
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const unsigned long long x = __shfl_xor_sync(FULL, best, o);
            best = x > best ? x : best;
        }
        if (live && sub == 0) {
            const uint32_t nl = ~(uint32_t)best;
            nxt[v] = nl;
            if (nl != cur[v]) ch = true;
        }
        __syncwarp();
    }
    if (ch) *changed = 1;
}

// L rows, pass 1: one CTA per CHUNK entries, insert into the row's global table
__global__ void __launch_bounds__(256)
k_cdlp_big_insert(const uint32_t *__restrict__ listL, const uint32_t *__restrict__ ins_row,
                  const uint8_t *__restrict__ ins_side, const uint64_t *__restrict__ ins_begin,
                  const uint64_t *__restrict__ tab_off, const uint64_t *__restrict__ rp0, const uint32_t *__restrict__ col0,
                  const uint64_t *__restrict__ rp1, const uint32_t *__restrict__ col1, const uint32_t *__restrict__ cur,
                  uint32_t *__restrict__ gkeys, uint32_t *__restrict__ gcnt)
{
    const uint32_t c = blockIdx.x;
    const uint32_t li = ins_row[c];
    const uint32_t v = listL[li];
    const bool side = ins_side[c] != 0;
    const uint64_t *rp = side ? rp1 : rp0;
    const uint32_t *col = side ? col1 : col0;
    const uint64_t b0 = ins_begin[c];
    const uint64_t row_end = rp[v + 1];
    const uint64_t e_end = (b0 + CHUNK < row_end) ? b0 + CHUNK : row_end;
    const uint64_t t0 = tab_off[li];
    const uint64_t tsize = tab_off[li + 1] - t0;
#pragma unroll 4
    for (uint64_t e = b0 + threadIdx.x; e < e_end; e += 256) {
        const uint32_t lab = cur[ld_stream(col + e)];
        uint64_t s = ((uint64_t)hash32(lab) * tsize) >> 32;
        for (;;) {
            const uint32_t old = atomicCAS(&gkeys[t0 + s], EMPTY, lab);
            if (old == EMPTY || old == lab) { atomicAdd(&gcnt[t0 + s], 1u); break; }
            s = (s + 1 == tsize) ? 0 : s + 1;
        }
    }
}

// L rows, pass 2: slot-parallel arg-max, table reset
__global__ void __launch_bounds__(256)
k_cdlp_big_scan(const uint32_t *__restrict__ scan_row, const uint64_t *__restrict__ scan_begin,
                const uint64_t *__restrict__ tab_off, uint32_t *__restrict__ gkeys, uint32_t *__restrict__ gcnt,
                unsigned long long *__restrict__ best_out)
{
    const uint32_t c = blockIdx.x;
    const uint32_t li = scan_row[c];
    const uint64_t s0 = scan_begin[c];
    const uint64_t t_end = tab_off[li + 1];
    const uint64_t s_end = (s0 + SCAN_CHUNK < t_end) ? s0 + SCAN_CHUNK : t_end;
    unsigned long long best = 0;
    for (uint64_t s = s0 + threadIdx.x; s < s_end; s += 256) {
        const uint32_t cc = __ldcg(gcnt + s);
        if (cc) {
            const unsigned long long kk = ((unsigned long long)cc << 32) | (uint32_t)~__ldcg(gkeys + s);
            best = kk > best ? kk : best;
            gkeys[s] = EMPTY;
            gcnt[s] = 0;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long x = __shfl_xor_sync(FULL, best, o);
        best = x > best ? x : best;
    }
    __shared__ unsigned long long red[8];
    if (lane_id() == 0) red[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int i = 1; i < 8; i++) best = red[i] > best ? red[i] : best;
        if (best) atomicMax(&best_out[li], best);
    }
}

__global__ void k_cdlp_big_final(const uint32_t *__restrict__ listL, uint64_t nL, unsigned long long *__restrict__ best,
                                 const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt, int *__restrict__ changed)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nL) return;
    const uint32_t v = listL[i];
    const uint32_t nl = ~(uint32_t)best[i];
    best[i] = 0;
    nxt[v] = nl;
    if (nl != cur[v]) *changed = 1;
}

__global__ void k_widen_u32_cdlp(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) out[v] = in[v];
}

static CdlpPlan *build_cdlp_plan(gx_graph *g)
{
    CdlpPlan *p = new CdlpPlan();
    const uint64_t n = g->n;
    const uint64_t *rp0 = g->out.rowptr.p;
    const uint64_t *rp1 = g->directed ? g->in.rowptr.p : nullptr;
    p->part = make_partition(rp0, rp1, n);
    const uint64_t v0 = p->part.lo, v1 = p->part.hi;
    DevBuf<unsigned long long> counts(3);
    counts.zero();
    DevBuf<uint32_t> dummy(1);
    GX_LAUNCH(k_cdlp_bin, grid_persistent(8), 256, 0, rp0, rp1, v0, v1, dummy.p, dummy.p, dummy.p, counts.p, 0);
    unsigned long long h[3];
    read_back(h, counts.p, sizeof(h));
    p->nS = h[0]; p->nM = h[1]; p->nL = h[2];
    p->listS.alloc(p->nS ? p->nS : 1);
    p->listM.alloc(p->nM ? p->nM : 1);
    p->listL.alloc(p->nL ? p->nL : 1);
    counts.zero();
    GX_LAUNCH(k_cdlp_bin, grid_persistent(8), 256, 0, rp0, rp1, v0, v1, p->listS.p, p->listM.p, p->listL.p, counts.p, 1);
    if (p->nL) {
        std::vector<uint32_t> L(p->nL);
        std::vector<uint64_t> h0(n + 1), h1;
        cudaStream_t s = ctx().stream;
        GX_CUDA(cudaMemcpyAsync(L.data(), p->listL.p, p->nL * sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        GX_CUDA(cudaMemcpyAsync(h0.data(), rp0, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        if (rp1) { h1.resize(n + 1); GX_CUDA(cudaMemcpyAsync(h1.data(), rp1, (n + 1) * sizeof(uint64_t), cudaMemcpyDeviceToHost, s)); }
        GX_CUDA(cudaStreamSynchronize(s));
        std::sort(L.begin(), L.end());
        std::vector<uint64_t> tab_off(p->nL + 1), ins_begin, scan_begin;
        std::vector<uint32_t> ins_row, scan_row;
        std::vector<uint8_t> ins_side;
        uint64_t off = 0;
        for (uint64_t i = 0; i < p->nL; i++) {
            const uint32_t v = L[i];
            uint64_t d = h0[v + 1] - h0[v];
            for (uint64_t b = h0[v]; b < h0[v + 1]; b += CHUNK) { ins_row.push_back((uint32_t)i); ins_side.push_back(0); ins_begin.push_back(b); }
            if (rp1) {
                d += h1[v + 1] - h1[v];
                for (uint64_t b = h1[v]; b < h1[v + 1]; b += CHUNK) { ins_row.push_back((uint32_t)i); ins_side.push_back(1); ins_begin.push_back(b); }
            }
            tab_off[i] = off;
            for (uint64_t sb = off; sb < off + 2 * d; sb += SCAN_CHUNK) { scan_row.push_back((uint32_t)i); scan_begin.push_back(sb); }
            off += 2 * d;
        }
        tab_off[p->nL] = off;
        p->slots = off;
        p->n_ins = ins_row.size();
        p->n_scan = scan_row.size();
        p->tab_off.alloc(p->nL + 1);
        p->ins_row.alloc(p->n_ins); p->ins_side.alloc(p->n_ins); p->ins_begin.alloc(p->n_ins);
        p->scan_row.alloc(p->n_scan); p->scan_begin.alloc(p->n_scan);
        p->gkeys.alloc(off); p->gcnt.alloc(off); p->best.alloc(p->nL);
        GX_CUDA(cudaMemcpyAsync(p->listL.p, L.data(), p->nL * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->tab_off.p, tab_off.data(), (p->nL + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->ins_row.p, ins_row.data(), p->n_ins * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->ins_side.p, ins_side.data(), p->n_ins * sizeof(uint8_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->ins_begin.p, ins_begin.data(), p->n_ins * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->scan_row.p, scan_row.data(), p->n_scan * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->scan_begin.p, scan_begin.data(), p->n_scan * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        p->gkeys.fill_byte(0xFF);
        p->gcnt.zero();
        p->best.zero();
        GX_CUDA(cudaStreamSynchronize(s));
    }
    p->built = true;
    return p;
}

} // namespace gx

using namespace gx;

void gx_cdlp_plan_free(void *p) { delete (CdlpPlan *)p; }

extern "C" int gx_cdlp(gx_graph *g, int itermax, uint64_t *label_host)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        GX_REQUIRE(itermax >= 0, "negative iteration count");
        Context &c = ctx();
        c.timing = gx_timing{};
        const uint64_t n = g->n;
        if (n == 0) return;
        ensure_in_adj(g);
        if (!g->cdlp_plan) {
            PhaseTimer tb(&c.timing.build_ms);
            g->cdlp_plan = build_cdlp_plan(g);
        }
        CdlpPlan &p = *(CdlpPlan *)g->cdlp_plan;
        const uint64_t *rp0 = g->out.rowptr.p, *rp1 = g->directed ? g->in.rowptr.p : nullptr;
        const uint32_t *col0 = g->out.col.p, *col1 = g->directed ? g->in.col.p : nullptr;
        constexpr size_t SMEM_M = (256 / 32) * 1024 * 8, SMEM_S = (256 / 8) * 128 * 8;
        GX_CUDA(cudaFuncSetAttribute(k_cdlp_rows<32, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_M));
        GX_CUDA(cudaFuncSetAttribute(k_cdlp_rows<8, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_S));
        g->res_u64.alloc(n);
        DevBuf<uint32_t> la(n), lb(n);
        DevBuf<int> changed(1);
        uint32_t iters = 0;
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            GX_LAUNCH(k_cdlp_init, grid_persistent(8), 256, 0, la.p, lb.p, n);
            uint32_t *cur = la.p, *nxt = lb.p;
            for (int it = 0; it < itermax; it++) {
                changed.zero();
                if (p.nL) {
                    GX_LAUNCH(k_cdlp_big_insert, (unsigned)p.n_ins, 256, 0, p.listL.p, p.ins_row.p, p.ins_side.p, p.ins_begin.p,
                              p.tab_off.p, rp0, col0, rp1, col1, cur, p.gkeys.p, p.gcnt.p);
                    GX_LAUNCH(k_cdlp_big_scan, (unsigned)p.n_scan, 256, 0, p.scan_row.p, p.scan_begin.p, p.tab_off.p, p.gkeys.p,
                              p.gcnt.p, p.best.p);
                    GX_LAUNCH(k_cdlp_big_final, grid_for(p.nL, 256), 256, 0, p.listL.p, p.nL, p.best.p, cur, nxt, changed.p);
                }
                if (p.nM)
                    GX_LAUNCH((k_cdlp_rows<32, 1024>), grid_persistent(3), 256, SMEM_M, p.listM.p, p.nM, rp0, col0, rp1, col1, cur, nxt, changed.p);
                if (p.nS)
                    GX_LAUNCH((k_cdlp_rows<8, 128>), grid_persistent(6), 256, SMEM_S, p.listS.p, p.nS, rp0, col0, rp1, col1, cur, nxt, changed.p);
                if (multi()) {
                    allgatherv(nxt, Dt::U32, p.part);       // owners publish their new labels
                    allreduce(changed.p, 1, Dt::I32, Red::Max);
                }
                uint32_t *t = cur; cur = nxt; nxt = t;
                iters++;
                int h = 0;
                read_back(&h, changed.p, sizeof(h));
                if (!h) break;
            }
            GX_LAUNCH(k_widen_u32_cdlp, grid_persistent(8), 256, 0, cur, n, g->res_u64.p);
        }
        const uint64_t m_eff = g->directed ? 2 * g->m : g->m;
        c.timing.iterations = iters;
        c.timing.edges_inspected = m_eff * iters;
        c.timing.algorithmic_bytes = (uint64_t)iters * (4 * m_eff + (g->directed ? 2 : 1) * 8 * (n + 1) + 8 * n);
        if (label_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(label_host, g->res_u64.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
