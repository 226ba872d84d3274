// algo_lcc.cu -- local clustering coefficient as masked plus.pair SpGEMM
// triangle counting.  Replaces LA_LCC (lcc.cpp:61-71) -> LAGraph_lcc(&d, A,
// symmetric = !directed, sanitize = false):
//   N(v) = (in U out neighbours) \ {v}, d = |N(v)|,
//   lcc(v) = #{(a,b) in E : a,b in N(v)} / (d (d-1)),   0.0 when d < 2.
// LAGraph evaluates C<C> = (A or A+A') * U' over the union structure C = A v A';
// here U = A v A' is oriented from lower to higher (degree, id) (graph.cu), every
// triangle {u,v,w} is found exactly once as the intersection N+(u) ^ N+(v) of an
// oriented entry u->v, and each corner receives the multiplicity (1 or 2
// directed entries) of the opposite side, carried in bit 31 of the column id.
// For a symmetric store every multiplicity is 2, which yields Graphalytics'
// 2*tri/(d(d-1)).  Counting is integer; one FP64 divide per vertex.
//
//   k_lcc_count   8-lane group per oriented entry: lanes stride the shorter list and look each
//                 element up in the longer one -- through the row's membership table (open
//                 addressing, 4 slots per entry, built once per graph) or, for rows under 16 entries,
//                 by binary search.  The entries are ordered by the owner of the longer list; a CTA
//                 copies an owner's table (<= 32 KB: every table on RMAT-22, where the largest
//                 oriented row has 1048 entries) into shared memory and serves the whole segment from it
//   k_lcc_final   lcc = num / (d (d-1))
// Algorithmic bytes: 4m + 8(n+1) + 4*sum_{u->v}(d+(u) + d+(v)) + 8n.
#include <cstdlib>

#include "graph.cuh"

namespace gx {

constexpr uint32_t LCC_RUN = 1024; // entries a CTA draws at a time (32 trips of its 32 groups; 64 / 256 / 1024 / 4096: 85 / 77 / 74.5 / 76 ms)
constexpr uint32_t IDMASK = ~LCC_MULT_BIT;

constexpr uint32_t LCC_SMEM_SLOTS = 8192; // membership tables up to this many slots (32 KB) are staged in shared memory

// One intersection: the lanes of an 8-lane group stride the shorter list [sa, sb) and look every element up in the
// longer row's membership table -- TAB_SMEM: the CTA's shared-memory copy of it (0.27 LSU wavefronts per probe instead
// of one: the run's probes are ~200x the table's size), else in global memory; rows under LCC_TAB_MIN entries have no
// table and are searched.  Adds the corner counts of the common neighbours, returns this lane's share for u and v.
template <bool TAB_SMEM, int UNROLL, int LCC_G>
__device__ __forceinline__ void lcc_intersect(const uint32_t *__restrict__ ocol, const uint32_t *__restrict__ tab, const uint32_t *s_tab,
                                              uint64_t sa, uint64_t sb, uint64_t la, uint64_t lb, uint64_t t0, uint64_t tmask64,
                                              unsigned sub, bool u_short, unsigned long long m_uv, unsigned long long *__restrict__ num,
                                              unsigned long long &su, unsigned long long &sv, int diag)
{
    // Index arithmetic in 32 bits relative to the lists' first entries (an oriented row has fewer than 2^31 entries,
    // a table fewer than 2^32 slots), counts in 32 bits (at most two per walked element).  A fifth fewer integer
    // instructions than the 64-bit form; the run time did not move (44.3 vs 44.5 ms at RMAT-22): what binds the walk
    // is the latency of the walked lists' elements, not the issue rate.
    const uint32_t *__restrict__ sl_p = ocol + sa; // the walked (shorter) list
    const uint32_t slen = (uint32_t)(sb - sa);
    const uint32_t tmask = (uint32_t)tmask64;       // table size (power of two) or 0
    const uint32_t *__restrict__ tb = tab + t0;
    uint32_t c_long = 0, c_short = 0; // multiplicities met on the side of the long / the short list's owner
    // UNROLL elements of the shorter list are requested before the first one is looked up: the walk is a chain of
    // dependent L2 / DRAM latencies otherwise (one load in flight per lane)
    for (uint32_t i = sub; i < slen; i += LCC_G * UNROLL) {
        uint32_t cs_[UNROLL];
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const uint32_t k = i + (uint32_t)j * LCC_G;
            cs_[j] = k < slen ? sl_p[k] : 0xFFFFFFFFu; // (no entry has this value: ids are < 2^31 - 1)
        }
#pragma unroll
        for (int j = 0; j < UNROLL; j++) {
            const uint32_t cs = cs_[j];
            if (cs == 0xFFFFFFFFu) continue;
            const uint32_t w = cs & IDMASK;
            uint32_t cl = 0xFFFFFFFFu; // the longer list's entry for w, if any
            if (tmask) {
                // open addressing at load 1/4: a miss ends after 1.4 probes on average, all in one line
                uint32_t sl = (uint32_t)lcc_tab_hash(w, tmask64);
                for (;;) {
                    const uint32_t c = TAB_SMEM ? s_tab[sl] : tb[sl];
                    if (c == 0xFFFFFFFFu || (c & IDMASK) == w) { cl = c; break; }
                    sl = (sl + 1) & (tmask - 1);
                }
            } else {
                uint64_t lo = la, hi = lb;
                while (lo < hi) {
                    const uint64_t mid = (lo + hi) >> 1;
                    if ((ocol[mid] & IDMASK) < w) lo = mid + 1; else hi = mid;
                }
                if (lo < lb) cl = ocol[lo];
            }
            if (cl != 0xFFFFFFFFu && (cl & IDMASK) == w) {
                // corner u gets mult(v,w), corner v gets mult(u,w), corner w gets mult(u,v); a multiplicity is 1 or 2
                c_short += 1u + (cs >> 31); // side (short owner, w)
                c_long += 1u + (cl >> 31);  // side (long owner, w)
                if (!diag) atomicAdd(&num[w], m_uv); // diag (GX_LCC_VAR=1): timing diagnostic, results are wrong
            }
        }
    }
    su += u_short ? c_long : c_short;
    sv += u_short ? c_short : c_long;
}

template <int UNROLL, int LCC_G>
__global__ void __launch_bounds__(256, 6) // 6 CTAs of 32 KB + 256 threads per SM: 48 warps, <= 40 registers
k_lcc_count(const uint64_t *__restrict__ orp, const uint32_t *__restrict__ ocol, const uint32_t *__restrict__ eu,
            const uint32_t *__restrict__ ev, const uint32_t *__restrict__ eowner, const uint64_t *__restrict__ tab_off,
            const uint32_t *__restrict__ tab, uint64_t e0, uint64_t om, uint32_t run, unsigned long long *__restrict__ next_run,
            unsigned long long *__restrict__ num, int diag)
{
    // oriented entries [e0, om) are this rank's share, ordered by the owner of the longer list of their intersection.
    // A CTA draws runs of `run` consecutive entries from a counter (the draw balances the load whatever a run costs) and
    // walks a run segment by segment -- a segment = the consecutive entries of one owner: the owner's membership table
    // is copied into shared memory once and every probe of the segment hits that copy.
    extern __shared__ uint32_t s_tab[];
    const unsigned sub = threadIdx.x & (LCC_G - 1), grp = threadIdx.x / LCC_G;
    __shared__ unsigned long long s_base;
    __shared__ unsigned long long s_seg_end;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(next_run, (unsigned long long)run);
        __syncthreads();
        const uint64_t base = e0 + s_base;
        if (base >= om) break;
        const uint64_t run_end = base + run < om ? base + run : om;
        uint64_t pos = base;
        while (pos < run_end) {
            // ---- the segment [pos, seg_end) of the owner of entry `pos` (owners ascend along the list)
            const uint32_t owner = eowner[pos];
            if (threadIdx.x == 0) s_seg_end = run_end;
            __syncthreads();
            // owners ascend along the list, so the segment's end is the one position whose owner differs while its
            // predecessor's does not: a single writer (every thread beyond the end used to queue an atomicMin on the
            // same shared word, ~250 serialised atomics per segment of a few dozen entries)
            for (uint64_t i = pos + 1 + threadIdx.x; i < run_end; i += 256)
                if (eowner[i] != owner) {
                    if (eowner[i - 1] == owner) s_seg_end = i;
                    break;
                }
            const uint64_t t0 = tab_off[owner], tmask = tab_off[owner + 1] - t0; // table size (power of two) or 0
            const bool staged = tmask != 0 && tmask <= LCC_SMEM_SLOTS;
            if (staged)
                for (uint32_t i = threadIdx.x; i < (uint32_t)tmask; i += 256) s_tab[i] = tab[t0 + i];
            __syncthreads();
            const uint64_t seg_end = s_seg_end;
            for (uint64_t r0 = pos; r0 < seg_end; r0 += 256 / LCC_G) { // same trip count for all groups of a warp
                const uint64_t gi = r0 + grp;
                unsigned long long su = 0, sv = 0;
                uint32_t u = 0, v = 0;
                const bool live = gi < seg_end;
                if (live) {
                    u = eu[gi];
                    const uint32_t cv = ev[gi];
                    v = cv & IDMASK;
                    const unsigned long long m_uv = (cv & LCC_MULT_BIT) ? 2ull : 1ull;
                    const uint64_t ua = orp[u], ub = orp[u + 1], va = orp[v], vb = orp[v + 1];
                    // lanes walk the shorter list (sa..sb), look up in the longer one (la..lb) = the owner's
                    const bool u_short = (ub - ua) <= (vb - va);
                    const uint64_t sa = u_short ? ua : va, sb = u_short ? ub : vb;
                    const uint64_t la = u_short ? va : ua, lb = u_short ? vb : ub;
                    if (staged) lcc_intersect<true, UNROLL, LCC_G>(ocol, tab, s_tab, sa, sb, la, lb, t0, tmask, sub, u_short, m_uv, num, su, sv, diag);
                    else lcc_intersect<false, UNROLL, LCC_G>(ocol, tab, s_tab, sa, sb, la, lb, t0, tmask, sub, u_short, m_uv, num, su, sv, diag);
                }
#pragma unroll
                for (int o = LCC_G / 2; o > 0; o >>= 1) {
                    su += __shfl_xor_sync(FULL, su, o);
                    sv += __shfl_xor_sync(FULL, sv, o);
                }
                if (live && sub == 0) {
                    if (su) atomicAdd(&num[u], su);
                    if (sv) atomicAdd(&num[v], sv);
                }
            }
            __syncthreads(); // the table copy is overwritten by the next segment
            pos = seg_end;
        }
    }
}

__global__ void k_lcc_final(const unsigned long long *__restrict__ num, const uint32_t *__restrict__ udeg, uint64_t n,
                            double *__restrict__ lcc)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) {
        const double d = (double)udeg[v];
        lcc[v] = udeg[v] < 2 ? 0.0 : (double)num[v] / (d * (d - 1.0));
    }
}

} // namespace gx

using namespace gx;

extern "C" int gx_lcc(gx_graph *g, double *lcc_host)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        Context &c = ctx();
        c.timing = gx_timing{};
        const uint64_t n = g->n;
        if (n == 0) return;
        ensure_lcc_cache(g);
        g->res_f64.alloc(n);
        DevBuf<unsigned long long> num(n), next_run(1);
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            num.zero();
            next_run.zero();
            uint32_t run = LCC_RUN;
            int unroll = 4; // elements of the shorter list in flight per lane (tuning knob: 1 / 2 / 4 / 8)
            if (const char *e = getenv("GX_LCC_UNROLL")) unroll = atoi(e);
            int lanes = 8; // lanes per oriented entry (tuning knob: 4 / 8 / 16)
            if (const char *e = getenv("GX_LCC_G")) lanes = atoi(e);
            auto kern = lanes == 4 ? (unroll <= 2 ? k_lcc_count<2, 4> : k_lcc_count<4, 4>)
                      : lanes == 16 ? (unroll <= 2 ? k_lcc_count<2, 16> : k_lcc_count<4, 16>)
                      : (unroll <= 1 ? k_lcc_count<1, 8> : unroll == 2 ? k_lcc_count<2, 8> : unroll <= 4 ? k_lcc_count<4, 8> : k_lcc_count<8, 8>);
            GX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(LCC_SMEM_SLOTS * sizeof(uint32_t))));
            // tuning knob; a multiple of 32 so that the four 8-lane groups of a warp make the same number of trips
            // (the shuffles inside the trip are warp-wide)
            if (const char *e = getenv("GX_LCC_RUN")) run = (uint32_t)atoi(e) >= 32 ? ((uint32_t)atoi(e) + 31u) & ~31u : 32;
            // the oriented entry list is split evenly over the ranks; corner counts are summed
            const Partition part = make_even_partition(g->om);
            if (part.hi > part.lo)
                GX_LAUNCH(kern, grid_persistent(6), 256, LCC_SMEM_SLOTS * sizeof(uint32_t), g->orowptr.p, g->ocol.p, g->lcc_eu.p,
                          g->lcc_ev.p, g->lcc_owner.p, g->ltab_off.p, g->ltab.p, part.lo, part.hi, run, next_run.p, num.p, getenv("GX_LCC_VAR") ? atoi(getenv("GX_LCC_VAR")) : 0);
            allreduce(num.p, n, Dt::U64, Red::Sum);
            GX_LAUNCH(k_lcc_final, grid_persistent(8), 256, 0, num.p, g->udeg.p, n, g->res_f64.p);
        }
        c.timing.iterations = 1;
        c.timing.edges_inspected = g->lcc_list_bytes / 4;
        c.timing.algorithmic_bytes = 4 * g->m + 8 * (n + 1) + g->lcc_list_bytes + 8 * n;
        if (lcc_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(lcc_host, g->res_f64.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
