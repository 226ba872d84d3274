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
//                 addressing, 4 slots per entry, built once per graph: 1-2 probes in one cache line)
//                 or, for rows under 16 entries, by binary search
//   k_lcc_final   lcc = num / (d (d-1))
// Algorithmic bytes: 4m + 8(n+1) + 4*sum_{u->v}(d+(u) + d+(v)) + 8n.
#include <cstdlib>

#include "graph.cuh"

namespace gx {

constexpr int LCC_G = 8;
constexpr uint32_t LCC_RUN = 1024; // entries a CTA draws at a time (32 trips of its 32 groups; 64 / 256 / 1024 / 4096: 85 / 77 / 74.5 / 76 ms)
constexpr uint32_t IDMASK = ~LCC_MULT_BIT;

__global__ void __launch_bounds__(256)
k_lcc_count(const uint64_t *__restrict__ orp, const uint32_t *__restrict__ ocol, const uint32_t *__restrict__ eu,
            const uint32_t *__restrict__ ev, const uint64_t *__restrict__ tab_off, const uint32_t *__restrict__ tab, uint64_t e0, uint64_t om,
            uint32_t run, unsigned long long *__restrict__ next_run, unsigned long long *__restrict__ num)
{
    // oriented entries [e0, om) are this rank's share.  A CTA draws runs of LCC_RUN consecutive entries
    // from a counter: consecutive entries share the owner of the longer list, so the CTA probes one
    // membership table for a while (L1 hits) and the draw balances the load whatever a run costs.
    const unsigned sub = threadIdx.x & (LCC_G - 1), grp = threadIdx.x / LCC_G;
    __shared__ unsigned long long s_base;
    for (;;) {
        __syncthreads();
        if (threadIdx.x == 0) s_base = atomicAdd(next_run, (unsigned long long)run);
        __syncthreads();
        const uint64_t base = e0 + s_base;
        if (base >= om) break;
      for (uint32_t r = grp; r < run; r += 256 / LCC_G) {
        const uint64_t gi = base + r;
        unsigned long long su = 0, sv = 0;
        uint32_t u = 0, v = 0;
        const bool live = gi < om;
        if (live) {
            u = eu[gi]; // entries in the order of the longer list's owner (graph.cu)
            const uint32_t cv = ev[gi];
            v = cv & IDMASK;
            const unsigned long long m_uv = (cv & LCC_MULT_BIT) ? 2ull : 1ull;
            uint64_t ua = orp[u], ub = orp[u + 1], va = orp[v], vb = orp[v + 1];
            // lanes walk the shorter list (sa..sb), search the longer one (la..lb)
            const bool u_short = (ub - ua) <= (vb - va);
            const uint64_t sa = u_short ? ua : va, sb = u_short ? ub : vb;
            const uint64_t la = u_short ? va : ua, lb = u_short ? vb : ub;
            // the longer list's membership table, if it has one (rows from LCC_TAB_MIN entries on)
            const uint32_t owner = u_short ? v : u;
            const uint64_t t0 = tab_off[owner], tmask = tab_off[owner + 1] - t0; // size (power of two) or 0
            for (uint64_t i = sa + sub; i < sb; i += LCC_G) {
                const uint32_t cs = ocol[i];
                const uint32_t w = cs & IDMASK;
                uint32_t cl = 0xFFFFFFFFu; // the longer list's entry for w, if any
                if (tmask) {
                    // open addressing at load 1/4: a miss ends after 1.4 probes on average, all in one line
                    uint64_t sl = lcc_tab_hash(w, tmask);
                    for (;;) {
                        const uint32_t c = tab[t0 + sl];
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
                {
                    if (cl != 0xFFFFFFFFu && (cl & IDMASK) == w) {
                        const unsigned long long m_s = (cs & LCC_MULT_BIT) ? 2ull : 1ull; // side (short owner, w)
                        const unsigned long long m_l = (cl & LCC_MULT_BIT) ? 2ull : 1ull; // side (long owner, w)
                        // corner u gets mult(v,w), corner v gets mult(u,w), corner w gets mult(u,v)
                        su += u_short ? m_l : m_s;
                        sv += u_short ? m_s : m_l;
                        atomicAdd(&num[w], m_uv);
                    }
                }
            }
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
            // tuning knob; a multiple of 32 so that the four 8-lane groups of a warp make the same number of trips
            // (the shuffles inside the trip are warp-wide)
            if (const char *e = getenv("GX_LCC_RUN")) run = (uint32_t)atoi(e) >= 32 ? ((uint32_t)atoi(e) + 31u) & ~31u : 32;
            // the oriented entry list is split evenly over the ranks; corner counts are summed
            const Partition part = make_even_partition(g->om);
            if (part.hi > part.lo)
                GX_LAUNCH(k_lcc_count, grid_persistent(8), 256, 0, g->orowptr.p, g->ocol.p, g->lcc_eu.p, g->lcc_ev.p, g->ltab_off.p, g->ltab.p, part.lo, part.hi,
                          run, next_run.p, num.p);
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
