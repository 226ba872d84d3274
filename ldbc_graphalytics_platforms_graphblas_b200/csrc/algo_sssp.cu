// algo_sssp.cu -- single-source shortest paths as min.plus relaxation over
// FP64 weights.  Replaces LA_SSSP (sssp.cpp:53-81): zero diagonal +
// LAGr_SingleSourceShortestPath(&d, G, src, delta = 2.5).  The bucket width
// only schedules work; the result is the fix-point d(v) = min_u fl(d(u)+w(u,v)),
// d(src) = 0, which is unique because rounded addition is monotone -- so the
// distances are bit-identical to LAGraph's (and to Dijkstra's).  The zero
// diagonal the wrapper inserts (sssp.cpp:60-62) never changes a minimum and is
// not materialised.  Unreached vertices stay +inf (printed `infinity`).
//
// Non-negative doubles order like their bit patterns, so relaxations are
// atomicMin on the uint64 image of the distance.  With weights in (0,1] and
// delta = 2.5 nearly every vertex falls in LAGraph's first bucket, i.e. the
// reference also runs frontier sweeps of min.plus until nothing changes.
//   k_sssp_relax      warp per frontier vertex (push over out-edges + weights);
//                     hubs are re-queued as CHUNK-entry pieces for whole CTAs;
//                     an improved vertex enters the next frontier once (flag)
//   k_sssp_relax_big  CTA per piece
// Work efficiency: plain frontier sweeps relax every edge ~4x on RMAT (a vertex is expanded again
// each time its distance drops).  On one GPU the frontier is therefore bucketed like
// delta-stepping's (LAGraph's own algorithm): vertices below the current threshold T are expanded
// now ("near"), the others only get their state marked ("far") and are collected by one pass
// over the state array when the near queue runs dry and T advances by delta.  delta only
// schedules work -- the fix-point, hence every bit of the result, is the same.
// Algorithmic bytes (one-pass bound): 12m + 8(n+1) + 16n.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "graph.cuh"

namespace gx {

constexpr uint32_t SSSP_BIG = 4096;
constexpr unsigned long long INF_BITS = 0x7FF0000000000000ull;

struct SsspCounters { unsigned long long next_count, big_count, relaxed, far_count, far_min; };

__global__ void k_sssp_init(unsigned long long *__restrict__ dist, uint64_t n, uint32_t src, uint32_t *__restrict__ queue,
                            uint32_t *__restrict__ inq)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) { dist[v] = (v == src) ? 0ull : INF_BITS; inq[v] = 0; }
    if (blockIdx.x == 0 && threadIdx.x == 0) queue[0] = src;
}

// relax one edge.  state: 0 = not queued, 1 = in the next near frontier, 2 = waiting beyond the
// threshold ("far").  The next frontier is compacted from the state array after the round
// (k_sssp_compact): appending winners to a queue from here funnels millions of atomics per round
// through one counter.  state == NULL (multi-GPU): the next frontier is derived from the
// min-reduced distances instead.
__device__ __forceinline__ bool sssp_relax_edge(unsigned long long *dist, uint32_t *state, uint32_t v, double nd,
                                                unsigned long long thresh)
{
    const unsigned long long nb = (unsigned long long)__double_as_longlong(nd);
    if (nb >= dist[v]) return false;
    const unsigned long long old = atomicMin(&dist[v], nb);
    if (nb >= old || state == nullptr) return false;
    if (nb < thresh) { if (state[v] != 1u) state[v] = 1u; return true; }
    atomicCAS(&state[v], 0u, 2u);
    return false;
}

__device__ __forceinline__ void sssp_append(bool won, uint32_t v, uint32_t *next_q, SsspCounters *cnt)
{
    const unsigned mask = __ballot_sync(FULL, won);
    if (mask == 0) return;
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(&cnt->next_count, (unsigned long long)__popc(mask));
    base = __shfl_sync(FULL, base, 0);
    if (won) next_q[base + __popc(mask & ((1u << lane_id()) - 1u))] = v;
}

__global__ void __launch_bounds__(256)
k_sssp_relax(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const double *__restrict__ w,
             const uint32_t *__restrict__ queue, uint64_t qn, uint64_t v0, uint64_t v1,
             unsigned long long *__restrict__ dist, uint32_t *__restrict__ inq, uint32_t *__restrict__ next_q,
             uint32_t *__restrict__ big_row, uint64_t *__restrict__ big_begin, SsspCounters *__restrict__ cnt,
             unsigned long long thresh)
{
    // the frontier queue is replicated (in any order); a rank expands the vertices of its row block [v0, v1)
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long relaxed = 0;
    for (; wid < qn; wid += nw) {
        const uint32_t u = queue[wid];
        if (u < v0 || u >= v1) continue;
        const uint64_t a = rowptr[u], b = rowptr[u + 1];
        if (b - a > SSSP_BIG) {
            const uint64_t nch = (b - a + CHUNK - 1) / CHUNK;
            unsigned long long pos = 0;
            if (lane_id() == 0) pos = atomicAdd(&cnt->big_count, (unsigned long long)nch);
            pos = __shfl_sync(FULL, pos, 0);
            for (uint64_t k = lane_id(); k < nch; k += 32) { big_row[pos + k] = u; big_begin[pos + k] = a + k * CHUNK; }
            continue;
        }
        const double du = __longlong_as_double((long long)dist[u]);
        for (uint64_t base = a; base < b; base += 32) {
            const uint64_t e = base + lane_id();
            if (e < b) {
                sssp_relax_edge(dist, inq, ld_stream(col + e), du + ld_stream_f64(w + e), thresh);
                relaxed++;
            }
        }
    }
    relaxed = warp_sum(relaxed);
    if (lane_id() == 0 && relaxed) atomicAdd(&cnt->relaxed, relaxed);
}

__global__ void __launch_bounds__(256)
k_sssp_relax_big(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const double *__restrict__ w,
                 const uint32_t *__restrict__ big_row, const uint64_t *__restrict__ big_begin,
                 unsigned long long *__restrict__ dist, uint32_t *__restrict__ inq, uint32_t *__restrict__ next_q,
                 SsspCounters *__restrict__ cnt, unsigned long long thresh)
{
    const unsigned long long nbig = cnt->big_count;
    unsigned long long relaxed = 0;
    for (unsigned long long c = blockIdx.x; c < nbig; c += gridDim.x) {
        const uint32_t u = big_row[c];
        const uint64_t b0 = big_begin[c];
        const uint64_t row_end = rowptr[u + 1];
        const uint64_t e_end = (b0 + CHUNK < row_end) ? b0 + CHUNK : row_end;
        const double du = __longlong_as_double((long long)dist[u]);
        for (uint64_t base = b0; base < e_end; base += 256) {
            const uint64_t e = base + threadIdx.x;
            if (e < e_end) {
                sssp_relax_edge(dist, inq, ld_stream(col + e), du + ld_stream_f64(w + e), thresh);
                relaxed++;
            }
        }
    }
    relaxed = warp_sum(relaxed);
    if (lane_id() == 0 && relaxed) atomicAdd(&cnt->relaxed, relaxed);
}

// Frontier members drop their "queued" flag in a kernel of their own, BEFORE the relax kernels:
// any improvement of u that lands while u is being expanded (possibly from a distance read a
// moment too early) then finds the flag clear and re-queues u for the next round.
__global__ void k_sssp_clear(const uint32_t *__restrict__ queue, uint64_t qn, uint32_t *__restrict__ inq)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < qn; i += stride) inq[queue[i]] = 0;
}

// next near frontier = vertices in state 1; one global atomic per 256 vertices that hold any
__global__ void __launch_bounds__(256)
k_sssp_compact(const uint32_t *__restrict__ state, uint64_t n, uint32_t *__restrict__ queue, SsspCounters *__restrict__ cnt)
{
    __shared__ unsigned s_cnt[8];
    __shared__ unsigned long long s_base;
    const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
    const uint64_t nround = (n + 255) & ~255ull;
    for (uint64_t base = (uint64_t)blockIdx.x * 256; base < nround; base += (uint64_t)gridDim.x * 256) {
        const uint64_t v = base + threadIdx.x;
        const bool in = v < n && state[v] == 1u;
        const unsigned mask = __ballot_sync(FULL, in);
        if (lane == 0) s_cnt[wib] = __popc(mask);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
#pragma unroll
            for (int i = 0; i < 8; i++) { const unsigned c = s_cnt[i]; s_cnt[i] = tot; tot += c; }
            s_base = tot ? atomicAdd(&cnt->next_count, (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (in) queue[s_base + s_cnt[wib] + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)v;
        __syncthreads();
    }
}

__global__ void k_sssp_weight_sum(const double *__restrict__ w, uint64_t m, double *__restrict__ sum)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    double s = 0.0;
    for (; e < m; e += stride) s += w[e];
    s = warp_sum(s);
    if (lane_id() == 0) atomicAdd(sum, s);
}

// T advanced: far vertices now below it move to the near queue; the rest is counted and its
// smallest distance recorded so that empty buckets can be skipped
__global__ void k_sssp_collect_far(const unsigned long long *__restrict__ dist, uint32_t *__restrict__ state, uint64_t n,
                                   unsigned long long thresh, uint32_t *__restrict__ queue, SsspCounters *__restrict__ cnt)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nround = (n + 31) & ~31ull;
    unsigned long long far = 0, fmin = ~0ull;
    for (; v < nround; v += stride) {
        bool take = false;
        if (v < n && state[v] == 2u) {
            const unsigned long long d = dist[v];
            if (d < thresh) { take = true; state[v] = 1u; }
            else { far++; fmin = d < fmin ? d : fmin; }
        }
        sssp_append(take, (uint32_t)v, queue, cnt);
    }
    far = warp_sum(far);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(FULL, fmin, o); fmin = x < fmin ? x : fmin; }
    if (lane_id() == 0 && far) { atomicAdd(&cnt->far_count, far); atomicMin(&cnt->far_min, fmin); }
}

// Multi-GPU: vertices whose min-reduced distance dropped during the round form the next frontier
// (same set on every rank: dist and prev are replicated).
__global__ void k_sssp_diff(const unsigned long long *__restrict__ dist, unsigned long long *__restrict__ prev, uint64_t n,
                            uint32_t *__restrict__ next_q, SsspCounters *__restrict__ cnt)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nround = (n + 31) & ~31ull;
    for (; v < nround; v += stride) {
        bool ch = false;
        if (v < n) {
            const unsigned long long d = dist[v];
            ch = d < prev[v];
            if (ch) prev[v] = d;
        }
        sssp_append(ch, (uint32_t)v, next_q, cnt);
    }
}

__global__ void k_sssp_out(const unsigned long long *__restrict__ dist, uint64_t n, double *__restrict__ out)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) out[v] = __longlong_as_double((long long)dist[v]);
}

} // namespace gx

using namespace gx;

extern "C" int gx_sssp(gx_graph *g, uint64_t src, double *dist_host)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        GX_REQUIRE(src < g->n, "source vertex out of range");
        GX_REQUIRE(g->weighted, "SSSP needs a weighted graph (graph.mtx of type real / GrB_FP64)");
        Context &c = ctx();
        c.timing = gx_timing{};
        const uint64_t n = g->n, m = g->m;
        {
            PhaseTimer tb(&c.timing.build_ms);
            ensure_plan(g->out, n);
        }
        const Partition &part = g->out.plan.part;
        if (!g->have_mean_weight && m) {
            DevBuf<double> sum(1);
            sum.zero();
            GX_LAUNCH(k_sssp_weight_sum, grid_persistent(8), 256, 0, g->out.w.p, m, sum.p);
            double h = 0;
            read_back(&h, sum.p, sizeof(h));
            g->mean_weight = h / (double)m;
            g->have_mean_weight = true;
        }
        g->res_f64.alloc(n);
        DevBuf<unsigned long long> dist(n);
        DevBuf<uint32_t> inq(n), q0(n), q1(n);
        DevBuf<unsigned long long> prev(multi() ? n : 0);
        const uint64_t big_cap = m / CHUNK + m / SSSP_BIG + 16;
        DevBuf<uint32_t> big_row(big_cap);
        DevBuf<uint64_t> big_begin(big_cap);
        DevBuf<SsspCounters> cnt(1);
        uint64_t relaxed = 0;
        uint32_t rounds = 0;
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            GX_LAUNCH(k_sssp_init, grid_persistent(8), 256, 0, dist.p, n, (uint32_t)src, q0.p, inq.p);
            uint32_t *queue = q0.p, *next_q = q1.p;
            uint64_t qn = 1;
            if (multi()) GX_CUDA(cudaMemcpyAsync(prev.p, dist.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c.stream));
            uint32_t *inq_p = multi() ? nullptr : inq.p;
            // bucket width: a few average edge weights per average degree (Davidson et al.'s near-far rule);
            // several GPUs run plain sweeps (threshold = +inf)
            double delta = 8.0 * (double)g->mean_weight * (double)n / (double)(m ? m : 1);
            if (const char *e = getenv("GX_SSSP_DELTA")) delta = atof(e); // tuning knob
            const bool buckets = !multi() && delta > 0.0;
            double T = buckets ? delta : INFINITY;
            auto bits = [](double x) { unsigned long long b; memcpy(&b, &x, sizeof(b)); return b; };
            for (;;) {
                while (qn) {
                    cnt.zero();
                    // each rank relaxes the out-edges of the frontier vertices in its row block, on its replica
                    if (!multi()) GX_LAUNCH(k_sssp_clear, grid_for(qn, 256), 256, 0, queue, qn, inq.p);
                    GX_LAUNCH(k_sssp_relax, grid_for(qn * 32, 256), 256, 0, g->out.rowptr.p, g->out.col.p, g->out.w.p, queue, qn,
                              part.lo, part.hi, dist.p, inq_p, next_q, big_row.p, big_begin.p, cnt.p, bits(T));
                    GX_LAUNCH(k_sssp_relax_big, grid_persistent(4), 256, 0, g->out.rowptr.p, g->out.col.p, g->out.w.p, big_row.p,
                              big_begin.p, dist.p, inq_p, next_q, cnt.p, bits(T));
                    if (!multi()) GX_LAUNCH(k_sssp_compact, grid_persistent(8), 256, 0, inq.p, n, next_q, cnt.p);
                    if (multi()) {
                        // non-negative doubles order like their bit patterns: min over the replicas, then diff
                        allreduce(dist.p, n, Dt::U64, Red::Min);
                        GX_CUDA(cudaMemsetAsync(&cnt.p->next_count, 0, sizeof(unsigned long long), c.stream));
                        GX_LAUNCH(k_sssp_diff, grid_persistent(8), 256, 0, dist.p, prev.p, n, next_q, cnt.p);
                    }
                    SsspCounters h;
                    read_back(&h, cnt.p, sizeof(h));
                    qn = h.next_count;
                    relaxed += h.relaxed;
                    uint32_t *t = queue; queue = next_q; next_q = t;
                    rounds++;
                }
                if (!buckets) break;
                // the near queue ran dry: advance the threshold and collect what now lies below it
                T += delta;
                SsspCounters h;
                for (;;) {
                    cnt.zero();
                    GX_CUDA(cudaMemsetAsync(&cnt.p->far_min, 0xFF, sizeof(unsigned long long), c.stream));
                    GX_LAUNCH(k_sssp_collect_far, grid_persistent(8), 256, 0, dist.p, inq.p, n, bits(T), queue, cnt.p);
                    read_back(&h, cnt.p, sizeof(h));
                    if (h.next_count || !h.far_count) break;
                    double fmin;
                    memcpy(&fmin, &h.far_min, sizeof(fmin));
                    T = fmin + delta; // skip the empty buckets
                }
                qn = h.next_count;
                if (!qn) break;
            }
            GX_LAUNCH(k_sssp_out, grid_persistent(8), 256, 0, dist.p, n, g->res_f64.p);
        }
        c.timing.iterations = rounds;
        c.timing.edges_inspected = relaxed;
        c.timing.algorithmic_bytes = 12 * m + 8 * (n + 1) + 16 * n;
        if (dist_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(dist_host, g->res_f64.p, n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
