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
// each time its distance drops).  On one GPU the run is therefore delta-stepping proper (Meyer &
// Sanders; LAGraph's own algorithm): buckets [T - delta, T) are settled in order.  Inside a bucket
// only LIGHT entries (w <= delta) are relaxed, again whenever a member's distance drops; once the
// bucket is stable the HEAVY entries of everything it expanded are relaxed ONCE, with final
// distances -- a heavy entry cannot reach back into the bucket.  A cached copy of the adjacency
// keeps every row's light entries in front (stable partition, built once per graph and delta:
// SsspCache).  Targets beyond T only get their state marked ("far") and are collected by one
// pass over the state array when T advances.  delta only schedules work -- the fix-point, hence
// every bit of the result, is the same.  Several GPUs run the same delta-stepping with owner-held distances:
// a rank expands the due vertices of its row block, improvements to other ranks' vertices are forwarded by
// system-scope atomicMin over NVLink peer mappings, and {queue size, smallest waiting distance} are exchanged
// once per round through peer-mapped mailboxes (sssp_multi_delta below).
// Algorithmic bytes (one-pass bound): 12m + 8(n+1) + 16n.
#include <cub/device/device_scan.cuh>
#include <thrust/iterator/transform_iterator.h>

#include <cmath>
#include <cstdlib>
#include <cstring>

#include "graph.cuh"

namespace gx {

constexpr uint32_t SSSP_BIG = 4096;
constexpr unsigned long long INF_BITS = 0x7FF0000000000000ull;

struct SsspCounters { unsigned long long next_count, big_count, relaxed, far_count, far_min, r_count; };

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
//
// Several GPUs with delta-stepping (P != NULL): every rank holds a full-length distance array, mapped into all
// peers.  A rank's copy is AUTHORITATIVE for the vertices of its row block and a filter (a stale upper bound)
// for all others: an improvement that passes the local atomicMin is forwarded to the owner's copy by a
// system-scope atomicMin over NVLink -- no all-reduce of the distances, only improvements travel.
struct SsspPeers {
    unsigned long long *dist[MAX_PEERS]; // rank r's distance array as seen from here
    uint64_t bound[MAX_PEERS + 1];       // row-block boundaries
    int nranks, rank;
};

__device__ __forceinline__ bool sssp_relax_edge(unsigned long long *dist, uint32_t *state, uint32_t v, double nd,
                                                unsigned long long thresh, const SsspPeers *__restrict__ P = nullptr)
{
    const unsigned long long nb = (unsigned long long)__double_as_longlong(nd);
    if (nb >= dist[v]) return false;
    if (P) {
        const unsigned long long was = atomicMin_system(&dist[v], nb);
        if (nb >= was) return false;
        int o = 0;
        while (o + 1 < P->nranks && v >= P->bound[o + 1]) o++;
        if (o != P->rank) atomicMin_system(&P->dist[o][v], nb);
        return false;
    }
    const unsigned long long old = atomicMin(&dist[v], nb);
    if (nb >= old || state == nullptr) return false;
    if (nb < thresh) { if (state[v] != 1u) state[v] = 1u; return true; }
    atomicCAS(&state[v], 0u, 2u);
    return false;
}

// read-only adjacency loads the compiler may batch (ld_stream is `asm volatile`, i.e. strictly ordered)
__device__ __forceinline__ uint32_t ld_adj(const uint32_t *p)
{
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_adj_f64(const double *p)
{
    double v;
    asm("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// Relax entries e0, e0 + step, ... (up to 4, below `hi`): the 4 entry loads and then the 4 distance
// loads are in flight together; only actual improvements go on to the atomic.
// Entries with a weight <= min_w are skipped (the heavy phase walks whole rows of the graph's own adjacency and
// leaves out the light entries; min_w < 0 relaxes everything).
__device__ __forceinline__ unsigned sssp_relax4(const uint32_t *__restrict__ col, const double *__restrict__ w, uint64_t e0,
                                                uint64_t step, uint64_t hi, double du, unsigned long long *dist,
                                                uint32_t *state, unsigned long long thresh, double min_w,
                                                const SsspPeers *__restrict__ P = nullptr)
{
    uint32_t v[4];
    double nd[4];
    unsigned long long dv[4];
    bool ok[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const uint64_t e = e0 + (uint64_t)j * step;
        ok[j] = e < hi;
        const double we = ok[j] ? ld_adj_f64(w + e) : 0.0;
        ok[j] = ok[j] && we > min_w;
        v[j] = ok[j] ? ld_adj(col + e) : 0u;
        nd[j] = du + we;
    }
#pragma unroll
    for (int j = 0; j < 4; j++) dv[j] = ok[j] ? dist[v[j]] : 0ull;
    unsigned cntd = 0;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (!ok[j]) continue;
        cntd++;
        if ((unsigned long long)__double_as_longlong(nd[j]) < dv[j]) sssp_relax_edge(dist, state, v[j], nd[j], thresh, P);
    }
    return cntd;
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

// ---- delta-stepping on one GPU: light / heavy entries -----------------------------------------
// The light entries (w <= delta: ~10 % of them at the default bucket width) are compacted into an adjacency of
// their own, row by row; the heavy phase walks the graph's own rows and skips the light ones.  Building this reads
// the weights twice and the column ids of the light entries once (~27 bytes per entry), against a full partitioned
// copy of every row (12 bytes written per entry plus row ids: 4x the traffic and 10x the allocation) before.
struct SsspCache {
    double delta = 0.0;
    uint64_t ml = 0;          // light entries
    DevBuf<uint64_t> lrowptr; // n+1
    DevBuf<uint32_t> lcol;    // ml
    DevBuf<double> lw;        // ml
};

struct LightFlag {
    double delta;
    __host__ __device__ __forceinline__ uint32_t operator()(const double &x) const { return x <= delta ? 1u : 0u; }
};

// lrowptr[v] = light entries before row v among the entries [e0, e1) (rows [v0, v1]: the whole graph on one GPU,
// the rank's row block on several); L = exclusive prefix count of light entries over that entry range
__global__ void k_sssp_lrowptr(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ L, const double *__restrict__ w,
                               uint64_t v0, uint64_t v1, uint64_t e0, uint64_t e1, double delta, uint64_t *__restrict__ lrowptr)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t total = (uint64_t)L[e1 - e0 - 1] + (w[e1 - 1] <= delta ? 1u : 0u);
    for (; v <= v1; v += stride) {
        const uint64_t e = rowptr[v];
        lrowptr[v] = e < e1 ? (uint64_t)L[e - e0] : total;
    }
}

__global__ void k_sssp_compact_light(const uint32_t *__restrict__ col, const double *__restrict__ w, const uint32_t *__restrict__ L,
                                     uint64_t e0, uint64_t e1, double delta, uint32_t *__restrict__ lcol, double *__restrict__ lw)
{
    uint64_t e = e0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < e1; e += stride) {
        const double we = ld_stream_f64(w + e);
        if (we <= delta) { const uint32_t d = L[e - e0]; lcol[d] = col[e]; lw[d] = we; }
    }
}

// G lanes per queue vertex; HEAVY selects which part of the row is relaxed.  Ranges above 32 * G
// entries are re-queued as CHUNK pieces for whole CTAs (k_sssp_relax_pieces).  Light expansion stamps
// the vertex with the current epoch: the stamped vertices are the bucket's members whose heavy
// entries are still due.
// (rowptr, col, w): the light adjacency for the light rounds, the graph's own for the heavy phase (min_w = delta)
template <int G, bool HEAVY>
__global__ void __launch_bounds__(256)
k_sssp_expand(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col,
              const double *__restrict__ w, const uint32_t *__restrict__ queue, const unsigned long long *__restrict__ qn_p,
              unsigned long long *__restrict__ dist, uint32_t *__restrict__ state, uint32_t *__restrict__ stamp, uint32_t epoch,
              uint32_t *__restrict__ big_row, uint64_t *__restrict__ big_begin, uint64_t *__restrict__ big_end,
              SsspCounters *__restrict__ cnt, unsigned long long thresh, double min_w, const SsspPeers *__restrict__ P)
{
    const unsigned sub = threadIdx.x & (G - 1);
    uint64_t gi = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
    const uint64_t ng = ((uint64_t)gridDim.x * blockDim.x) / G;
    const uint64_t qn = *qn_p;
    unsigned long long relaxed = 0;
    for (; gi < qn; gi += ng) {
        const uint32_t u = queue[gi];
        const uint64_t lo = rowptr[u], hi = rowptr[u + 1];
        if (!HEAVY && stamp && sub == 0) stamp[u] = epoch;
        if (hi - lo > 32u * G) { // more than 32 trips of the group: whole CTAs take it over
            const uint64_t nch = (hi - lo + CHUNK - 1) / CHUNK;
            const unsigned gmask = G == 32 ? FULL : (((1u << (G & 31)) - 1u) << ((lane_id() / G) * G)); // groups diverge here
            unsigned long long pos = 0;
            if (sub == 0) pos = atomicAdd(&cnt->big_count, (unsigned long long)nch);
            pos = __shfl_sync(gmask, pos, (lane_id() / G) * G);
            for (uint64_t k = sub; k < nch; k += G) { big_row[pos + k] = u; big_begin[pos + k] = lo + k * CHUNK; big_end[pos + k] = hi; }
            continue;
        }
        const double du = __longlong_as_double((long long)dist[u]);
        for (uint64_t e = lo + sub; e < hi; e += 4 * G) relaxed += sssp_relax4(col, w, e, G, hi, du, dist, state, thresh, min_w, P);
    }
    relaxed = warp_sum(relaxed);
    if (lane_id() == 0 && relaxed) atomicAdd(&cnt->relaxed, relaxed);
}

__global__ void __launch_bounds__(256)
k_sssp_relax_pieces(const uint32_t *__restrict__ col, const double *__restrict__ w, const uint32_t *__restrict__ big_row,
                    const uint64_t *__restrict__ big_begin, const uint64_t *__restrict__ big_end,
                    unsigned long long *__restrict__ dist, uint32_t *__restrict__ state, SsspCounters *__restrict__ cnt,
                    unsigned long long thresh, double min_w, const SsspPeers *__restrict__ P)
{
    const unsigned long long nbig = cnt->big_count;
    unsigned long long relaxed = 0;
    for (unsigned long long c = blockIdx.x; c < nbig; c += gridDim.x) {
        const uint64_t b0 = big_begin[c], end = big_end[c];
        const uint64_t e_end = (b0 + CHUNK < end) ? b0 + CHUNK : end;
        const double du = __longlong_as_double((long long)dist[big_row[c]]);
        for (uint64_t e = b0 + threadIdx.x; e < e_end; e += 4 * 256) relaxed += sssp_relax4(col, w, e, 256, e_end, du, dist, state, thresh, min_w, P);
    }
    relaxed = warp_sum(relaxed);
    if (lane_id() == 0 && relaxed) atomicAdd(&cnt->relaxed, relaxed);
}

// Block-level compaction, 4 consecutive vertices per thread: `take` bit j selects vertex v4 + j.  One
// global atomic per 1024 vertices that hold any (a per-warp atomic on one counter costs ~0.25 us each
// once hundreds of thousands of warps queue up behind it).  All 256 threads of the CTA must call.
__device__ __forceinline__ void block_compact4(unsigned take, uint64_t v4, uint32_t *__restrict__ queue,
                                               unsigned long long *__restrict__ count)
{
    __shared__ unsigned s_cnt[8];
    __shared__ unsigned long long s_base;
    // most 1024-vertex blocks of a round hold nothing to take: one barrier settles that (the scans run over all n
    // vertices every round and were a tenth of the run at half of the streaming rate)
    if (!__syncthreads_or((int)take)) return;
    const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
    const unsigned mine = __popc(take);
    unsigned incl = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned up = __shfl_up_sync(FULL, incl, d);
        if (lane >= (unsigned)d) incl += up;
    }
    if (lane == 31) s_cnt[wib] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned tot = 0;
#pragma unroll
        for (int i = 0; i < 8; i++) { const unsigned c = s_cnt[i]; s_cnt[i] = tot; tot += c; }
        s_base = tot ? atomicAdd(count, (unsigned long long)tot) : 0ull;
    }
    __syncthreads();
    unsigned long long pos = s_base + s_cnt[wib] + (incl - mine);
#pragma unroll
    for (int j = 0; j < 4; j++)
        if ((take >> j) & 1u) queue[pos++] = (uint32_t)(v4 + j);
    __syncthreads();
}

// queue of the vertices with arr[v] == value (arr is 16-byte aligned: a stream-ordered allocation)
__global__ void __launch_bounds__(256)
k_sssp_compact_eq(const uint32_t *__restrict__ arr, uint32_t value, uint64_t n, uint32_t *__restrict__ queue,
                  unsigned long long *__restrict__ count)
{
    const uint64_t nround = (n + 1023) & ~1023ull;
    for (uint64_t base = (uint64_t)blockIdx.x * 1024; base < nround; base += (uint64_t)gridDim.x * 1024) {
        const uint64_t v4 = base + 4ull * threadIdx.x;
        unsigned take = 0;
        if (v4 + 4 <= n) {
            const uint4 x = *(const uint4 *)(arr + v4);
            take = (x.x == value ? 1u : 0u) | (x.y == value ? 2u : 0u) | (x.z == value ? 4u : 0u) | (x.w == value ? 8u : 0u);
        } else {
            for (int j = 0; j < 4; j++)
                if (v4 + j < n && arr[v4 + j] == value) take |= 1u << j;
        }
        block_compact4(take, v4, queue, count);
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
__global__ void __launch_bounds__(256)
k_sssp_collect_far(const unsigned long long *__restrict__ dist, uint32_t *__restrict__ state, uint64_t n,
                   unsigned long long thresh, uint32_t *__restrict__ queue, SsspCounters *__restrict__ cnt)
{
    const uint64_t nround = (n + 1023) & ~1023ull;
    unsigned long long far = 0, fmin = ~0ull;
    for (uint64_t base = (uint64_t)blockIdx.x * 1024; base < nround; base += (uint64_t)gridDim.x * 1024) {
        const uint64_t v4 = base + 4ull * threadIdx.x;
        unsigned take = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t v = v4 + j;
            if (v < n && state[v] == 2u) {
                const unsigned long long d = dist[v];
                if (d < thresh) { take |= 1u << j; state[v] = 1u; }
                else { far++; fmin = d < fmin ? d : fmin; }
            }
        }
        block_compact4(take, v4, queue, &cnt->next_count);
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

namespace gx {

static SsspCache *build_sssp_cache(gx_graph *g, double delta)
{
    const uint64_t n = g->n;
    SsspCache *sc = new SsspCache();
    sc->delta = delta;
    sc->lrowptr.alloc(n + 1);
    // several GPUs: a rank expands the rows of its block only, so only their light entries are compacted
    const Partition &part = g->out.plan.part;
    const uint64_t v0 = multi() ? part.lo : 0, v1 = multi() ? part.hi : n;
    uint64_t ends[2] = {0, g->m};
    if (multi()) {
        read_back(&ends[0], g->out.rowptr.p + v0, sizeof(uint64_t));
        read_back(&ends[1], g->out.rowptr.p + v1, sizeof(uint64_t));
    }
    const uint64_t e0 = ends[0], e1 = ends[1], cnt = e1 - e0;
    GX_REQUIRE(cnt < 0xFFFFFFFFull, "SSSP light/heavy split needs fewer than 2^32 entries per rank");
    if (!cnt) { sc->lrowptr.zero(); sc->lcol.alloc(1); sc->lw.alloc(1); return sc; }
    DevBuf<uint32_t> L(cnt);
    {
        // L[e] = light entries before e: the flags are computed on the fly from the weights (no flag array)
        auto flags = thrust::make_transform_iterator((const double *)g->out.w.p + e0, LightFlag{delta});
        size_t tb = 0;
        GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flags, L.p, (int64_t)cnt, ctx().stream));
        DevBuf<char> tmp(tb);
        GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flags, L.p, (int64_t)cnt, ctx().stream));
        count_launch();
    }
    GX_LAUNCH(k_sssp_lrowptr, grid_persistent(8), 256, 0, g->out.rowptr.p, L.p, g->out.w.p, v0, v1, e0, e1, delta, sc->lrowptr.p);
    read_back(&sc->ml, sc->lrowptr.p + v1, sizeof(uint64_t));
    sc->lcol.alloc(sc->ml ? sc->ml : 1);
    sc->lw.alloc(sc->ml ? sc->ml : 1);
    GX_LAUNCH(k_sssp_compact_light, grid_persistent(8), 256, 0, g->out.col.p, g->out.w.p, L.p, e0, e1, delta, sc->lcol.p, sc->lw.p);
    return sc;
}

// one GPU: delta-stepping with light / heavy entries (see the header comment)
static void sssp_delta_stepping(gx_graph *g, const SsspCache &sc, uint64_t src, uint64_t &relaxed, uint32_t &rounds)
{
    Context &c = ctx();
    const uint64_t n = g->n, m = g->m;
    const double delta = sc.delta;
    DevBuf<unsigned long long> dist(n);
    DevBuf<uint32_t> state(n), stamp(n), q0(n), q1(n);
    const uint64_t big_cap = m / CHUNK + m / 128 + 16; // pieces of one launch: ranges longer than 32 * G entries, G >= 4
    DevBuf<uint32_t> big_row(big_cap);
    DevBuf<uint64_t> big_begin(big_cap), big_end(big_cap);
    DevBuf<SsspCounters> cnt(1);
    DevBuf<unsigned long long> qn_dev(1);
    auto bits = [](double x) { unsigned long long b; memcpy(&b, &x, sizeof(b)); return b; };
    const uint64_t *rp = g->out.rowptr.p;
    GX_LAUNCH(k_sssp_init, grid_persistent(8), 256, 0, dist.p, n, (uint32_t)src, q0.p, state.p);
    stamp.zero();
    uint32_t *queue = q0.p, *next_q = q1.p;
    uint64_t qn = 1;
    uint32_t epoch = 1;
    double T = delta;
    // lanes per queue vertex (tuning knobs): light rows hold a handful of entries, whole rows a few dozen -- several
    // vertices per warp keep more dependent row-offset -> entry -> distance chains in flight than one vertex per warp
    int lg = 8, hg = 16;
    if (const char *e = getenv("GX_SSSP_LG")) lg = atoi(e);
    if (const char *e = getenv("GX_SSSP_HG")) hg = atoi(e);
    bool expanded = false; // some vertex was light-expanded since the last heavy phase
    for (;;) {
        // ---- light rounds: the bucket's members relax their light entries until nothing below T moves
        while (qn) {
            cnt.zero();
            GX_CUDA(cudaMemcpyAsync(qn_dev.p, &qn, sizeof(qn), cudaMemcpyHostToDevice, c.stream));
            GX_LAUNCH(k_sssp_clear, grid_for(qn, 256), 256, 0, queue, qn, state.p);
            // (grid capped: every warp ends with one atomic on the shared counters)
            const unsigned g_light = grid_for(qn * (unsigned)lg, 256) < grid_persistent(8) ? grid_for(qn * (unsigned)lg, 256) : grid_persistent(8);
            if (lg == 4)
                GX_LAUNCH((k_sssp_expand<4, false>), g_light, 256, 0, sc.lrowptr.p, sc.lcol.p, sc.lw.p, queue, qn_dev.p,
                          dist.p, state.p, stamp.p, epoch, big_row.p, big_begin.p, big_end.p, cnt.p, bits(T), -1.0, nullptr);
            else
                GX_LAUNCH((k_sssp_expand<8, false>), g_light, 256, 0, sc.lrowptr.p, sc.lcol.p, sc.lw.p, queue, qn_dev.p,
                          dist.p, state.p, stamp.p, epoch, big_row.p, big_begin.p, big_end.p, cnt.p, bits(T), -1.0, nullptr);
            GX_LAUNCH(k_sssp_relax_pieces, grid_persistent(8), 256, 0, sc.lcol.p, sc.lw.p, big_row.p, big_begin.p, big_end.p, dist.p,
                      state.p, cnt.p, bits(T), -1.0, nullptr);
            GX_LAUNCH(k_sssp_compact_eq, grid_persistent(8), 256, 0, state.p, 1u, n, next_q, &cnt.p->next_count);
            SsspCounters h;
            read_back(&h, cnt.p, sizeof(h)); // also orders the host-side qn against its async copy
            qn = h.next_count;
            relaxed += h.relaxed;
            uint32_t *t = queue; queue = next_q; next_q = t;
            rounds++;
            expanded = true;
        }
        // ---- the bucket is stable: heavy entries of everything it expanded, once, with final distances
        if (expanded) {
            cnt.zero();
            GX_LAUNCH(k_sssp_compact_eq, grid_persistent(8), 256, 0, stamp.p, epoch, n, next_q, &cnt.p->r_count);
            if (hg == 8)
                GX_LAUNCH((k_sssp_expand<8, true>), grid_persistent(8), 256, 0, rp, g->out.col.p, g->out.w.p, next_q, &cnt.p->r_count,
                          dist.p, state.p, stamp.p, epoch, big_row.p, big_begin.p, big_end.p, cnt.p, bits(T), delta, nullptr);
            else if (hg == 16)
                GX_LAUNCH((k_sssp_expand<16, true>), grid_persistent(8), 256, 0, rp, g->out.col.p, g->out.w.p, next_q, &cnt.p->r_count,
                          dist.p, state.p, stamp.p, epoch, big_row.p, big_begin.p, big_end.p, cnt.p, bits(T), delta, nullptr);
            else
                GX_LAUNCH((k_sssp_expand<32, true>), grid_persistent(8), 256, 0, rp, g->out.col.p, g->out.w.p, next_q, &cnt.p->r_count,
                          dist.p, state.p, stamp.p, epoch, big_row.p, big_begin.p, big_end.p, cnt.p, bits(T), delta, nullptr);
            GX_LAUNCH(k_sssp_relax_pieces, grid_persistent(8), 256, 0, g->out.col.p, g->out.w.p, big_row.p, big_begin.p, big_end.p, dist.p,
                      state.p, cnt.p, bits(T), delta, nullptr);
            // a heavy entry adds more than delta to a distance >= T - delta, so nothing lands below T;
            // should rounding ever say otherwise, the vertex is in state 1 and the bucket simply goes on
            GX_LAUNCH(k_sssp_compact_eq, grid_persistent(8), 256, 0, state.p, 1u, n, queue, &cnt.p->next_count);
            SsspCounters h;
            read_back(&h, cnt.p, sizeof(h));
            relaxed += h.relaxed;
            rounds++;
            epoch++;
            expanded = false;
            qn = h.next_count;
            if (qn) continue;
        }
        // ---- advance the threshold and collect what now lies below it
        T += delta;
        SsspCounters h;
        for (;;) {
            cnt.zero();
            GX_CUDA(cudaMemsetAsync(&cnt.p->far_min, 0xFF, sizeof(unsigned long long), c.stream));
            GX_LAUNCH(k_sssp_collect_far, grid_persistent(8), 256, 0, dist.p, state.p, n, bits(T), queue, cnt.p);
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

// ---- several GPUs: delta-stepping, owner-computes, improvements forwarded by peer atomics ---------------------
// Per owned vertex two marks: the distance at which it was last light- / heavy-expanded.  A vertex is due for
// the light rounds of the bucket below T when its distance is < T and < its light mark; for the heavy phase that
// closes the bucket when its distance is < T and < its heavy mark.  The scan that builds a round's queue also
// returns the smallest distance >= T that is still waiting, so empty buckets are skipped.
struct SsspRound { unsigned long long qn, neg_far_min, relaxed; }; // the first two are max-reduced over the ranks

__global__ void k_sssp_m_init(unsigned long long *__restrict__ dist, uint64_t n, uint32_t src, unsigned long long *__restrict__ ldone,
                              unsigned long long *__restrict__ hdone, uint64_t own)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) {
        dist[v] = (v == src) ? 0ull : INF_BITS;
        if (v < own) { ldone[v] = INF_BITS; hdone[v] = INF_BITS; }
    }
}

template <bool HEAVY>
__global__ void __launch_bounds__(256)
k_sssp_m_build(const unsigned long long *__restrict__ dist, unsigned long long *__restrict__ done, uint64_t v0, uint64_t v1,
               unsigned long long thresh, uint32_t *__restrict__ queue, SsspCounters *__restrict__ cnt)
{
    const uint64_t own = v1 - v0, nround = (own + 1023) & ~1023ull;
    unsigned long long fmin = ~0ull;
    for (uint64_t base = (uint64_t)blockIdx.x * 1024; base < nround; base += (uint64_t)gridDim.x * 1024) {
        const uint64_t i4 = base + 4ull * threadIdx.x;
        unsigned take = 0;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const uint64_t i = i4 + j;
            if (i < own) {
                const unsigned long long d = dist[v0 + i];
                if (d < done[i]) {
                    if (d < thresh) { take |= 1u << j; done[i] = d; }
                    else if (!HEAVY) fmin = d < fmin ? d : fmin;
                }
            }
        }
        block_compact4(take, v0 + i4, queue, &cnt->next_count);
    }
    if (!HEAVY) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { const unsigned long long x = __shfl_xor_sync(FULL, fmin, o); fmin = x < fmin ? x : fmin; }
        if (lane_id() == 0 && fmin != ~0ull) atomicMin(&cnt->far_min, fmin);
    }
}

__global__ void k_sssp_m_pack(const SsspCounters *__restrict__ cnt, SsspRound *__restrict__ r)
{
    r->qn = cnt->next_count;
    r->neg_far_min = ~cnt->far_min;
    r->relaxed = cnt->relaxed;
}

// the round's exchange through the peer mailboxes (comm.cuh): {queue size, ~smallest waiting distance} max-reduced
// over the ranks, result into host-mapped memory; the counters are reset for the next round on the way
__global__ void __launch_bounds__(32) k_sssp_m_exchange(MailTable t, SsspCounters *__restrict__ cnt, unsigned long long seq,
                                                         unsigned long long *__restrict__ host_out)
{
    const unsigned long long qn = cnt->next_count, nf = ~cnt->far_min, relaxed = cnt->relaxed;
    __syncwarp();
    if (threadIdx.x == 0) {
        cnt->next_count = 0; cnt->big_count = 0; cnt->relaxed = 0; cnt->far_count = 0; cnt->r_count = 0;
        cnt->far_min = ~0ull;
    }
    peer_mail_exchange(t, qn, nf, relaxed, seq, host_out);
}

__global__ void k_sssp_m_out(const unsigned long long *__restrict__ dist, uint64_t v0, uint64_t v1, double *__restrict__ out)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < v1; v += stride) out[v] = __longlong_as_double((long long)dist[v]);
}

// returns false when the peer mapping is unavailable (the caller then runs the replicated sweeps)
static bool sssp_multi_delta(gx_graph *g, const SsspCache &sc, uint64_t src, uint64_t &relaxed, uint32_t &rounds)
{
    Context &c = ctx();
    const uint64_t n = g->n, m = g->m;
    const double delta = sc.delta;
    const Partition &part = g->out.plan.part;
    const uint64_t v0 = part.lo, v1 = part.hi, own = v1 - v0;
    PeerBuf db;
    peer_alloc(db, n * sizeof(unsigned long long));
    if (!db.shared) { peer_free(db); return false; }
    struct Release { PeerBuf &b; ~Release() { peer_free(b); } } release{db};
    unsigned long long *dist = (unsigned long long *)db.local;
    DevBuf<unsigned long long> ldone(own ? own : 1), hdone(own ? own : 1);
    DevBuf<uint32_t> queue(own + 1024);
    const uint64_t big_cap = m / CHUNK + m / 128 + 16;
    DevBuf<uint32_t> big_row(big_cap);
    DevBuf<uint64_t> big_begin(big_cap), big_end(big_cap);
    DevBuf<SsspCounters> cnt(1);
    DevBuf<SsspRound> round(1);
    DevBuf<SsspPeers> peers(1);
    {
        SsspPeers h{};
        h.nranks = c.nranks;
        h.rank = c.rank;
        for (int r = 0; r < c.nranks; r++) { h.dist[r] = (unsigned long long *)db.peer[r]; h.bound[r] = part.b[r]; }
        h.bound[c.nranks] = part.b[c.nranks];
        GX_CUDA(cudaMemcpyAsync(peers.p, &h, sizeof(h), cudaMemcpyHostToDevice, c.stream));
        GX_CUDA(cudaStreamSynchronize(c.stream));
    }
    auto bits = [](double x) { unsigned long long b; memcpy(&b, &x, sizeof(b)); return b; };
    // the per-round exchange: peer mailboxes (GX_SSSP_MAIL=0: NCCL all-reduce + device-to-host copy)
    PeerMail no_mail;
    const char *me = getenv("GX_SSSP_MAIL");
    PeerMail &mail = (me && me[0] == '0') ? no_mail : context_mail();
    GX_LAUNCH(k_sssp_m_init, grid_persistent(8), 256, 0, dist, n, (uint32_t)src, ldone.p, hdone.p, own);
    cnt.zero();
    GX_CUDA(cudaMemsetAsync(&cnt.p->far_min, 0xFF, sizeof(unsigned long long), c.stream));
    round.zero();
    allreduce(round.p, 2, Dt::U64, Red::Max); // nobody may forward into a rank's array before that rank has initialised it
    double T = delta;
    bool expanded = false; // some rank light-expanded a vertex since the last heavy phase
    // one round: build the queue of due vertices, expand them, exchange {queue size, smallest waiting distance};
    // the all-reduce is also the barrier after which every forwarded improvement of the round has landed
    auto run_round = [&](bool heavy, SsspRound &h) {
        if (!mail.ok) {
            cnt.zero();
            GX_CUDA(cudaMemsetAsync(&cnt.p->far_min, 0xFF, sizeof(unsigned long long), c.stream));
        }
        if (heavy) {
            GX_LAUNCH(k_sssp_m_build<true>, grid_persistent(8), 256, 0, dist, hdone.p, v0, v1, bits(T), queue.p, cnt.p);
            GX_LAUNCH((k_sssp_expand<16, true>), grid_persistent(8), 256, 0, g->out.rowptr.p, g->out.col.p, g->out.w.p, queue.p,
                      &cnt.p->next_count, dist, nullptr, nullptr, 0u, big_row.p, big_begin.p, big_end.p, cnt.p, bits(T), delta, peers.p);
            GX_LAUNCH(k_sssp_relax_pieces, grid_persistent(8), 256, 0, g->out.col.p, g->out.w.p, big_row.p, big_begin.p, big_end.p, dist,
                      nullptr, cnt.p, bits(T), delta, peers.p);
        } else {
            GX_LAUNCH(k_sssp_m_build<false>, grid_persistent(8), 256, 0, dist, ldone.p, v0, v1, bits(T), queue.p, cnt.p);
            GX_LAUNCH((k_sssp_expand<8, false>), grid_persistent(8), 256, 0, sc.lrowptr.p, sc.lcol.p, sc.lw.p, queue.p, &cnt.p->next_count,
                      dist, nullptr, nullptr, 0u, big_row.p, big_begin.p, big_end.p, cnt.p, bits(T), -1.0, peers.p);
            GX_LAUNCH(k_sssp_relax_pieces, grid_persistent(8), 256, 0, sc.lcol.p, sc.lw.p, big_row.p, big_begin.p, big_end.p, dist, nullptr,
                      cnt.p, bits(T), -1.0, peers.p);
        }
        if (mail.ok) {
            const unsigned long long seq = ++mail.seq;
            GX_LAUNCH(k_sssp_m_exchange, 1, 32, 0, mail.table, cnt.p, seq, mail.host_dev);
            unsigned long long out[3];
            peer_mail_wait(mail, seq, out);
            h.qn = out[0]; h.neg_far_min = out[1]; h.relaxed = out[2];
        } else {
            GX_LAUNCH(k_sssp_m_pack, 1, 1, 0, cnt.p, round.p);
            allreduce(round.p, 2, Dt::U64, Red::Max);
            read_back(&h, round.p, sizeof(h));
        }
        relaxed += h.relaxed;
        rounds++;
    };
    for (;;) {
        SsspRound h;
        for (;;) { // light rounds until no rank has a due vertex below T
            run_round(false, h);
            if (!h.qn) break;
            expanded = true;
        }
        if (expanded) { // the bucket is stable: heavy entries of everything in it that moved, once
            SsspRound hh;
            run_round(true, hh);
            expanded = false;
            continue; // (re-scan: the heavy phase may have created due vertices, and the waiting minimum moved)
        }
        const unsigned long long fmin_bits = ~h.neg_far_min;
        if (fmin_bits == ~0ull) break; // nothing is waiting anywhere
        double fmin;
        memcpy(&fmin, &fmin_bits, sizeof(fmin));
        T += delta;
        if (fmin >= T) T = fmin + delta; // skip the empty buckets
    }
    GX_LAUNCH(k_sssp_m_out, grid_persistent(8), 256, 0, dist, v0, v1, g->res_f64.p);
    allgatherv(g->res_f64.p, Dt::F64, part);
    GX_CUDA(cudaStreamSynchronize(c.stream));
    return true;
}

} // namespace gx

using namespace gx;

void gx_sssp_cache_free(void *p) { delete (SsspCache *)p; }

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
        // bucket width: a few average edge weights per average degree (Davidson et al.'s near-far rule)
        double delta = 8.0 * (double)g->mean_weight * (double)n / (double)(m ? m : 1);
        if (const char *e = getenv("GX_SSSP_DELTA")) delta = atof(e); // tuning knob; 0 = plain sweeps
        const char *lh = getenv("GX_SSSP_LH");                        // GX_SSSP_LH=0: near/far without the light/heavy split
        const bool light_heavy = delta > 0.0 && m > 0 && !(lh && lh[0] == '0') && (!multi() || c.nranks <= MAX_PEERS);
        uint64_t relaxed = 0;
        uint32_t rounds = 0;
        bool done = false;
        if (light_heavy) {
            SsspCache *sc = (SsspCache *)g->sssp_cache;
            if (!sc || sc->delta != delta) {
                PhaseTimer tb(&c.timing.build_ms);
                if (sc) { delete sc; g->sssp_cache = nullptr; }
                g->sssp_cache = sc = build_sssp_cache(g, delta);
            }
            PhaseTimer tk(&c.timing.kernel_ms);
            if (multi()) done = sssp_multi_delta(g, *sc, src, relaxed, rounds);
            else { sssp_delta_stepping(g, *sc, src, relaxed, rounds); done = true; }
        }
        if (!done) {
        DevBuf<unsigned long long> dist(n);
        DevBuf<uint32_t> inq(n), q0(n), q1(n);
        DevBuf<unsigned long long> prev(multi() ? n : 0);
        const uint64_t big_cap = m / CHUNK + m / SSSP_BIG + 16;
        DevBuf<uint32_t> big_row(big_cap);
        DevBuf<uint64_t> big_begin(big_cap);
        DevBuf<SsspCounters> cnt(1);
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            GX_LAUNCH(k_sssp_init, grid_persistent(8), 256, 0, dist.p, n, (uint32_t)src, q0.p, inq.p);
            uint32_t *queue = q0.p, *next_q = q1.p;
            uint64_t qn = 1;
            if (multi()) GX_CUDA(cudaMemcpyAsync(prev.p, dist.p, n * sizeof(unsigned long long), cudaMemcpyDeviceToDevice, c.stream));
            uint32_t *inq_p = multi() ? nullptr : inq.p;
            // several GPUs run plain sweeps (threshold = +inf)
            const bool buckets = !multi() && delta > 0.0;
            double T = buckets ? delta : INFINITY;
            auto bits = [](double x) { unsigned long long b; memcpy(&b, &x, sizeof(b)); return b; };
            for (;;) {
                while (qn) {
                    cnt.zero();
                    // each rank relaxes the out-edges of the frontier vertices in its row block, on its replica
                    if (!multi()) GX_LAUNCH(k_sssp_clear, grid_for(qn, 256), 256, 0, queue, qn, inq.p);
                    GX_LAUNCH(k_sssp_relax, grid_for(qn * 32, 256) < grid_persistent(16) ? grid_for(qn * 32, 256) : grid_persistent(16), 256, 0, g->out.rowptr.p, g->out.col.p, g->out.w.p, queue, qn,
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
