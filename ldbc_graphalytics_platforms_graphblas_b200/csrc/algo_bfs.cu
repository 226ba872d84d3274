// algo_bfs.cu -- level-only BFS as push/pull structural SpMV with a direction
// switch.  Replaces LA_BFS (bfs.cpp:70-83) -> LAGr_BreadthFirstSearch(&level,
// NULL, G, src): level[src] = 0, no entry (GX_UNREACHED_LEVEL) when unreached.
// Levels are identical whichever direction a level is expanded in, so the
// switch is a pure performance decision (LAGraph's own push/pull rule, which
// the reference's wrapper disables by not caching AT, bfs.cpp:79-80).
//
//   push (sparse frontier queue, out-edges): warp per frontier vertex, hubs
//        re-queued as CHUNK-entry pieces for whole CTAs; atomicCAS claims
//        a vertex, warp-ballot compaction appends it to the next queue
//   pull (bitmap frontier, in-edges): thread per unvisited vertex, early exit
//        on the first parent in the frontier bitmap; one warp owns one
//        32-vertex bitmap word, written with a ballot (no atomics); rows longer
//        than ROW_SPLIT go to a warp-per-row kernel
// One GPU: the whole search is ONE cooperative launch (k_bfs_run): the level loop, the direction rule and the
// frontier bookkeeping run on the device with grid-wide barriers between the phases of a level, so a level costs a
// few barriers (~2 us each) instead of 3-4 launches, a memset and a device-to-host read of the counters.  Several
// GPUs keep the host loop (every pull level ends in an all-gather of the next frontier's bitmap words).
// Algorithmic bytes (one-pass bound): 4 m_reach + 8(n+1) + 4n + 2 (n/8) levels.
#include <cooperative_groups.h>

#include "graph.cuh"

namespace gx {

constexpr int32_t UNVIS = -1;
constexpr uint32_t PUSH_BIG = 4096; // frontier vertices with more out-edges are chunked

struct BfsCounters { unsigned long long nf, mf, next_count, big_count; };

__global__ void k_bfs_seed(int32_t *level, uint32_t *queue, uint32_t src)
{
    level[src] = 0;
    queue[0] = src;
}

// claim v for `depth`; returns true for the unique winner
__device__ __forceinline__ bool bfs_claim(int32_t *level, uint32_t v, int32_t depth)
{
    if (level[v] != UNVIS) return false;
    return atomicCAS(&level[v], UNVIS, depth) == UNVIS;
}

// warp-aggregated append of the winners of one 32-lane step
__device__ __forceinline__ void bfs_append(bool won, uint32_t v, uint32_t *next_q, BfsCounters *cnt)
{
    unsigned mask = __ballot_sync(FULL, won);
    if (mask == 0) return;
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(&cnt->next_count, (unsigned long long)__popc(mask));
    base = __shfl_sync(FULL, base, 0);
    if (won) next_q[base + __popc(mask & ((1u << lane_id()) - 1u))] = v;
}

// The same through a per-warp staging buffer in shared memory: the winners of several steps leave with one atomic on
// the queue's cursor.  In the big level of a push-only search nearly every 32-edge step has a winner, and several
// hundred thousand warp-steps queueing on one address cost more than the edges themselves.
constexpr unsigned BFS_STAGE = 128; // entries per warp; flushed when fewer than 32 slots are left
struct BfsStage {
    uint32_t *buf;
    unsigned n; // warp-uniform
};
__device__ __forceinline__ BfsStage bfs_stage()
{
    __shared__ uint32_t s_stage[32][BFS_STAGE]; // up to 32 warps per CTA
    return BfsStage{s_stage[threadIdx.x >> 5], 0u};
}
__device__ __forceinline__ void bfs_stage_flush(BfsStage &st, uint32_t *next_q, BfsCounters *cnt)
{
    if (st.n == 0) return;
    __syncwarp();
    unsigned long long base = 0;
    if (lane_id() == 0) base = atomicAdd(&cnt->next_count, (unsigned long long)st.n);
    base = __shfl_sync(FULL, base, 0);
    for (unsigned i = lane_id(); i < st.n; i += 32) next_q[base + i] = st.buf[i];
    __syncwarp();
    st.n = 0;
}
__device__ __forceinline__ void bfs_stage_append(BfsStage &st, bool won, uint32_t v, uint32_t *next_q, BfsCounters *cnt)
{
    const unsigned mask = __ballot_sync(FULL, won);
    if (mask == 0) return;
    if (won) st.buf[st.n + __popc(mask & ((1u << lane_id()) - 1u))] = v;
    st.n += __popc(mask);
    if (st.n > BFS_STAGE - 32) bfs_stage_flush(st, next_q, cnt);
}

__device__ __forceinline__ void
bfs_push_phase(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const uint32_t *__restrict__ queue,
               uint64_t qn, int32_t *__restrict__ level, int32_t depth, uint32_t *__restrict__ next_q,
               uint32_t *__restrict__ big_row, uint64_t *__restrict__ big_begin, BfsCounters *__restrict__ cnt)
{
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    unsigned long long nf = 0, mf = 0;
    BfsStage st = bfs_stage();
    for (; wid < qn; wid += nw) {
        const uint32_t u = queue[wid];
        const uint64_t a = rowptr[u], b = rowptr[u + 1];
        if (b - a > PUSH_BIG) {
            const uint64_t nch = (b - a + CHUNK - 1) / CHUNK;
            unsigned long long pos = 0;
            if (lane_id() == 0) pos = atomicAdd(&cnt->big_count, (unsigned long long)nch);
            pos = __shfl_sync(FULL, pos, 0);
            for (uint64_t k = lane_id(); k < nch; k += 32) { big_row[pos + k] = u; big_begin[pos + k] = a + k * CHUNK; }
            continue;
        }
        for (uint64_t base = a; base < b; base += 32) {
            const uint64_t e = base + lane_id();
            bool won = false;
            uint32_t v = 0;
            if (e < b) {
                v = ld_stream(col + e);
                won = bfs_claim(level, v, depth);
                if (won) { nf++; mf += rowptr[v + 1] - rowptr[v]; }
            }
            bfs_stage_append(st, won, v, next_q, cnt);
        }
    }
    bfs_stage_flush(st, next_q, cnt);
    nf = warp_sum(nf);
    mf = warp_sum(mf);
    if (lane_id() == 0 && nf) { atomicAdd(&cnt->nf, nf); atomicAdd(&cnt->mf, mf); }
}

__global__ void __launch_bounds__(256)
k_bfs_push(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const uint32_t *__restrict__ queue,
           uint64_t qn, int32_t *__restrict__ level, int32_t depth, uint32_t *__restrict__ next_q,
           uint32_t *__restrict__ big_row, uint64_t *__restrict__ big_begin, BfsCounters *__restrict__ cnt)
{
    bfs_push_phase(rowptr, col, queue, qn, level, depth, next_q, big_row, big_begin, cnt);
}

__device__ __forceinline__ void
bfs_push_big_phase(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const uint32_t *__restrict__ big_row,
                   const uint64_t *__restrict__ big_begin, int32_t *__restrict__ level, int32_t depth,
                   uint32_t *__restrict__ next_q, BfsCounters *__restrict__ cnt)
{
    const unsigned long long nbig = cnt->big_count;
    unsigned long long nf = 0, mf = 0;
    BfsStage st = bfs_stage();
    for (unsigned long long c = blockIdx.x; c < nbig; c += gridDim.x) {
        const uint64_t b0 = big_begin[c];
        const uint64_t row_end = rowptr[big_row[c] + 1];
        const uint64_t e_end = (b0 + CHUNK < row_end) ? b0 + CHUNK : row_end;
        for (uint64_t base = b0; base < e_end; base += blockDim.x) {
            const uint64_t e = base + threadIdx.x;
            bool won = false;
            uint32_t v = 0;
            if (e < e_end) {
                v = ld_stream(col + e);
                won = bfs_claim(level, v, depth);
                if (won) { nf++; mf += rowptr[v + 1] - rowptr[v]; }
            }
            bfs_stage_append(st, won, v, next_q, cnt);
        }
    }
    bfs_stage_flush(st, next_q, cnt);
    nf = warp_sum(nf);
    mf = warp_sum(mf);
    if (lane_id() == 0 && nf) { atomicAdd(&cnt->nf, nf); atomicAdd(&cnt->mf, mf); }
}

__global__ void __launch_bounds__(256)
k_bfs_push_big(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, const uint32_t *__restrict__ big_row,
               const uint64_t *__restrict__ big_begin, int32_t *__restrict__ level, int32_t depth,
               uint32_t *__restrict__ next_q, BfsCounters *__restrict__ cnt)
{
    bfs_push_big_phase(rowptr, col, big_row, big_begin, level, depth, next_q, cnt);
}

// frontier (level == cur) and visited (level != UNVIS) bitmaps from the level array
__device__ __forceinline__ void bfs_bitmaps_phase(const int32_t *__restrict__ level, uint64_t n, int32_t cur,
                                                  uint32_t *__restrict__ front, uint32_t *__restrict__ visited)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nround = (n + 31) & ~31ull;
    for (; v < nround; v += stride) {
        int32_t l = v < n ? level[v] : 0; // padding bits read as "visited", never in the frontier
        unsigned f = __ballot_sync(FULL, v < n && l == cur);
        unsigned s = __ballot_sync(FULL, l != UNVIS);
        if (lane_id() == 0) { front[v >> 5] = f; visited[v >> 5] = s; }
    }
}

__global__ void k_bfs_bitmaps(const int32_t *__restrict__ level, uint64_t n, int32_t cur, uint32_t *__restrict__ front,
                              uint32_t *__restrict__ visited)
{
    bfs_bitmaps_phase(level, n, cur, front, visited);
}

__device__ __forceinline__ void
bfs_pull_phase(const uint64_t *__restrict__ in_rowptr, const uint32_t *__restrict__ in_col,
               const uint64_t *__restrict__ out_rowptr, uint64_t n, uint64_t v0, uint64_t v1,
               const uint32_t *__restrict__ front, uint32_t *__restrict__ visited, uint32_t *__restrict__ next,
               int32_t *__restrict__ level, int32_t depth, BfsCounters *__restrict__ cnt)
{
    // [v0, v1) is this rank's row block; v0 is a multiple of 32, so a warp still owns whole words
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nround = (v1 + 31) & ~31ull;
    unsigned long long nf = 0, mf = 0, scanned = 0;
    for (; v < nround; v += stride) {
        const uint32_t vis = visited[v >> 5];
        bool found = false;
        if (vis != 0xFFFFFFFFu && !((vis >> lane_id()) & 1u) && v < n) {
            const uint64_t a = in_rowptr[v], b = in_rowptr[v + 1];
            if (b - a <= ROW_SPLIT) {
                // the first entry alone (in the dense levels it is a parent more often than not), then four entries and
                // their frontier words in flight per trip: the walk is a chain of dependent loads otherwise.
                // `scanned` counts up to the first parent, as a one-at-a-time walk would
                if (a < b) {
                    const uint32_t u0 = in_col[a];
                    scanned++;
                    found = (front[u0 >> 5] >> (u0 & 31u)) & 1u;
                }
                for (uint64_t e = a + 1; e < b && !found; e += 4) {
                    uint32_t u[4], fw[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) u[j] = e + j < b ? in_col[e + j] : 0xFFFFFFFFu;
#pragma unroll
                    for (int j = 0; j < 4; j++) fw[j] = u[j] != 0xFFFFFFFFu ? front[u[j] >> 5] : 0u;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if (found || u[j] == 0xFFFFFFFFu) continue;
                        scanned++;
                        if ((fw[j] >> (u[j] & 31u)) & 1u) found = true;
                    }
                }
            }
            if (found) { level[v] = depth; nf++; mf += out_rowptr[v + 1] - out_rowptr[v]; }
        }
        const unsigned fm = __ballot_sync(FULL, found);
        if (lane_id() == 0) { next[v >> 5] = fm; if (fm) visited[v >> 5] = vis | fm; }
    }
    nf = warp_sum(nf);
    mf = warp_sum(mf);
    scanned = warp_sum(scanned);
    if (lane_id() == 0 && (nf | scanned)) { atomicAdd(&cnt->nf, nf); atomicAdd(&cnt->mf, mf); atomicAdd(&cnt->next_count, scanned); }
}

__global__ void __launch_bounds__(256)
k_bfs_pull(const uint64_t *__restrict__ in_rowptr, const uint32_t *__restrict__ in_col,
           const uint64_t *__restrict__ out_rowptr, uint64_t n, uint64_t v0, uint64_t v1,
           const uint32_t *__restrict__ front, uint32_t *__restrict__ visited, uint32_t *__restrict__ next,
           int32_t *__restrict__ level, int32_t depth, BfsCounters *__restrict__ cnt)
{
    bfs_pull_phase(in_rowptr, in_col, out_rowptr, n, v0, v1, front, visited, next, level, depth, cnt);
}

// rows the thread-per-vertex kernel skipped: one warp per long row, ballot early exit
__device__ __forceinline__ void
bfs_pull_long_phase(const uint64_t *__restrict__ in_rowptr, const uint32_t *__restrict__ in_col,
                    const uint64_t *__restrict__ out_rowptr, const uint32_t *__restrict__ long_rows, uint64_t n_long,
                    const uint32_t *__restrict__ front, uint32_t *__restrict__ visited, uint32_t *__restrict__ next,
                    int32_t *__restrict__ level, int32_t depth, BfsCounters *__restrict__ cnt)
{
    uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (; wid < n_long; wid += nw) {
        const uint32_t v = long_rows[wid];
        if ((visited[v >> 5] >> (v & 31u)) & 1u) continue;
        const uint64_t a = in_rowptr[v], b = in_rowptr[v + 1];
        bool found = false;
        unsigned long long scanned = 0;
        for (uint64_t base = a; base < b && !found; base += 32) {
            const uint64_t e = base + lane_id();
            bool hit = false;
            if (e < b) { const uint32_t u = in_col[e]; hit = (front[u >> 5] >> (u & 31u)) & 1u; }
            found = __any_sync(FULL, hit);
            scanned += (b - base < 32) ? (b - base) : 32;
        }
        if (lane_id() == 0) {
            atomicAdd(&cnt->next_count, scanned);
            if (found) {
                level[v] = depth;
                atomicOr(&next[v >> 5], 1u << (v & 31u));
                atomicOr(&visited[v >> 5], 1u << (v & 31u));
                atomicAdd(&cnt->nf, 1ull);
                atomicAdd(&cnt->mf, (unsigned long long)(out_rowptr[v + 1] - out_rowptr[v]));
            }
        }
    }
}

__global__ void __launch_bounds__(256)
k_bfs_pull_long(const uint64_t *__restrict__ in_rowptr, const uint32_t *__restrict__ in_col,
                const uint64_t *__restrict__ out_rowptr, const uint32_t *__restrict__ long_rows, uint64_t n_long,
                const uint32_t *__restrict__ front, uint32_t *__restrict__ visited, uint32_t *__restrict__ next,
                int32_t *__restrict__ level, int32_t depth, BfsCounters *__restrict__ cnt)
{
    bfs_pull_long_phase(in_rowptr, in_col, out_rowptr, long_rows, n_long, front, visited, next, level, depth, cnt);
}

// Multi-GPU pull: after the owners' bitmap words were all-gathered, every rank applies the words
// it does not own to its replica of level/visited and recounts the new frontier (vertices and
// out-edges) over all words, so that all ranks take the same direction decision.
__global__ void k_bfs_merge(const uint32_t *__restrict__ next, uint32_t *__restrict__ visited, int32_t *__restrict__ level,
                            const uint64_t *__restrict__ out_rowptr, uint64_t n, uint64_t w0, uint64_t w1, int32_t depth,
                            BfsCounters *__restrict__ cnt)
{
    const uint64_t words = (n + 31) / 32;
    uint64_t w = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long nf = 0, mf = 0;
    for (; w < words; w += stride) {
        uint32_t bits = next[w];
        if (!bits) continue;
        const bool remote = w < w0 || w >= w1;
        if (remote) visited[w] |= bits;
        while (bits) {
            const uint32_t v = (uint32_t)(w * 32) + (uint32_t)(__ffs(bits) - 1);
            bits &= bits - 1;
            if (remote) level[v] = depth;
            nf++;
            mf += out_rowptr[v + 1] - out_rowptr[v];
        }
    }
    nf = warp_sum(nf);
    mf = warp_sum(mf);
    if (lane_id() == 0 && nf) { atomicAdd(&cnt->nf, nf); atomicAdd(&cnt->mf, mf); }
}

// queue of the vertices with level == cur (pull -> push switch)
__global__ void k_bfs_level_to_queue(const int32_t *__restrict__ level, uint64_t n, int32_t cur,
                                     uint32_t *__restrict__ queue, BfsCounters *__restrict__ cnt)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t nround = (n + 31) & ~31ull;
    for (; v < nround; v += stride) {
        bool in = v < n && level[v] == cur;
        bfs_append(in, (uint32_t)v, queue, cnt);
    }
}

__global__ void k_bfs_widen(const int32_t *__restrict__ level, uint64_t n, int64_t *__restrict__ out)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) { int32_t l = level[v]; out[v] = l == UNVIS ? GX_UNREACHED_LEVEL : (int64_t)l; }
}


// ----------------------------------------------------------------------------- one GPU: the whole search in one launch
struct BfsRunArgs {
    const uint64_t *out_rowptr; const uint32_t *out_col;
    const uint64_t *in_rowptr; const uint32_t *in_col;
    const uint32_t *long_rows; uint64_t n_long;
    uint64_t n, m; uint32_t src; int can_pull;
    int32_t *level; uint32_t *q0, *q1, *bm_front, *bm_next, *bm_vis;
    uint32_t *big_row; uint64_t *big_begin;
    BfsCounters *cnt;          // 3 slots used in turn (level d uses slot d % 3, zeroes slot (d + 1) % 3)
    unsigned long long *qcur;  // append cursor of the pull -> push conversion
    int64_t *out;              // widened levels
    unsigned long long *stats; // levels, edges inspected, m_reach
};

// Same level loop, direction rule and phases as the host loop of gx_bfs below (several GPUs); every thread keeps the
// loop state in registers -- it is a function of the counters all threads read after the same barrier.
// CTAS: co-resident CTAs per SM the launch bound asks for (4 / 6 / 8 -> 64 / 40 / 32 registers; the loop state spills
// at 6 and 8, but the pull phase is a chain of dependent loads and wants the threads)
template <int THREADS, int CTAS>
__global__ void __launch_bounds__(THREADS, CTAS) k_bfs_run(const BfsRunArgs a)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const uint64_t gtid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, gth = (uint64_t)gridDim.x * blockDim.x;
    const uint64_t n = a.n;
    for (uint64_t v = gtid; v < n; v += gth) a.level[v] = UNVIS;
    if (gtid < 3) a.cnt[gtid] = BfsCounters{0, 0, 0, 0};
    if (gtid == 0) *a.qcur = 0;
    grid.sync();
    if (gtid == 0) { a.level[a.src] = 0; a.q0[0] = a.src; }
    uint64_t nf = 1, mf = a.out_rowptr[a.src + 1] - a.out_rowptr[a.src], m_unvisited = a.m, prev_nf = 0;
    uint64_t m_reach = mf, inspected = 0;
    bool pull = false, have_queue = true, have_bitmaps = false;
    uint32_t *queue = a.q0, *next_q = a.q1, *front = a.bm_front, *next = a.bm_next;
    int32_t depth = 0;
    uint32_t levels = 0;
    grid.sync();
    while (nf > 0) {
        depth++;
        m_unvisited = m_unvisited > mf ? m_unvisited - mf : 0;
        if (!pull) { if (a.can_pull && mf > m_unvisited / 8 && nf > prev_nf && nf > 1) pull = true; }
        else if (nf < n / 500 + 1 && nf < prev_nf) pull = false;
        prev_nf = nf;
        BfsCounters *cnt = a.cnt + depth % 3;
        if (gtid == 0) a.cnt[(depth + 1) % 3] = BfsCounters{0, 0, 0, 0}; // last read two barriers ago
        if (!pull) {
            if (!have_queue) {
                const uint64_t nround = (n + 31) & ~31ull;
                for (uint64_t v = gtid; v < nround; v += gth) {
                    const bool in = v < n && a.level[v] == depth - 1;
                    const unsigned mask = __ballot_sync(FULL, in);
                    if (mask) {
                        unsigned long long base = 0;
                        if (lane_id() == 0) base = atomicAdd(a.qcur, (unsigned long long)__popc(mask));
                        base = __shfl_sync(FULL, base, 0);
                        if (in) queue[base + __popc(mask & ((1u << lane_id()) - 1u))] = (uint32_t)v;
                    }
                }
                grid.sync();
                if (gtid == 0) *a.qcur = 0;
            }
            bfs_push_phase(a.out_rowptr, a.out_col, queue, nf, a.level, depth, next_q, a.big_row, a.big_begin, cnt);
            grid.sync();
            if (cnt->big_count) {
                bfs_push_big_phase(a.out_rowptr, a.out_col, a.big_row, a.big_begin, a.level, depth, next_q, cnt);
            }
            uint32_t *t = queue; queue = next_q; next_q = t;
            have_queue = true;
            have_bitmaps = false;
            inspected += mf;
        } else {
            if (!have_bitmaps) { bfs_bitmaps_phase(a.level, n, depth - 1, front, a.bm_vis); grid.sync(); }
            bfs_pull_phase(a.in_rowptr, a.in_col, a.out_rowptr, n, 0, n, front, a.bm_vis, next, a.level, depth, cnt);
            if (a.n_long) {
                grid.sync(); // the long rows OR their bits into words the pass above stored whole
                bfs_pull_long_phase(a.in_rowptr, a.in_col, a.out_rowptr, a.long_rows, a.n_long, front, a.bm_vis, next, a.level,
                                    depth, cnt);
            }
            uint32_t *t = front; front = next; next = t;
            have_bitmaps = true;
            have_queue = false;
        }
        grid.sync();
        const volatile BfsCounters *vc = cnt;
        if (pull) inspected += vc->next_count;
        nf = vc->nf;
        mf = vc->mf;
        m_reach += mf;
        levels++;
    }
    for (uint64_t v = gtid; v < n; v += gth) { const int32_t l = a.level[v]; a.out[v] = l == UNVIS ? GX_UNREACHED_LEVEL : (int64_t)l; }
    if (gtid == 0) { a.stats[0] = levels; a.stats[1] = inspected; a.stats[2] = m_reach; }
}

} // namespace gx

using namespace gx;

extern "C" int gx_bfs(gx_graph *g, uint64_t src, int64_t *level_host)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(g != nullptr, "graph is NULL");
        GX_REQUIRE(src < g->n, "source vertex out of range");
        Context &c = ctx();
        c.timing = gx_timing{};
        const uint64_t n = g->n, m = g->m;
        // LAGraph's rule (LG_BreadthFirstSearch_SSGrB): pull steps only when the transposed adjacency is
        // already cached -- a single BFS never pays for a transposition (bfs.cpp:79-80 caches nothing, so the
        // reference runs push-only).  Undirected graphs pull on the one adjacency they have.
        const bool can_pull = !g->directed || g->have_in;
        Adj &in = g->in_adj();
        if (can_pull) {
            PhaseTimer tb(&c.timing.build_ms);
            ensure_plan(in, n);
        }
        const uint64_t words = (n + 31) / 32;
        g->res_i64.alloc(n);
        DevBuf<int32_t> level(n);
        DevBuf<uint32_t> q0(n), q1(n), bm_front(words), bm_next(words), bm_vis(words);
        const uint64_t big_cap = m / CHUNK + m / PUSH_BIG + 16;
        DevBuf<uint32_t> big_row(big_cap);
        DevBuf<uint64_t> big_begin(big_cap);
        DevBuf<BfsCounters> cnt(1);
        uint64_t m_reach = 0, inspected = 0;
        uint32_t levels = 0;
        // GX_BFS_COOP: 0 = host loop; otherwise the shape of the one-launch search, threads per CTA * 10 + CTAs per SM
        // (fewer, larger CTAs make the grid barrier cheaper: it costs one atomic per CTA)
        const char *ce = getenv("GX_BFS_COOP");
        // default: one launch up to 2^23 vertices (RMAT-22: 0.32 -> 0.21 ms); beyond, the pull levels are long enough
        // that the host loop's full occupancy (2048 threads per SM at 32 registers) wins over its round trips
        int shape = ce ? atoi(ce) : (n <= (1ull << 23) ? 5122 : 0);
        struct RunShape { int code, threads, ctas; const void *fn; };
        static const RunShape shapes[] = {{2564, 256, 4, (const void *)k_bfs_run<256, 4>}, {2568, 256, 8, (const void *)k_bfs_run<256, 8>},
                                          {5122, 512, 2, (const void *)k_bfs_run<512, 2>}, {5124, 512, 4, (const void *)k_bfs_run<512, 4>},
                                          {10241, 1024, 1, (const void *)k_bfs_run<1024, 1>}, {10242, 1024, 2, (const void *)k_bfs_run<1024, 2>}};
        static int coop_ok[6] = {0, 0, 0, 0, 0, 0}; // 0 unknown, 1 co-resident at that shape, -1 not
        const RunShape *rs = nullptr;
        int want_ctas = 0;
        if (!multi() && shape > 0) {
            int si = 4;
            for (int i = 0; i < 6; i++) if (shapes[i].code == shape) si = i;
            if (!coop_ok[si]) {
                int dev = 0, can = 0, occ = 0;
                GX_CUDA(cudaGetDevice(&dev));
                GX_CUDA(cudaDeviceGetAttribute(&can, cudaDevAttrCooperativeLaunch, dev));
                GX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, shapes[si].fn, shapes[si].threads, 0));
                coop_ok[si] = (can && occ >= shapes[si].ctas) ? 1 : -1;
            }
            if (coop_ok[si] > 0) { rs = &shapes[si]; want_ctas = rs->ctas; }
        }
        if (!multi() && want_ctas > 0) {
            DevBuf<BfsCounters> cnt3(3);
            DevBuf<unsigned long long> aux(4); // qcur, stats[3]
            BfsRunArgs a;
            a.out_rowptr = g->out.rowptr.p; a.out_col = g->out.col.p;
            a.in_rowptr = can_pull ? in.rowptr.p : nullptr; a.in_col = can_pull ? in.col.p : nullptr;
            a.long_rows = can_pull ? in.plan.long_rows.p : nullptr; a.n_long = can_pull ? in.plan.n_long : 0;
            a.n = n; a.m = m; a.src = (uint32_t)src; a.can_pull = can_pull ? 1 : 0;
            a.level = level.p; a.q0 = q0.p; a.q1 = q1.p; a.bm_front = bm_front.p; a.bm_next = bm_next.p; a.bm_vis = bm_vis.p;
            a.big_row = big_row.p; a.big_begin = big_begin.p; a.cnt = cnt3.p; a.qcur = aux.p; a.out = g->res_i64.p;
            a.stats = aux.p + 1;
            unsigned long long st[3] = {0, 0, 0};
            {
                PhaseTimer tk(&c.timing.kernel_ms);
                void *params[] = {(void *)&a};
                const bool prof__ = profiling();
                if (prof__) prof_begin("k_bfs_run");
                GX_CUDA(cudaLaunchCooperativeKernel(rs->fn, dim3((unsigned)c.num_sms * (unsigned)want_ctas), dim3((unsigned)rs->threads),
                                                    params, 0, c.stream));
                if (prof__) prof_end();
                count_launch();
                GX_CUDA(cudaMemcpyAsync(c.pinned_scratch, aux.p + 1, sizeof(st), cudaMemcpyDeviceToHost, c.stream));
            }
            memcpy(st, c.pinned_scratch, sizeof(st)); // the timer's stop synchronised the stream
            levels = (uint32_t)st[0]; inspected = st[1]; m_reach = st[2];
        } else {
            PhaseTimer tk(&c.timing.kernel_ms);
            level.fill_byte(0xFF);
            GX_LAUNCH(k_bfs_seed, 1, 1, 0, level.p, q0.p, (uint32_t)src);
            uint64_t deg_src = 0;
            {
                uint64_t rp[2];
                read_back(rp, g->out.rowptr.p + src, sizeof(rp));
                deg_src = rp[1] - rp[0];
            }
            uint64_t nf = 1, mf = deg_src, m_unvisited = m, prev_nf = 0;
            m_reach = deg_src;
            bool pull = false, have_queue = true, have_bitmaps = false;
            uint32_t *queue = q0.p, *next_q = q1.p;
            uint32_t *front = bm_front.p, *next = bm_next.p;
            int32_t depth = 0;
            while (nf > 0) {
                depth++;
                m_unvisited = m_unvisited > mf ? m_unvisited - mf : 0;
                // direction rule (Beamer; LAGraph uses alpha = 8, beta = 500 on the same quantities)
                if (!pull) { if (can_pull && mf > m_unvisited / 8 && nf > prev_nf && nf > 1) pull = true; }
                else if (nf < n / 500 + 1 && nf < prev_nf) pull = false;
                prev_nf = nf;
                cnt.zero();
                if (!pull) {
                    if (!have_queue) {
                        GX_LAUNCH(k_bfs_level_to_queue, grid_persistent(8), 256, 0, level.p, n, depth - 1, queue, cnt.p);
                        cnt.zero(); // next_count was used as the append cursor
                    }
                    GX_LAUNCH(k_bfs_push, grid_for(nf * 32, 256), 256, 0, g->out.rowptr.p, g->out.col.p, queue, nf, level.p,
                              depth, next_q, big_row.p, big_begin.p, cnt.p);
                    GX_LAUNCH(k_bfs_push_big, grid_persistent(4), 256, 0, g->out.rowptr.p, g->out.col.p, big_row.p,
                              big_begin.p, level.p, depth, next_q, cnt.p);
                    uint32_t *t = queue; queue = next_q; next_q = t;
                    have_queue = true;
                    have_bitmaps = false;
                    inspected += mf;
                } else {
                    if (!have_bitmaps) GX_LAUNCH(k_bfs_bitmaps, grid_persistent(8), 256, 0, level.p, n, depth - 1, front, bm_vis.p);
                    // pull levels are split by row block; push levels (tiny frontiers) run replicated
                    const Partition &part = in.plan.part;
                    // several GPUs: a rank writes only the words of its row block; with the others zeroed the exchange is
                    // one all-reduce(max) of n/8 bytes (a group of per-owner broadcasts took 180 us per level on 8 GPUs)
                    if (multi()) GX_CUDA(cudaMemsetAsync(next, 0, words * sizeof(uint32_t), c.stream));
                    GX_LAUNCH(k_bfs_pull, grid_persistent(8), 256, 0, in.rowptr.p, in.col.p, g->out.rowptr.p, n, part.lo, part.hi,
                              front, bm_vis.p, next, level.p, depth, cnt.p);
                    if (in.plan.n_long)
                        GX_LAUNCH(k_bfs_pull_long, grid_for(in.plan.n_long * 32, 256), 256, 0, in.rowptr.p, in.col.p,
                                  g->out.rowptr.p, in.plan.long_rows.p, in.plan.n_long, front, bm_vis.p, next, level.p, depth,
                                  cnt.p);
                    if (multi()) {
                        allreduce(next, words, Dt::U32, Red::Max);
                        cnt.zero();
                        GX_LAUNCH(k_bfs_merge, grid_persistent(4), 256, 0, next, bm_vis.p, level.p, g->out.rowptr.p, n,
                                  part.lo / 32, part.hi == n ? words : part.hi / 32, depth, cnt.p);
                    }
                    uint32_t *t = front; front = next; next = t;
                    have_bitmaps = true;
                    have_queue = false;
                }
                BfsCounters h;
                read_back(&h, cnt.p, sizeof(h));
                if (pull) inspected += h.next_count;
                nf = h.nf;
                mf = h.mf;
                m_reach += mf;
                levels++;
            }
            GX_LAUNCH(k_bfs_widen, grid_persistent(8), 256, 0, level.p, n, g->res_i64.p);
        }
        c.timing.iterations = levels;
        c.timing.edges_inspected = inspected;
        c.timing.algorithmic_bytes = 4 * m_reach + 8 * (n + 1) + 4 * n + 2 * (n / 8) * (uint64_t)levels;
        if (level_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(level_host, g->res_i64.p, n * sizeof(int64_t), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
