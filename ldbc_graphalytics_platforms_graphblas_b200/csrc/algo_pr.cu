// algo_pr.cu -- PageRank as a pull plus.second SpMV over the in-edge adjacency
// with dangling-node correction.  Replaces LA_PR (pr.cpp:47-66) ->
// LAGr_PageRankGX(&r, &iters, G, (float)damping, itermax): exactly `iters`
// iterations, FP64, r0 = 1/n, d = outdeg / damping,
//   teleport' = (1-damping)/n + damping/n * sum_{sinks} r,   w = r ./ d,
//   r = teleport' + A' (plus.second) w.
//
// What bounds it on B200: the 4-byte column ids stream at 5.8 TB/s, but every entry also gathers one
// 8-byte w[source] at a random address, and a load whose 32 lanes hit 32 different cache lines
// occupies the SM's LSU pipe for 32 cycles (profiles/r1_microbench_stream_patterns.txt: 284 G
// gathers/s from an L2-resident vector = one per cycle per SM).  With the hottest sources in shared
// memory (0.27 cycles per gather) a 256-entry tile costs ~195 LSU cycles: >= 175 us per iteration on
// RMAT-22; the kernel runs at 233 us (DESIGN.md 7 has the experiments that led here).
//
// Shape of the kernel.  The entries of the rank's non-empty rows are cut into 256-entry tiles
// regardless of row borders (RMAT hubs span hundreds of tiles, typical rows a dozen entries); the
// last tile is padded with the id of a slot that always holds 0, so there are no bounds checks.
// A warp takes a tile; lane L owns the 8 CONSECUTIVE entries [8L, 8L+8): one 256-bit evict-first
// column load and then 8 independent gathers per lane are in flight before anything is consumed
// (the group-per-row shape of the first version kept ~1 in flight and was 27 % slower).  Row
// borders inside a tile are a precomputed 256-bit mask (one bit per entry that starts a row).
// A lane closes the rows that start and end among its 8 entries by itself; rows crossing lanes
// are closed by a warp segmented scan of the lanes' open sums (5 shuffle steps per 256
// entries); rows crossing tiles leave a head / tail partial per tile that k_pr_tile_fin adds in
// tile order (hub rows: by the warp).  Rows without entries (r = teleport') are done by the warps
// once they run out of tiles.  The epilogue multiplies by a precomputed 1/d (an FP64 division is
// ~30 instructions per closed row).  Every sum has a fixed order and the work lists are sorted, so
// results are bit-reproducible.  w lives in an index space of per-rank segments, each sorted by
// out-degree; the hottest slots of every segment -- 24 K in total, the source of 47 % of RMAT-22's
// entries -- are staged in shared memory by every CTA and carry the ids [0, hot) in the stored
// adjacency copy, all other slots are stored as slot + hot (one compare in the kernel).
//
// Kernels per iteration:
//   k_pr_tiles     persistent, one 1024-thread CTA per SM: teleport' from the previous sink
//                  partials, hot stage, warp-per-tile gather + segmented sums + fused epilogue
//                  r -> w' = r * (1/d) and sink mass, rows without entries in the tail
//   k_pr_tile_fin  thread per tile-crossing row (sum of its partials)
//   several GPUs:  k_pr_tele + all-reduce of the sink mass (also the barrier); w' reaches the other ranks
//                  by peer stores over NVLink -- from the two kernels themselves on 2 GPUs, by one
//                  copy kernel per iteration (k_pr_push_segment) on more
// Algorithmic bytes per iteration: 4m + 8(n+1) + 28n.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/transform_iterator.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "graph.cuh"

namespace gx {


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
// block-wide sum of `count` doubles in a fixed order (every thread gets the result);
// s_red needs (blockDim.x / 32) + 1 slots
__device__ __forceinline__ double block_sum_ordered(const double *__restrict__ parts, unsigned count, double *s_red)
{
    const unsigned nw = blockDim.x >> 5;
    double s = 0.0;
    for (unsigned i = threadIdx.x; i < count; i += blockDim.x) s += parts[i];
    s = warp_sum(s);
    if (lane_id() == 0) s_red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x < 32) {
        double x = threadIdx.x < nw ? s_red[threadIdx.x] : 0.0;
        x = warp_sum(x);
        if (threadIdx.x == 0) s_red[nw] = x;
    }
    __syncthreads();
    return s_red[nw];
}

// several GPUs only: this rank's sink partials folded into one scalar (all-reduced afterwards)
__global__ void k_pr_tele(const double *__restrict__ sink_part, unsigned nparts, double *__restrict__ sink_sum)
{
    __shared__ double s_red[33];
    const double x = block_sum_ordered(sink_part, nparts, s_red);
    if (threadIdx.x == 0) *sink_sum = x;
}

// the same with the all-reduce inside: the partial goes through the peer mailboxes (comm.cuh), ~45 us of NCCL launch
// + protocol per iteration on 8 GPUs become one NVLink round trip, and the host queues the whole loop without waiting
__global__ void k_pr_tele_mail(const double *__restrict__ sink_part, unsigned nparts, double *__restrict__ sink_sum, MailTable t,
                               unsigned long long seq)
{
    __shared__ double s_red[33];
    const double x = block_sum_ordered(sink_part, nparts, s_red);
    if (threadIdx.x < 32) {
        const double s = peer_mail_sum_f64(t, x, seq);
        if (threadIdx.x == 0) *sink_sum = s;
    }
}

struct PrScalars { double teleport, damping, n; };

// Where a new w' value goes: this rank's copy and, on several GPUs, every peer's copy of the vector
// (stores over NVLink peer mappings).  The exchange of w' is thereby part of the kernels that
// compute it; the all-reduce of the sink mass that opens the next iteration is the only barrier:
// it completes on a rank after every rank finished these kernels, i.e. after all their peer
// stores, and a rank reuses a buffer only two iterations later.
struct WOut {
    double *peer[MAX_PEERS - 1];  // the other ranks' copies
    int npeer;                    // 0 on one GPU
};

// k_pr_tiles takes the destinations by value (kernel parameter space, no dependent load before the stores)
struct WOutV {
    double *p[MAX_PEERS];
    int n;
};

template <bool PEERS>
__device__ __forceinline__ void w_store_v(const WOutV &o, uint32_t slot, double v)
{
    if constexpr (!PEERS) {
        o.p[0][slot] = v; // one GPU: no predicated-off stores in the instruction stream
    } else {
#pragma unroll
        for (int r = 0; r < MAX_PEERS; r++)
            if (r < o.n) o.p[r][slot] = v;
    }
}

template <bool PEERS>
__device__ __forceinline__ void w_store(double *__restrict__ own, const WOut *__restrict__ o, uint32_t slot, double v)
{
    own[slot] = v;
    if (PEERS)
        for (int r = 0; r < o->npeer; r++) o->peer[r][slot] = v;
}

// All-gather of w' by peer stores: this rank's finished segment is copied into every peer's buffer with
// 16-byte loads and stores over NVLink (one launch, no protocol, full-width stores -- unlike the 8-byte
// scattered stores of the fused path).  Peer p is served by the CTAs with blockIdx.x % npeer == p, so all
// links are busy at once.  The all-reduce of the sink mass that opens the next iteration is the barrier.
// A rank's segment has two pieces (often-gathered slots | the rest, see PrTiles); both counts are multiples of 32.
__global__ void __launch_bounds__(256)
k_pr_push_segment(const double *__restrict__ own, const WOut *__restrict__ o, uint64_t first, uint64_t count,
                  uint64_t first2, uint64_t count2)
{
    const int npeer = o->npeer;
    if (npeer == 0) return;
    const int p = blockIdx.x % npeer;
    const unsigned rank_in_peer = blockIdx.x / npeer, ctas_per_peer = gridDim.x / npeer; // gridDim.x is a multiple of npeer
    const uint64_t n2 = count / 2, m2 = count2 / 2;
    const double2 *src = (const double2 *)(own + first); // pieces start at multiples of 32 slots: 16-byte aligned
    double2 *dst = (double2 *)(o->peer[p] + first);
    const double2 *srcb = (const double2 *)(own + first2);
    double2 *dstb = (double2 *)(o->peer[p] + first2);
    for (uint64_t i = (uint64_t)rank_in_peer * 256 + threadIdx.x; i < n2 + m2; i += (uint64_t)ctas_per_peer * 256) {
        if (i < n2) dst[i] = src[i]; else dstb[i - n2] = srcb[i - n2];
    }
}

__global__ void k_fill_f64(double *__restrict__ p, uint64_t n, double v)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) p[i] = v;
}


constexpr int PT_TILE = 256;
constexpr int PT_WARPS = 32;          // one 1024-thread CTA per SM (64 registers per thread)
constexpr uint32_t PT_HOT = 24576;    // hottest entries of w staged in shared memory (192 KB)
constexpr uint64_t PT_WARM_MB = 48;   // often-gathered slots kept in L2 when w does not fit (all ranks' warm pieces together)
constexpr uint64_t PT_HINT_MIN_MB = 56; // w smaller than this stays in L2 by itself: one load policy

struct PrTiles {
    uint64_t K = 0, M = 0, n_tiles = 0, n_span = 0, n_empty = 0; // non-empty rows, entries, ...
    DevBuf<uint32_t> ne_rows;   // K: vertex of the k-th non-empty row of the block
    DevBuf<uint64_t> ne_ptr;    // K+1: local entry offset of its first entry (ne_ptr[0] = 0, ne_ptr[K] = M)
    DevBuf<uint32_t> pi;        // n: vertex -> slot in the index space w lives in (see k_pt_make_pi)
    uint64_t seg = 0, slots = 0; // slots per rank segment (equal, padded), nranks * seg
    // A segment is stored in two pieces so that the often-gathered slots of ALL ranks are one address range:
    // [rank 0 | rank 1 | ...] x wps slots (each rank's highest out-degrees), then [rank 0 | rank 1 | ...] x tps slots
    // (the rest); seg = wps + tps.  Gathers below `warm_end` ask L2 to keep their lines (evict_last), the tail is
    // read evict_first: on RMAT-25 w is 136 MB -- larger than L2 -- and its first third takes 90 % of the gathers.
    uint64_t wps = 0, tps = 0;
    bool hint = false;           // the two-policy gather is compiled in (w larger than about half of L2)
    // Several GPUs, compact form (3+ ranks): a rank keeps w only for the slots it gathers -- its own segment plus, per
    // other owner, the slots some row of its block refers to, in slot order: local ids [lbase[o], lbase[o+1]) belong to
    // owner o.  On RMAT-25 / 8 ranks that is ~40 % of the vector: it stays in L2, and the per-iteration exchange moves
    // only what is read.  An owner knows what a consumer needs without being told: consumer c needs source v iff v has
    // an out-entry into c's row block (the out-adjacency is replicated), so both sides derive the same ascending lists.
    bool compact = false;
    uint64_t L = 0, Lmax = 0;             // local slots of this rank / the largest over the ranks (equal buffer sizes)
    uint64_t lbase[MAX_PEERS + 1] = {0};
    DevBuf<uint32_t> push_list;           // per consumer: offsets inside the own segment, ascending, concatenated
    uint64_t push_off[MAX_PEERS + 1] = {0};
    uint64_t push_dst[MAX_PEERS] = {0};   // where this rank's range starts in consumer c's local space
    uint32_t hps = 0, hot = 0;   // hot entries per segment / in total (hps * nranks): the stored ids are h < hot for the
                                 // hottest hps slots of every segment and slot + hot for all others
    PeerBuf wbuf[2];             // the two copies of w (read / written in turn), mapped into all ranks
    bool have_wbuf = false;
    ~PrTiles() { if (have_wbuf) { peer_free(wbuf[0]); peer_free(wbuf[1]); } }
    DevBuf<uint32_t> col;       // M: pi(source) of this rank's slice of the in-edges (16-byte aligned tiles)
    DevBuf<uint32_t> tile_k0;   // n_tiles: non-empty row holding the tile's first entry; bit 31: that row starts there
    DevBuf<uint32_t> mask;      // M bits (32 bytes per tile): entry starts a row (tile-first entries excluded)
    DevBuf<uint32_t> slot_k;    // K: pi(vertex of row k), where its w' goes
    DevBuf<uint32_t> span_k;    // n_span: non-empty rows lying in more than one tile
    // operands of k_pr_tile_fin, one entry per thread so that its loads are independent:
    DevBuf<uint32_t> fin_v, fin_slot, fin_t0, fin_nt; // n_span + n_empty: vertex, slot of w', first tile, following tiles
    DevBuf<uint32_t> empty_rows; // n_empty
};

// row lists of [v0, v1): flag = 1 appends non-empty rows (in order, via their rank among
// non-empty rows), flag = 0 appends empty rows (any order)
__global__ void k_pt_mark(const uint64_t *__restrict__ rowptr, uint64_t v0, uint64_t v1, uint32_t *__restrict__ nonempty)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < v1; v += stride) nonempty[v - v0] = rowptr[v + 1] > rowptr[v] ? 1u : 0u;
}

__global__ void k_pt_fill(const uint64_t *__restrict__ rowptr, uint64_t v0, uint64_t v1, const uint32_t *__restrict__ rank_ne,
                          uint64_t e0, uint32_t *__restrict__ ne_rows, uint64_t *__restrict__ ne_ptr,
                          uint32_t *__restrict__ empty_rows)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < v1; v += stride) {
        const uint64_t b = rowptr[v], e = rowptr[v + 1];
        const uint32_t k = rank_ne[v - v0]; // number of non-empty rows before v
        if (e > b) { ne_rows[k] = (uint32_t)v; ne_ptr[k] = b - e0; }
        else empty_rows[(v - v0) - k] = (uint32_t)v; // rank among empty rows
    }
}

// Sort key of the index space w lives in: owner rank first (each rank's rows then occupy one
// contiguous segment, so the all-gather needs no re-ordering), inside a segment by descending
// out-degree (the often-gathered entries share cache lines), ties by vertex id.
// (the sort is stable and the vertices enter it in id order, so the id need not be part of the key: a
// 32-bit key, 24 + log2(ranks) significant bits, with the vertex as payload)
__global__ void k_pt_degree_keys(const uint64_t *__restrict__ out_rowptr, uint64_t n, const uint64_t *__restrict__ bounds,
                                 int nranks, uint32_t *__restrict__ keys, uint32_t *__restrict__ vertex)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) {
        uint32_t owner = 0;
        while ((int)owner + 1 < nranks && v >= bounds[owner + 1]) owner++;
        const uint64_t od = out_rowptr[v + 1] - out_rowptr[v];
        const uint32_t inv = 0xFFFFFFu - (uint32_t)(od > 0xFFFFFFull ? 0xFFFFFFull : od);
        keys[v] = (owner << 24) | inv;
        vertex[v] = (uint32_t)v;
    }
}

// slot of the i-th vertex in sorted order: rank r's vertices fill the first (b[r+1] - b[r]) slots of the
// segment [r * seg, (r+1) * seg) -- equal, padded segments, so the exchange is one plain all-gather
__global__ void k_pt_make_pi(const uint32_t *__restrict__ sorted_keys, const uint32_t *__restrict__ sorted_vertex, uint64_t n,
                             const uint64_t *__restrict__ bounds, uint64_t wps, uint64_t tps, uint64_t nranks,
                             uint32_t *__restrict__ pi)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n; i += stride) {
        const uint64_t owner = sorted_keys[i] >> 24, off = i - bounds[owner];
        pi[sorted_vertex[i]] = (uint32_t)(off < wps ? owner * wps + off : nranks * wps + owner * tps + (off - wps));
    }
}

// Stored id of an entry: the hottest `hps` slots of EVERY rank's segment (each segment is sorted by
// out-degree) get the ids [0, hot) of the shared-memory stage, every other slot is stored as slot + hot,
// so the kernel's test stays one compare.  Entries beyond `count` pad the last tile: they gather the
// always-zero slot `zero_slot`.
__global__ void k_pt_relabel_slice(const uint32_t *__restrict__ col, const uint32_t *__restrict__ pi, uint64_t count,
                                   uint64_t padded, uint32_t zero_slot, uint32_t wps, uint32_t warm_slots, uint32_t hps,
                                   uint32_t hot, uint32_t *__restrict__ out)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < padded; e += stride) {
        uint32_t slot = zero_slot;
        if (e < count) slot = pi[col[e]];
        uint32_t id = slot + hot;
        if (slot < warm_slots) { // the hot stage holds the first hps slots of every rank's warm piece
            const uint32_t owner = slot < wps ? 0u : slot / wps, off = slot - owner * wps; // (one GPU: no division)
            if (off < hps) id = owner * hps + off;
        }
        out[e] = id;
    }
}

// w0 in the degree-sorted space (replicated on every rank; compact form: only the slots this rank keeps)
__global__ void k_pt_scatter(const double *__restrict__ w_nat, const uint32_t *__restrict__ pi, uint64_t n,
                             double *__restrict__ w_perm)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) {
        const uint32_t s = pi[v];
        if (s != 0xFFFFFFFFu) w_perm[s] = w_nat[v];
    }
}

// ---- compact form ---------------------------------------------------------------------------------------------
struct RankBounds { uint64_t b[MAX_PEERS + 1]; int nranks; };

__device__ __forceinline__ unsigned owner_of(const RankBounds &rb, uint64_t v)
{
    unsigned o = 0;
    while ((int)o + 1 < rb.nranks && v >= rb.b[o + 1]) o++;
    return o;
}

// owner side: which ranks read w[v]?  Bit c is set when v has an out-entry into rank c's row block.  A thread per
// out-entry of the own vertices (their row ids expanded by a max-scan first); the mask lives at the vertex's offset
// inside the own segment and is only touched by an atomic when the bit is still missing.
__global__ void k_pt_row_heads(const uint64_t *__restrict__ rowptr, uint64_t v0, uint64_t v1, uint64_t e0, uint32_t *__restrict__ row_of)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < v1; v += stride) {
        const uint64_t a = rowptr[v];
        if (rowptr[v + 1] > a) row_of[a - e0] = (uint32_t)v;
    }
}

struct MaxOfU32 {
    __host__ __device__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

__global__ void __launch_bounds__(256)
k_pt_needmask(const uint32_t *__restrict__ out_col, const uint32_t *__restrict__ row_of, uint64_t count, const RankBounds rb,
              const uint32_t *__restrict__ pi, uint64_t seg_first, uint32_t *__restrict__ needmask)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < count; e += stride) {
        const unsigned bit = 1u << owner_of(rb, ld_stream(out_col + e));
        const uint64_t x = pi[row_of[e]] - seg_first;
        if (!(needmask[x] & bit)) atomicOr(&needmask[x], bit);
    }
}

struct BitOf {
    unsigned bit;
    __host__ __device__ uint8_t operator()(uint32_t m) const { return (uint8_t)((m >> bit) & 1u); }
};

// consumer side: the slots the rows of this rank's block gather from
__global__ void k_pt_mark_needed(const uint32_t *__restrict__ col, uint64_t count, const uint32_t *__restrict__ pi,
                                 uint8_t *__restrict__ mark)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < count; e += stride) {
        const uint32_t s = pi[ld_stream(col + e)];
        if (!mark[s]) mark[s] = 1; // (most entries find their source marked already: a load instead of a store)
    }
}

// pi: global slot -> local id (0xFFFFFFFF for slots this rank does not keep)
__global__ void k_pt_localise(uint32_t *__restrict__ pi, uint64_t n, const uint8_t *__restrict__ mark, const uint32_t *__restrict__ loc)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) { const uint32_t s = pi[v]; pi[v] = mark[s] ? loc[s] : 0xFFFFFFFFu; }
}

struct LocalBases { uint64_t b[MAX_PEERS + 1]; int nranks; };

// stored id of an entry in the compact form: the first hps local slots of every owner's range are the hot stage
__global__ void k_pt_relabel_compact(const uint32_t *__restrict__ col, const uint32_t *__restrict__ pil, uint64_t count,
                                     uint64_t padded, uint32_t zero_local, const LocalBases lb, uint32_t hps, uint32_t hot,
                                     uint32_t *__restrict__ out)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < padded; e += stride) {
        uint32_t id = zero_local + hot;
        if (e < count) {
            const uint32_t l = pil[col[e]];
            unsigned o = 0;
            while ((int)o + 1 < lb.nranks && l >= lb.b[o + 1]) o++;
            const uint32_t j = l - (uint32_t)lb.b[o];
            id = j < hps ? o * hps + j : l + hot;
        }
        out[e] = id;
    }
}

// The per-iteration exchange of the compact form: for every consumer, the slots it reads out of this rank's finished
// segment are gathered (ascending offsets: mostly neighbouring lines) and stored as one contiguous range into the
// consumer's buffer over NVLink.  Consumer p is served by the CTAs with blockIdx.x % npeer == p.
struct PushLists {
    const uint32_t *list[MAX_PEERS - 1];
    uint64_t count[MAX_PEERS - 1];
    double *dst[MAX_PEERS - 1];
    int npeer;
};
__global__ void __launch_bounds__(256) k_pr_push_lists(const double *__restrict__ own, const PushLists pl)
{
    if (pl.npeer == 0) return;
    const int p = blockIdx.x % pl.npeer;
    const unsigned rank_in_peer = blockIdx.x / pl.npeer, ctas_per_peer = gridDim.x / pl.npeer;
    const uint32_t *__restrict__ list = pl.list[p];
    double *__restrict__ dst = pl.dst[p];
    const uint64_t cnt = pl.count[p];
    for (uint64_t i = (uint64_t)rank_in_peer * 256 + threadIdx.x; i < cnt; i += (uint64_t)ctas_per_peer * 256) dst[i] = own[list[i]];
}

__global__ void k_pt_tile_k0(const uint64_t *__restrict__ ne_ptr, uint64_t K, uint64_t n_tiles, uint32_t *__restrict__ tile_k0)
{
    uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; t < n_tiles; t += stride) {
        const uint64_t B = t * PT_TILE;
        uint64_t lo = 0, hi = K; // largest k with ne_ptr[k] <= B
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (ne_ptr[mid] <= B) lo = mid; else hi = mid;
        }
        tile_k0[t] = (uint32_t)lo | (ne_ptr[lo] == B ? 0x80000000u : 0u);
    }
}

// one bit per entry that starts a row, except entries that open a tile (those are tile_k0's bit 31)
__global__ void k_pt_mask(const uint64_t *__restrict__ ne_ptr, uint64_t K, uint32_t *__restrict__ mask)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; k < K; k += stride) {
        const uint64_t pos = ne_ptr[k];
        if (pos % PT_TILE) atomicOr(&mask[pos >> 5], 1u << (pos & 31));
    }
}

__global__ void k_pt_slots(const uint32_t *__restrict__ ne_rows, const uint32_t *__restrict__ pi, uint64_t K,
                           uint32_t *__restrict__ slot_k)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; k < K; k += stride) slot_k[k] = pi ? pi[ne_rows[k]] : ne_rows[k];
}

// per call: 1 / d in non-empty-row order (0 for a sink), so that a tile's epilogue operands are indexed by k
// and w' = r * (1/d) costs one multiply instead of an FP64 division sequence (~30 instructions) per row
__global__ void k_pt_gather_d(const double *__restrict__ d, const uint32_t *__restrict__ ne_rows, uint64_t K,
                              double *__restrict__ inv_k)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; k < K; k += stride) { const double dv = d[ne_rows[k]]; inv_k[k] = dv == 0.0 ? 0.0 : 1.0 / dv; }
}

// epilogue of non-empty row k: r = teleport' + s, w' = r / d, sink mass
template <bool NO_LOADS, bool PEERS>
__device__ __forceinline__ void pt_close(uint32_t k, double s, double tele, const double *__restrict__ d_k,
                                         const uint32_t *__restrict__ slot_k, const uint32_t *__restrict__ ne_rows,
                                         const WOutV &w_new, double *__restrict__ rank, double &sink)
{
    const double r = tele + s;
    const double iv = NO_LOADS ? 0.5 : d_k[k]; // 1 / d, 0 for a sink
    const uint32_t slot = NO_LOADS ? k : slot_k[k];
    if (iv == 0.0) sink += r;
    w_store_v<PEERS>(w_new, slot, r * iv);
    if (rank) rank[ne_rows[k]] = r;
}

__global__ void k_pt_collect_span(const uint64_t *__restrict__ ne_ptr, uint64_t K, uint32_t *__restrict__ list,
                                  unsigned long long *__restrict__ count, uint64_t cap)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; k < K; k += stride) {
        if (ne_ptr[k] / PT_TILE != (ne_ptr[k + 1] - 1) / PT_TILE) {
            const unsigned long long pos = atomicAdd(count, 1ull);
            if (pos < cap) list[pos] = (uint32_t)k;
        }
    }
}

// s += v if bit J of m is set -- one predicate-setting logic op and one predicated DADD; written as
// `if (...) s += v` the compiler selects 0.0 first (two compares, two FSELs and the add per entry)
template <int J>
__device__ __forceinline__ void add_if_bit(double &s, double v, uint32_t m)
{
    asm("{\n\t.reg .pred p;\n\t.reg .b32 t;\n\tand.b32 t, %2, %3;\n\tsetp.ne.u32 p, t, 0;\n\t@p add.f64 %0, %0, %1;\n\t}"
        : "+d"(s)
        : "d"(v), "r"(m), "n"(1u << J));
}

// sum of the lane's values whose positions are set in m, in ascending order
__device__ __forceinline__ double masked_sum8(const double (&val)[8], uint32_t m)
{
    double s = 0.0;
    add_if_bit<0>(s, val[0], m); add_if_bit<1>(s, val[1], m); add_if_bit<2>(s, val[2], m); add_if_bit<3>(s, val[3], m);
    add_if_bit<4>(s, val[4], m); add_if_bit<5>(s, val[5], m); add_if_bit<6>(s, val[6], m); add_if_bit<7>(s, val[7], m);
    return s;
}

struct PtArgs {
    const uint32_t *col;
    const uint64_t *ne_ptr;
    const uint32_t *ne_rows;
    const uint32_t *tile_k0;
    const uint8_t *mask;    // 32 bytes per tile
    const uint32_t *slot_k;
    const double *d_k;
    const double *w;        // degree-sorted space
    const double *w_cold;   // w - hot: stored ids >= hot are slot + hot
    uint32_t hps, wps;      // hot entries per rank segment, slots of a rank's warm piece (PrTiles)
    uint32_t hot_base[MAX_PEERS]; // first slot of owner o's hottest entries (o * wps; compact form: lbase[o])
    uint32_t warm_end;      // stored ids in [hot, warm_end) ask L2 to keep their lines, the rest are read evict-first
    const double *sink_in;  // sink partials of the previous step (one scalar after the multi-GPU all-reduce)
    unsigned n_sink_in;
    double *tele_out;
    WOutV w_new;            // where w' goes: this rank's copy (+ every peer's on several GPUs, fused exchange)
    double *rank;
    double *head_part;
    double *tail_part;
    double *sink_out;
    uint64_t K, M, n_tiles;
    uint32_t hot;
    PrScalars sc;
    // rows without entries (r = teleport'): done by the warps of k_pr_tiles once they run out of tiles
    const uint32_t *empty_v, *empty_slot;
    const double *empty_inv;
    uint64_t n_empty;
};

// One 256-entry tile whose 8 values per lane are in registers: row sums from the row-start bits
// (in-lane pieces, warp segmented scan for rows crossing lanes, head / tail partials for rows crossing
// tiles) and the fused epilogue r -> w' = r / d, sink mass.
template <int VAR, bool PEERS>
__device__ __forceinline__ void pt_tile_rows(const PtArgs &a, uint64_t t, unsigned lane, uint32_t kk, uint32_t kk_next,
                                             uint32_t flags, const double (&val)[8], double tele, double &sink)
{
    const uint32_t k0 = kk & 0x7FFFFFFFu;
    const bool k0_starts_here = (kk >> 31) != 0;   // otherwise row k0 began in an earlier tile
    const bool ends_here = (kk_next >> 31) != 0;   // a row starts right after this tile (or the block ends)
    // rows starting before this lane's first entry (exclusive prefix of the per-lane counts)
    const uint32_t cnt = __popc(flags);
    uint32_t incl = cnt;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const uint32_t up = __shfl_up_sync(FULL, incl, dlt);
        if (lane >= (unsigned)dlt) incl += up;
    }
    const uint32_t before = incl - cnt;
    const uint32_t total_starts = __shfl_sync(FULL, incl, 31);
    // ---- rows that end inside the lane: one trip per row start (a tile has ~10, a lane 0..2), so
    // the epilogue code runs a couple of times per tile instead of once per entry position
    uint32_t kcur = k0 + before; // row the lane's first entry belongs to
    uint32_t f = flags, i0 = 0;
    double head = 0.0;
    bool seen = false;
    while (__any_sync(FULL, f != 0)) {
        if (f) {
            const uint32_t i = __ffs(f) - 1;
            f &= f - 1;
            const double s = masked_sum8(val, ((1u << i) - 1u) & ~((1u << i0) - 1u)); // entries [i0, i)
            if (!seen) { head = s; seen = true; } // the row running into the lane: closed after the scan
            else pt_close<(VAR & 2) != 0, PEERS>(kcur, s, tele, a.d_k, a.slot_k, a.ne_rows, a.w_new, a.rank, sink);
            kcur++;
            i0 = i;
        }
    }
    const double acc = masked_sum8(val, 0xFFu & ~((1u << i0) - 1u)); // open sum at the end of the lane: entries [i0, 8)
    // ---- rows crossing lanes: segmented inclusive scan of the lanes' open sums
    double sv = acc;
    bool sf = seen;
#pragma unroll
    for (int dlt = 1; dlt < 32; dlt <<= 1) {
        const double pv = __shfl_up_sync(FULL, sv, dlt);
        const int pf = __shfl_up_sync(FULL, (int)sf, dlt);
        if (lane >= (unsigned)dlt) { if (!sf) sv += pv; sf = sf || pf; }
    }
    double carry = __shfl_up_sync(FULL, sv, 1);
    if (lane == 0) carry = 0.0;
    if (seen) {
        // the row running into this lane ends at the lane's first row start
        const double tot = carry + head;
        if (before == 0 && !k0_starts_here) a.head_part[t] = tot;
        else pt_close<(VAR & 2) != 0, PEERS>(k0 + before, tot, tele, a.d_k, a.slot_k, a.ne_rows, a.w_new, a.rank, sink);
    }
    if (lane == 31) {
        // the row still open at the end of the tile
        const uint32_t klast = k0 + total_starts;
        const bool began_here = total_starts > 0 || k0_starts_here;
        if (ends_here) {
            if (began_here) pt_close<(VAR & 2) != 0, PEERS>(klast, sv, tele, a.d_k, a.slot_k, a.ne_rows, a.w_new, a.rank, sink);
            else a.head_part[t] = sv;
        } else {
            if (began_here) a.tail_part[t] = sv; else a.head_part[t] = sv;
        }
    }
}

__device__ __forceinline__ double ld_gather_policy(const double *p, uint64_t policy)
{
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
    return v;
}

// the 8 gathers of a lane: hottest sources from shared memory, the rest through L1/L2.  HINT (w larger than about
// half of L2): the warm range is read with an evict-last policy, the rarely gathered tail evict-first, so that the
// tail's lines (most of the vector, a tenth of the gathers) do not push the warm ones out of L2.
template <int VAR, bool HINT>
__device__ __forceinline__ void pt_gather8(const PtArgs &a, const double *s_hot, const uint32_t (&idx)[8], double (&val)[8],
                                           uint64_t pol_keep, uint64_t pol_stream)
{
#pragma unroll
    for (int i = 0; i < 8; i++) {
        if (VAR & 4) val[i] = (double)idx[i];
        else if (idx[i] < a.hot) val[i] = s_hot[idx[i]];
        else if (!HINT) val[i] = __ldg(a.w_cold + idx[i]);
        else if (idx[i] < a.warm_end) val[i] = ld_gather_policy(a.w_cold + idx[i], pol_keep);
        else val[i] = ld_gather_policy(a.w_cold + idx[i], pol_stream);
    }
}

// VAR (GX_PR_VAR): 2 and 4 are TIMING DIAGNOSTICS that break the result (tools/pr_ab.py): 2 = epilogue
// operands not loaded, 4 = no gathers -- what each dependent memory phase of a tile costs.
template <int VAR, bool PEERS, bool HINT>
__global__ void __launch_bounds__(PT_WARPS * 32, 1) k_pr_tiles(const PtArgs a)
{
    uint64_t pol_keep = 0, pol_stream = 0;
    if (HINT) {
        asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol_keep));
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol_stream));
    }
    extern __shared__ double s_hot[];
    __shared__ double s_red[PT_WARPS + 1];
    // teleport' = (1-damping)/n + damping * sum_{sinks} r / n; every CTA adds the partials in the same order
    const double tele = a.sc.teleport + a.sc.damping * block_sum_ordered(a.sink_in, a.n_sink_in, s_red) / a.sc.n;
    if (blockIdx.x == 0 && threadIdx.x == 0) *a.tele_out = tele;
    // the hottest sources (highest out-degree) are read from shared memory instead of through L1/L2
    for (uint32_t h = threadIdx.x; h < a.hot; h += PT_WARPS * 32) {
        const uint32_t owner = h / a.hps;
        s_hot[h] = a.w[(uint64_t)a.hot_base[owner] + (h - owner * a.hps)]; // one GPU: w[h]
    }
    __syncthreads();
    const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
    const uint64_t nwarp = (uint64_t)gridDim.x * PT_WARPS;
    double sink = 0.0;
    for (uint64_t t = (uint64_t)blockIdx.x * PT_WARPS + wib; t < a.n_tiles; t += nwarp) {
        // ---- everything the tile needs is requested at once: first row, row-start bits, column ids
        const uint32_t kk = a.tile_k0[t];
        const uint32_t kk_next = (t + 1 < a.n_tiles) ? a.tile_k0[t + 1] : 0x80000000u;
        const uint32_t flags = a.mask[t * 32 + lane]; // bit i: a row starts at entry 8*lane+i
        uint32_t idx[8];
        ld_stream8(a.col + t * PT_TILE + 8u * lane, idx); // tiles are whole (padded) and 1 KB apart: 32-byte aligned
        double val[8];
        pt_gather8<VAR, HINT>(a, s_hot, idx, val, pol_keep, pol_stream); // 8 independent gathers per lane
        pt_tile_rows<VAR, PEERS>(a, t, lane, kk, kk_next, flags, val, tele, sink);
    }
    // rows without entries: r = teleport'; a thread each, spread over all warps of the grid (the tail of
    // the persistent kernel, when the LSU pipe is no longer busy with gathers)
    for (uint64_t i = ((uint64_t)blockIdx.x * PT_WARPS + wib) * 32 + lane; i < a.n_empty; i += nwarp * 32) {
        const double iv = a.empty_inv[i];
        if (iv == 0.0) sink += tele;
        w_store_v<PEERS>(a.w_new, a.empty_slot[i], tele * iv);
        if (a.rank) a.rank[a.empty_v[i]] = tele;
    }
    __syncthreads();
    sink = warp_sum(sink);
    if (lane == 0) s_red[wib] = sink;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0;
        for (int i = 0; i < PT_WARPS; i++) x += s_red[i];
        a.sink_out[blockIdx.x] = x;
    }
}

// per-thread operands of k_pr_tile_fin: tile-crossing rows first, then rows without entries
__global__ void k_pt_fin_plan(const uint64_t *__restrict__ ne_ptr, const uint32_t *__restrict__ ne_rows,
                              const uint32_t *__restrict__ span_k, uint64_t n_span, const uint32_t *__restrict__ empty_rows,
                              uint64_t n_empty, const uint32_t *__restrict__ pi, uint32_t *__restrict__ fin_v,
                              uint32_t *__restrict__ fin_slot, uint32_t *__restrict__ fin_t0, uint32_t *__restrict__ fin_nt)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < n_span + n_empty; i += stride) {
        uint32_t v, t0 = 0, nt = 0;
        if (i < n_span) {
            const uint32_t k = span_k[i];
            v = ne_rows[k];
            t0 = (uint32_t)(ne_ptr[k] / PT_TILE);
            nt = (uint32_t)((ne_ptr[k + 1] - 1) / PT_TILE) - t0;
        } else {
            v = empty_rows[i - n_span];
        }
        fin_v[i] = v; fin_slot[i] = pi[v]; fin_t0[i] = t0; fin_nt[i] = nt;
    }
}

// rows crossing tile borders (tail of the first tile + heads of the next ones, added in tile
// order); one thread each, all operands indexed by the thread
constexpr uint32_t FIN_LONG = 32;
template <bool PEERS>
__global__ void __launch_bounds__(256)
k_pr_tile_fin(const uint32_t *__restrict__ fin_v, const uint32_t *__restrict__ fin_slot, const uint32_t *__restrict__ fin_t0,
              const uint32_t *__restrict__ fin_nt, const double *__restrict__ fin_d, uint64_t n_span, uint64_t n_fin,
              const double *__restrict__ head_part, const double *__restrict__ tail_part,
              const double *__restrict__ tele_p, double *__restrict__ w_new, const WOut *__restrict__ peers,
              double *__restrict__ rank, double *__restrict__ sink_part)
{
    const double tele = *tele_p;
    double sink = 0.0;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; // one row per thread
    const bool live = i < n_fin;
    uint32_t t0 = 0, nt = 0;
    double s = 0.0;
    if (live && i < n_span) {
        t0 = fin_t0[i];
        nt = fin_nt[i];
        s = tail_part[t0];
        if (nt <= FIN_LONG) {
#pragma unroll 4
            for (uint32_t t = 1; t <= nt; t++) s += head_part[t0 + t];
        }
    }
    // hub rows span hundreds of tiles: the warp adds their partials together (lane-strided loads, then a
    // fixed-order tree), instead of one thread walking them one dependent load after the other
    for (unsigned todo = __ballot_sync(FULL, nt > FIN_LONG); todo; todo &= todo - 1) {
        const int src = __ffs(todo) - 1;
        const uint32_t bt0 = __shfl_sync(FULL, t0, src), bnt = __shfl_sync(FULL, nt, src);
        double p = 0.0;
#pragma unroll 4
        for (uint32_t t = 1 + lane_id(); t <= bnt; t += 32) p += head_part[bt0 + t];
        p = warp_sum(p);
        if ((int)lane_id() == src) s += p;
    }
    if (live) {
        const double dv = fin_d[i]; // 1 / d, 0 for a sink
        const double r = tele + s;
        if (dv == 0.0) sink += r;
        w_store<PEERS>(w_new, peers, fin_slot[i], r * dv);
        if (rank) rank[fin_v[i]] = r;
    }
    __shared__ double red[8];
    sink = warp_sum(sink);
    if (lane_id() == 0) red[threadIdx.x >> 5] = sink;
    __syncthreads();
    if (threadIdx.x == 0) {
        double x = 0.0;
        for (unsigned i2 = 0; i2 < (blockDim.x >> 5); i2++) x += red[i2];
        sink_part[blockIdx.x] = x;
    }
}

// GX_TIMING_DEBUG=1: wall-clock per phase of the plan construction on stderr (rank 0)
struct PlanLog {
    bool on = getenv("GX_TIMING_DEBUG") != nullptr && ctx().rank == 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(ctx().stream);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gx timing] PR plan: %-24s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

struct Widen8 {
    __host__ __device__ uint32_t operator()(uint8_t x) const { return x; }
};

// compact form: the slots this rank keeps (consumer side), what it sends to whom (owner side), and pi rewritten to
// local ids.  Collective: the count matrix (owner x consumer) is all-gathered.
static void build_compact(gx_graph *g, PrTiles *pt, const Adj &in, uint64_t v0, uint64_t v1, uint64_t e0)
{
    Context &c = ctx();
    const int N = c.nranks, r = c.rank;
    const uint64_t seg = pt->seg, n = g->n, slots = pt->slots;
    PlanLog log;
    RankBounds rb;
    rb.nranks = N;
    for (int i = 0; i <= N; i++) rb.b[i] = in.plan.part.b[i];
    // ---- owner side: who reads which of my slots
    DevBuf<uint32_t> needmask(seg);
    needmask.zero();
    if (v1 > v0) {
        uint64_t oe[2];
        read_back(&oe[0], g->out.rowptr.p + v0, sizeof(uint64_t));
        read_back(&oe[1], g->out.rowptr.p + v1, sizeof(uint64_t));
        const uint64_t cnt = oe[1] - oe[0];
        if (cnt) {
            DevBuf<uint32_t> row_of(cnt);
            GX_CUDA(cudaMemsetAsync(row_of.p, 0, cnt * sizeof(uint32_t), c.stream));
            GX_LAUNCH(k_pt_row_heads, grid_persistent(8), 256, 0, g->out.rowptr.p, v0, v1, oe[0], row_of.p);
            size_t tb = 0;
            GX_CUDA(cub::DeviceScan::InclusiveScan(nullptr, tb, row_of.p, row_of.p, MaxOfU32(), (int64_t)cnt, c.stream));
            DevBuf<char> tmp(tb);
            GX_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, tb, row_of.p, row_of.p, MaxOfU32(), (int64_t)cnt, c.stream));
            count_launch();
            GX_LAUNCH(k_pt_needmask, grid_persistent(8), 256, 0, g->out.col.p + oe[0], row_of.p, cnt, rb, pt->pi.p, (uint64_t)r * seg,
                      needmask.p);
            GX_CUDA(cudaStreamSynchronize(c.stream)); // the scoped buffers
        }
    }
    log.mark("need masks");
    std::vector<uint64_t> row(N, 0);
    {
        DevBuf<uint32_t> lists(seg * (uint64_t)(N - 1));
        DevBuf<uint64_t> nsel(1);
        thrust::counting_iterator<uint32_t> ids(0);
        size_t tb = 0;
        {
            auto flags = thrust::make_transform_iterator((const uint32_t *)needmask.p, BitOf{0u});
            GX_CUDA(cub::DeviceSelect::Flagged(nullptr, tb, ids, flags, lists.p, nsel.p, (int64_t)seg, c.stream));
        }
        DevBuf<char> tmp(tb);
        uint64_t off = 0;
        for (int cns = 0; cns < N; cns++) {
            pt->push_off[cns] = off;
            if (cns == r) { row[cns] = seg; continue; }
            auto flags = thrust::make_transform_iterator((const uint32_t *)needmask.p, BitOf{(unsigned)cns});
            GX_CUDA(cub::DeviceSelect::Flagged(tmp.p, tb, ids, flags, lists.p + off, nsel.p, (int64_t)seg, c.stream));
            count_launch();
            uint64_t cnt = 0;
            read_back(&cnt, nsel.p, sizeof(cnt));
            row[cns] = cnt;
            off += cnt;
        }
        pt->push_off[N] = off;
        pt->push_list.alloc(off ? off : 1);
        if (off) GX_CUDA(cudaMemcpyAsync(pt->push_list.p, lists.p, off * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
        GX_CUDA(cudaStreamSynchronize(c.stream)); // `lists` goes out of scope
    }
    log.mark("push lists");
    // ---- everybody's counts: mat[o * N + cns] = slots of owner o that consumer cns keeps
    std::vector<uint64_t> mat((size_t)N * N);
    {
        DevBuf<uint64_t> dm((size_t)N * N);
        GX_CUDA(cudaMemcpyAsync(dm.p + (size_t)r * N, row.data(), N * sizeof(uint64_t), cudaMemcpyHostToDevice, c.stream));
        allgather_equal(dm.p, Dt::U64, (uint64_t)N);
        read_back(mat.data(), dm.p, mat.size() * sizeof(uint64_t));
    }
    log.mark("count matrix");
    // ---- consumer side: the slots my rows gather from (+ my own segment), numbered in slot order
    DevBuf<uint8_t> mark(slots + 1);
    mark.zero();
    GX_CUDA(cudaMemsetAsync(mark.p + (uint64_t)r * seg, 1, seg, c.stream));
    if (pt->M) GX_LAUNCH(k_pt_mark_needed, grid_persistent(8), 256, 0, in.col.p + e0, pt->M, pt->pi.p, mark.p);
    DevBuf<uint32_t> loc(slots + 1);
    {
        auto wide = thrust::make_transform_iterator((const uint8_t *)mark.p, Widen8{});
        size_t tb = 0;
        GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, wide, loc.p, (int64_t)(slots + 1), c.stream));
        DevBuf<char> tmp(tb);
        GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, wide, loc.p, (int64_t)(slots + 1), c.stream));
        count_launch();
    }
    log.mark("mark + scan");
    for (int o = 0; o <= N; o++) {
        uint32_t x = 0;
        read_back(&x, loc.p + (uint64_t)o * seg, sizeof(x));
        pt->lbase[o] = x;
    }
    pt->L = pt->lbase[N];
    for (int o = 0; o < N; o++)
        GX_REQUIRE(pt->lbase[o + 1] - pt->lbase[o] == mat[(size_t)o * N + r],
                   "PageRank compact plan: owner and consumer disagree on the slots to exchange");
    pt->Lmax = 0;
    for (int cns = 0; cns < N; cns++) {
        uint64_t tot = 0;
        for (int o = 0; o < N; o++) tot += mat[(size_t)o * N + cns];
        pt->Lmax = std::max(pt->Lmax, tot);
        uint64_t before = 0;
        for (int o = 0; o < r; o++) before += mat[(size_t)o * N + cns];
        pt->push_dst[cns] = before;
    }
    GX_LAUNCH(k_pt_localise, grid_persistent(8), 256, 0, pt->pi.p, n, mark.p, loc.p);
    GX_CUDA(cudaStreamSynchronize(c.stream));
    log.mark("local ids");
}

static PrTiles *build_pr_tiles(gx_graph *g)
{
    PrTiles *pt = new PrTiles();
    Adj &in = g->in_adj();
    const uint64_t v0 = in.plan.part.lo, v1 = in.plan.part.hi, nv = v1 - v0;
    uint64_t ends[2] = {0, in.col.n}; // the whole adjacency on one GPU
    if (multi()) {
        read_back(&ends[0], in.rowptr.p + v0, sizeof(uint64_t));
        read_back(&ends[1], in.rowptr.p + v1, sizeof(uint64_t));
    }
    const uint64_t e0 = ends[0];
    pt->M = ends[1] - ends[0];
    pt->n_tiles = (pt->M + PT_TILE - 1) / PT_TILE;
    // rank of every row among the non-empty rows of the block
    DevBuf<uint32_t> flag(nv ? nv : 1), rank_ne(nv ? nv : 1);
    uint32_t tail[2] = {0, 0};
    if (nv) {
        GX_LAUNCH(k_pt_mark, grid_persistent(8), 256, 0, in.rowptr.p, v0, v1, flag.p);
        size_t tb = 0;
        GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tb, flag.p, rank_ne.p, (int64_t)nv, ctx().stream));
        DevBuf<char> tmp(tb);
        GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tb, flag.p, rank_ne.p, (int64_t)nv, ctx().stream));
        read_back(&tail[0], rank_ne.p + (nv - 1), sizeof(uint32_t));
        read_back(&tail[1], flag.p + (nv - 1), sizeof(uint32_t));
    }
    pt->K = (uint64_t)tail[0] + tail[1];
    GX_REQUIRE(pt->K < 0x80000000ull, "PageRank tiles need fewer than 2^31 non-empty rows per rank (bit 31 of tile_k0 is a flag)");
    pt->n_empty = nv - pt->K;
    pt->ne_rows.alloc(pt->K ? pt->K : 1);
    pt->ne_ptr.alloc(pt->K + 1);
    pt->empty_rows.alloc(pt->n_empty ? pt->n_empty : 1);
    if (nv)
        GX_LAUNCH(k_pt_fill, grid_persistent(8), 256, 0, in.rowptr.p, v0, v1, rank_ne.p, e0, pt->ne_rows.p, pt->ne_ptr.p,
                  pt->empty_rows.p);
    GX_CUDA(cudaMemcpyAsync(pt->ne_ptr.p + pt->K, &pt->M, sizeof(uint64_t), cudaMemcpyHostToDevice, ctx().stream));
    GX_CUDA(cudaStreamSynchronize(ctx().stream)); // &pt->M is read by the copy
    // out-degree order of the sources: w lives in that index space, its head goes to shared memory
    {
        const uint64_t n = g->n;
        pt->pi.alloc(n);
        DevBuf<uint32_t> keys(n), keys_alt(n), vtx(n), vtx_alt(n);
        DevBuf<uint64_t> bounds(in.plan.part.b.size());
        GX_CUDA(cudaMemcpyAsync(bounds.p, in.plan.part.b.data(), in.plan.part.b.size() * sizeof(uint64_t), cudaMemcpyHostToDevice,
                                ctx().stream));
        GX_REQUIRE(ctx().nranks <= 256, "too many ranks");
        GX_LAUNCH(k_pt_degree_keys, grid_persistent(8), 256, 0, g->out.rowptr.p, n, bounds.p, ctx().nranks, keys.p, vtx.p);
        cub::DoubleBuffer<uint32_t> dk(keys.p, keys_alt.p), dv(vtx.p, vtx_alt.p);
        {
            const int end_bit = 24 + (ctx().nranks > 1 ? bits_for((uint64_t)ctx().nranks) : 0);
            size_t tb = 0;
            GX_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, (int64_t)n, 0, end_bit, ctx().stream));
            DevBuf<char> tmp(tb);
            GX_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tb, dk, dv, (int64_t)n, 0, end_bit, ctx().stream));
            count_launch();
        }
        uint64_t seg = 0;
        for (int r = 0; r < ctx().nranks; r++) seg = std::max<uint64_t>(seg, in.plan.part.b[r + 1] - in.plan.part.b[r]);
        seg = (seg + 31) & ~31ull;
        uint32_t hot_cap = PT_HOT;
        if (const char *e = getenv("GX_PR_HOT")) hot_cap = (uint32_t)atoi(e) < PT_HOT ? (uint32_t)atoi(e) : PT_HOT; // tuning knob
        pt->hps = (uint32_t)std::min<uint64_t>(seg, hot_cap / (uint32_t)ctx().nranks);
        pt->hot = pt->hps * (uint32_t)ctx().nranks;
        // compact form (see PrTiles): from 3 ranks on when the peer mapping works; GX_PR_COMPACT=0 / 1 forces it off / on
        {
            const char *ce = getenv("GX_PR_COMPACT");
            bool want = multi() && ctx().nranks <= MAX_PEERS && (ce ? ce[0] != '0' : ctx().nranks >= 3);
            if (want) want = context_mail().ok; // (collective: every rank evaluates the same condition)
            pt->compact = want;
        }
        // warm piece: the often-gathered head of every segment, PR_WARM_MB in total (GX_PR_WARM_MB; 0 = one piece)
        uint64_t warm_mb = PT_WARM_MB;
        if (const char *e = getenv("GX_PR_WARM_MB")) warm_mb = (uint64_t)atoll(e);
        if (pt->compact) warm_mb = 0; // the kept part of w fits L2: one piece, one load policy
        uint64_t wps = ((warm_mb << 20) / sizeof(double) / (uint64_t)ctx().nranks) & ~31ull;
        wps = std::max<uint64_t>(wps, (pt->hps + 31u) & ~31u);
        if (warm_mb == 0 || wps >= seg) wps = seg;
        pt->wps = wps;
        pt->tps = seg - wps;
        pt->seg = seg;
        pt->slots = pt->seg * (uint64_t)ctx().nranks;
        pt->hint = pt->tps > 0 && pt->slots * sizeof(double) > (PT_HINT_MIN_MB << 20);
        if (const char *e = getenv("GX_PR_HINT")) pt->hint = e[0] != '0';
        if (pt->compact) pt->hint = false;
        GX_REQUIRE(pt->slots + pt->hot < 0xFFFFFFFEull, "vertex slot space exceeds 32 bits");
        GX_LAUNCH(k_pt_make_pi, grid_persistent(8), 256, 0, dk.Current(), dv.Current(), n, bounds.p, pt->wps, pt->tps,
                  (uint64_t)ctx().nranks, pt->pi.p);
    }
    if (pt->compact) build_compact(g, pt, in, v0, v1, e0);
    // this rank's slice of the column ids as pi(source), tile-aligned at offset 0
    // (padded to whole tiles with the index of a slot that always holds 0: the kernel needs no bounds checks)
    const uint64_t padded = pt->n_tiles * PT_TILE;
    pt->col.alloc(padded ? padded : 1);
    if (pt->compact) {
        LocalBases lb;
        lb.nranks = ctx().nranks;
        for (int o = 0; o <= ctx().nranks; o++) lb.b[o] = pt->lbase[o];
        GX_LAUNCH(k_pt_relabel_compact, grid_persistent(8), 256, 0, in.col.p + e0, pt->pi.p, pt->M, padded, (uint32_t)pt->L, lb,
                  pt->hps, pt->hot, pt->col.p);
    } else if (pt->M)
        GX_LAUNCH(k_pt_relabel_slice, grid_persistent(8), 256, 0, in.col.p + e0, pt->pi.p, pt->M, padded, (uint32_t)pt->slots,
                  (uint32_t)pt->wps, (uint32_t)(pt->wps * (uint64_t)ctx().nranks), pt->hps, pt->hot, pt->col.p);
    pt->tile_k0.alloc(pt->n_tiles ? pt->n_tiles : 1);
    if (pt->n_tiles)
        GX_LAUNCH(k_pt_tile_k0, grid_for(pt->n_tiles, 256), 256, 0, pt->ne_ptr.p, pt->K, pt->n_tiles, pt->tile_k0.p);
    pt->mask.alloc(pt->n_tiles ? pt->n_tiles * (PT_TILE / 32) : 1);
    pt->mask.zero();
    pt->slot_k.alloc(pt->K ? pt->K : 1);
    if (pt->K) {
        GX_LAUNCH(k_pt_mask, grid_persistent(8), 256, 0, pt->ne_ptr.p, pt->K, pt->mask.p);
        GX_LAUNCH(k_pt_slots, grid_persistent(8), 256, 0, pt->ne_rows.p, pt->pi.p, pt->K, pt->slot_k.p);
    }
    DevBuf<unsigned long long> cnt(1);
    cnt.zero();
    pt->span_k.alloc(pt->n_tiles ? pt->n_tiles : 1); // a spanning row crosses a tile border: at most n_tiles of them
    if (pt->K) GX_LAUNCH(k_pt_collect_span, grid_persistent(8), 256, 0, pt->ne_ptr.p, pt->K, pt->span_k.p, cnt.p, pt->n_tiles);
    unsigned long long ns = 0;
    read_back(&ns, cnt.p, sizeof(ns));
    pt->n_span = ns;
    sort_keys32(pt->span_k, ns, bits_for(pt->K + 1)); // the atomics appended in any order; a fixed order keeps the sink sums reproducible
    const uint64_t n_fin = pt->n_span + pt->n_empty;
    pt->fin_v.alloc(n_fin ? n_fin : 1); pt->fin_slot.alloc(n_fin ? n_fin : 1);
    pt->fin_t0.alloc(n_fin ? n_fin : 1); pt->fin_nt.alloc(n_fin ? n_fin : 1);
    if (n_fin)
        GX_LAUNCH(k_pt_fin_plan, grid_persistent(8), 256, 0, pt->ne_ptr.p, pt->ne_rows.p, pt->span_k.p, pt->n_span,
                  pt->empty_rows.p, pt->n_empty, pt->pi.p, pt->fin_v.p, pt->fin_slot.p, pt->fin_t0.p, pt->fin_nt.p);
    return pt;
}

static void pagerank_tiles(gx_graph *g, double damping, int iters)
{
    Context &c = ctx();
    const uint64_t n = g->n;
    Adj &in = g->in_adj();
    const RowPlan &plan = in.plan;
    const PrTiles &pt = *(PrTiles *)g->pr_cache;
    const uint64_t v0 = plan.part.lo, v1 = plan.part.hi;
    int var = 0;
    if (const char *e = getenv("GX_PR_VAR")) var = atoi(e);
    const uint32_t hot = pt.hot;
    const size_t smem = (size_t)hot * sizeof(double);
    using TilesFn = void (*)(const PtArgs);
    static const TilesFn tiles_tab[4][2] = {{k_pr_tiles<0, false, false>, k_pr_tiles<0, true, false>},
                                            {k_pr_tiles<2, false, false>, k_pr_tiles<2, true, false>},
                                            {k_pr_tiles<4, false, false>, k_pr_tiles<4, true, false>},
                                            {k_pr_tiles<0, false, true>, k_pr_tiles<0, true, true>}};
    const int vi = var == 2 ? 1 : var == 4 ? 2 : pt.hint ? 3 : 0;
    const unsigned tile_threads = PT_WARPS * 32;
    TilesFn tiles_fn[2] = {tiles_tab[vi][0], tiles_tab[vi][1]};
    GX_CUDA(cudaFuncSetAttribute(tiles_fn[0], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    GX_CUDA(cudaFuncSetAttribute(tiles_fn[1], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PrTiles &ptm = *(PrTiles *)g->pr_cache;
    // slots of a buffer: the whole index space + the always-zero slot the tile padding gathers; compact form: the
    // largest local space over the ranks (equal sizes keep the ranks' parked-buffer caches in step) + room for the
    // hot-stage load to run past a short owner range
    const uint64_t wcap = pt.compact ? pt.Lmax + pt.hot + 1 : pt.slots + 1;
    const uint64_t zero_slot = pt.compact ? pt.L : pt.slots;
    if (!ptm.have_wbuf) {
        peer_alloc(ptm.wbuf[0], wcap * sizeof(double));
        peer_alloc(ptm.wbuf[1], wcap * sizeof(double));
        ptm.have_wbuf = true;
    }
    // fused exchange (peer stores from the kernels) on 2 GPUs -- measured 4 % ahead of the all-gather there, but
    // 8-byte stores scattered over 7 peers lose to it on 8 GPUs (16.1 vs 10.3 ms per PageRank at RMAT-25);
    // GX_PR_FUSED=0 / 1 forces it off / on.  Never without the peer mapping.
    const char *fe = getenv("GX_PR_FUSED");
    const bool want_fused = fe ? fe[0] != '0' : c.nranks <= 2;
    const bool cpush = pt.compact; // compact form: per-consumer gather lists (k_pr_push_lists)
    GX_REQUIRE(!cpush || (pt.wbuf[0].shared && pt.wbuf[1].shared), "PageRank compact plan without mapped peer buffers");
    const bool fused = multi() && !cpush && pt.wbuf[0].shared && pt.wbuf[1].shared && want_fused;
    // otherwise the finished segment is pushed to the peers by one copy kernel (GX_PR_PUSH=0: ncclAllGather)
    const char *pe = getenv("GX_PR_PUSH");
    const bool push = multi() && !cpush && !fused && pt.wbuf[0].shared && pt.wbuf[1].shared && !(pe && pe[0] == '0');
    const uint64_t n_fin = pt.n_span + pt.n_empty;
    DevBuf<double> d(n), w_nat(n), sink_sum(1), tele(1), d_k(pt.K ? pt.K : 1), d_fin(n_fin ? n_fin : 1);
    double *wv[2] = {(double *)pt.wbuf[0].local, (double *)pt.wbuf[1].local};
    if (multi()) { // padding slots are exchanged but never gathered
        GX_CUDA(cudaMemsetAsync(wv[0], 0, wcap * sizeof(double), c.stream));
        GX_CUDA(cudaMemsetAsync(wv[1], 0, wcap * sizeof(double), c.stream));
    }
    GX_CUDA(cudaMemsetAsync(wv[0] + zero_slot, 0, sizeof(double), c.stream)); // the zero slot (never written)
    GX_CUDA(cudaMemsetAsync(wv[1] + zero_slot, 0, sizeof(double), c.stream));
    const unsigned g_tiles = (unsigned)c.num_sms;
    const unsigned g_fin = pt.n_span ? grid_for(pt.n_span, 256) : 0; // rows without entries ride along in k_pr_tiles
    const unsigned g_init = grid_persistent(8);
    const unsigned nparts = (g_tiles + g_fin > g_init) ? g_tiles + g_fin : g_init;
    DevBuf<double> sinkA(nparts), sinkB(nparts), head_part(pt.n_tiles ? pt.n_tiles : 1), tail_part(pt.n_tiles ? pt.n_tiles : 1);
    PhaseTimer tk(&c.timing.kernel_ms);
    // teleport' = (1-damping)/n + damping * sum_{sinks} r / n   (LAGr_PageRankGX)
    const PrScalars sc{(1.0 - damping) / (double)n, damping, (double)n};
    sinkA.zero();
    sinkB.zero();
    GX_LAUNCH(k_pr_init, g_init, 256, 0, g->out.rowptr.p, n, v0, v1, damping, d.p, w_nat.p, sinkA.p);
    GX_LAUNCH(k_pt_scatter, grid_persistent(8), 256, 0, w_nat.p, pt.pi.p, n, wv[0]);
    if (pt.K) GX_LAUNCH(k_pt_gather_d, grid_persistent(8), 256, 0, d.p, pt.ne_rows.p, pt.K, d_k.p);
    if (n_fin) GX_LAUNCH(k_pt_gather_d, grid_persistent(8), 256, 0, d.p, pt.fin_v.p, n_fin, d_fin.p);
    if (iters == 0) GX_LAUNCH(k_fill_f64, grid_persistent(4), 256, 0, g->res_f64.p, n, 1.0 / (double)n);
    int cur = 0;
    double *s_in = sinkA.p, *s_out = sinkB.p;
    // peer pointer tables of the two buffers, read by the kernels from global memory
    DevBuf<WOut> peer_tab(2);
    {
        WOut h[2];
        for (int b = 0; b < 2; b++) {
            h[b].npeer = 0;
            for (int r = 0; (fused || push) && r < c.nranks; r++)
                if (r != c.rank) h[b].peer[h[b].npeer++] = (double *)pt.wbuf[b].peer[r];
        }
        GX_CUDA(cudaMemcpyAsync(peer_tab.p, h, sizeof(h), cudaMemcpyHostToDevice, c.stream));
        GX_CUDA(cudaStreamSynchronize(c.stream));
    }
    // compact form: what each consumer gets, per buffer (the lists are offsets inside the own segment)
    PushLists plists[2];
    for (int b = 0; b < 2; b++) {
        plists[b].npeer = 0;
        for (int r = 0; cpush && r < c.nranks; r++) {
            if (r == c.rank) continue;
            const int k = plists[b].npeer++;
            plists[b].list[k] = pt.push_list.p + pt.push_off[r];
            plists[b].count[k] = pt.push_off[r + 1] - pt.push_off[r];
            plists[b].dst[k] = (double *)pt.wbuf[b].peer[r] + pt.push_dst[r];
        }
    }
    if (fused || push || cpush) {
        // nobody may store into a rank's buffers before that rank has initialised them
        sink_sum.zero();
        allreduce(sink_sum.p, 1, Dt::F64, Red::Sum);
    }
    // the per-iteration sum of the sink mass, which is also the barrier: peer mailboxes (GX_PR_MAIL=0: NCCL all-reduce)
    PeerMail *mail = nullptr;
    if (multi() && c.nranks <= MAX_PEERS) {
        const char *me = getenv("GX_PR_MAIL");
        if (!(me && me[0] == '0')) mail = &context_mail();
    }
    for (int it = 0; it < iters; it++) {
        const double *sink_in = s_in;
        unsigned n_sink_in = nparts;
        if (multi()) {
            // the sink mass is spread over the ranks: fold the local partials, all-reduce the scalar
            if (mail && mail->ok) {
                GX_LAUNCH(k_pr_tele_mail, 1, 256, 0, s_in, nparts, sink_sum.p, mail->table, ++mail->seq);
            } else {
                GX_LAUNCH(k_pr_tele, 1, 256, 0, s_in, nparts, sink_sum.p);
                allreduce(sink_sum.p, 1, Dt::F64, Red::Sum);
            }
            sink_in = sink_sum.p;
            n_sink_in = 1;
        }
        GX_CUDA(cudaMemsetAsync(s_out, 0, nparts * sizeof(double), c.stream));
        double *rank = (it == iters - 1) ? g->res_f64.p : nullptr;
        PtArgs a;
        a.col = pt.col.p; a.ne_ptr = pt.ne_ptr.p; a.ne_rows = pt.ne_rows.p; a.tile_k0 = pt.tile_k0.p;
        a.mask = (const uint8_t *)pt.mask.p; a.slot_k = pt.slot_k.p; a.d_k = d_k.p;
        double *w_old = wv[cur], *w_new = wv[cur ^ 1];
        const WOut *wout = peer_tab.p + (cur ^ 1);
        a.w = w_old; a.w_cold = w_old - hot; a.hps = pt.hps ? pt.hps : 1; a.wps = (uint32_t)pt.wps;
        for (int r = 0; r < MAX_PEERS; r++)
            a.hot_base[r] = r < c.nranks ? (uint32_t)(pt.compact ? pt.lbase[r] : (uint64_t)r * pt.wps) : 0u;
        a.warm_end = hot + (uint32_t)(pt.wps * (uint64_t)c.nranks); a.sink_in = sink_in; a.n_sink_in = n_sink_in; a.tele_out = tele.p;
        a.w_new.n = 1;
        a.w_new.p[0] = w_new;
        if (fused) {
            a.w_new.n = c.nranks;
            for (int r = 0; r < c.nranks; r++) a.w_new.p[r] = (double *)pt.wbuf[cur ^ 1].peer[r];
        }
        a.rank = rank;
        a.head_part = head_part.p; a.tail_part = tail_part.p; a.sink_out = s_out;
        a.K = pt.K; a.M = pt.M; a.n_tiles = pt.n_tiles; a.hot = hot; a.sc = sc;
        a.empty_v = pt.fin_v.p + pt.n_span; a.empty_slot = pt.fin_slot.p + pt.n_span; a.empty_inv = d_fin.p + pt.n_span;
        a.n_empty = pt.n_empty;
        {
            const bool prof__ = profiling();
            if (prof__) prof_begin("k_pr_tiles");
            tiles_fn[fused ? 1 : 0]<<<g_tiles, tile_threads, smem, c.stream>>>(a);
            if (prof__) prof_end();
            count_launch();
            GX_CUDA(cudaGetLastError());
        }
        if (g_fin && fused)
            GX_LAUNCH(k_pr_tile_fin<true>, g_fin, 256, 0, pt.fin_v.p, pt.fin_slot.p, pt.fin_t0.p, pt.fin_nt.p, d_fin.p, pt.n_span,
                      pt.n_span, head_part.p, tail_part.p, tele.p, w_new, wout, rank, s_out + g_tiles);
        else if (g_fin)
            GX_LAUNCH(k_pr_tile_fin<false>, g_fin, 256, 0, pt.fin_v.p, pt.fin_slot.p, pt.fin_t0.p, pt.fin_nt.p, d_fin.p, pt.n_span,
                      pt.n_span, head_part.p, tail_part.p, tele.p, w_new, wout, rank, s_out + g_tiles);
        // the ranks exchange their segments of the new w (their slices of r after the last iteration);
        // a rank's rows are one contiguous segment of the index space w lives in, with the row block's bounds
        if (it + 1 < iters) {
            if (cpush) {
                const unsigned np = (unsigned)plists[cur ^ 1].npeer;
                GX_LAUNCH(k_pr_push_lists, np * ((2 * (unsigned)c.num_sms + np - 1) / np), 256, 0, w_new + pt.lbase[c.rank], plists[cur ^ 1]);
            } else if (push) {
                const unsigned np = (unsigned)c.nranks - 1;
                GX_LAUNCH(k_pr_push_segment, np * ((2 * (unsigned)c.num_sms + np - 1) / np), 256, 0, w_new, wout,
                          (uint64_t)c.rank * pt.wps, pt.wps, (uint64_t)c.nranks * pt.wps + (uint64_t)c.rank * pt.tps, pt.tps);
            } else if (!fused) {
                allgather_equal(w_new, Dt::F64, pt.wps);
                if (pt.tps) allgather_equal(w_new + (uint64_t)c.nranks * pt.wps, Dt::F64, pt.tps);
            }
        }
        else allgatherv(rank, Dt::F64, plan.part);
        cur ^= 1;
        double *t = s_in; s_in = s_out; s_out = t;
    }
}

} // namespace gx

using namespace gx;

void gx_pr_cache_free(void *p) { delete (PrTiles *)p; }

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
        {
            PhaseTimer tb(&c.timing.build_ms);
            ensure_plan(g->in_adj(), n);
            if (!g->pr_cache) g->pr_cache = build_pr_tiles(g);
        }
        g->res_f64.alloc(n);
        pagerank_tiles(g, damping, iters);
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
