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
// Rows are binned by entry count d once per graph; a bin's table is sized for its longest row, so that the shared
// memory of a CTA -- hence the number of warps an SM keeps in flight -- follows the bin (every row is a chain of
// dependent misses, offsets -> column ids -> labels, that only other rows can overlap):
//   T4..T32  d <= 4 / 8 / 16 / 32   G = 4/8/16/32 lanes per row, one label per lane, counted with a
//                  single __match_any_sync on (group, label) -- registers only, no table
//   M1 M2    d <= 128 / 512         warp per row, 256- / 1024-slot open-addressing table per warp in shared memory
//   C1 C2 C3 d <= 1024 / 2048 / 4096  CTA per row, 2048- / 4096- / 8192-slot table in shared memory
//   H        d  > 4096   (hubs) 4096-entry pieces: a 512-thread CTA aggregates its piece in an 8192-slot smem table,
//                  compacts the distinct labels, prefetches their slots of the row's global table (2d slots of
//                  {label + 1, epoch | count}) and then claims / adds; a per-row finalize
// All table kernels keep CDLP_U (4) column ids and then 4 labels in flight per lane.
// Shared-memory atomics cost ~2 cycles per lane, so equal labels met by one warp instruction are merged first
// (__match_any_sync, leader lane adds the count): once labels converge a row's neighbours share a few labels and
// most atomics disappear.  The arg-max key is (count << 32) | ~label, so max() picks the highest count and, among
// equals, the smallest label -- bit-exact with the sorted-run scan.  No table is ever scanned: the add that
// completes a label's count returns the full count, so the largest key any add of a row has returned is the row's
// arg-max; shared-memory tables are cleared with vector stores, the hub tables are never cleared -- a slot tagged
// with another iteration's epoch is free.
// Active rows.  L_{t+1}(v) depends only on L_t of v's neighbours, so a row none of whose neighbours
// changed in the last iteration keeps its label.  After each iteration k_cdlp_stat adds up the
// entries of the rows that changed; once that is below 1/8 of all entries the changed rows mark their
// neighbours in a byte map (k_cdlp_mark_rows, k_cdlp_mark_pieces for the hubs), the marked rows of every bin and
// the pieces of the marked hubs are compacted (k_cdlp_compact_active) and the next iteration recomputes only those
// -- on RMAT graphs iterations 5..10 touch ~5-15 % of the entries.  The labels are the same bit for bit: skipped
// rows would have recomputed the value they already hold.
// Algorithmic bytes per iteration: 4 m' + 8(n+1) [x2 directed] + 4n + 4n.
#include <cstdlib>
#include <vector>

#include "graph.cuh"

namespace gx {

constexpr uint32_t EMPTY = 0xFFFFFFFFu;
constexpr int CDLP_BINS = 10; // T4 T8 T16 T32 | M1 M2 | C1 C2 C3 | H
constexpr uint32_t CDLP_M_MAX = 512, CDLP_C_MAX = 4096;
// The table of a bin is sized for the bin's longest row: a CTA's shared memory, hence the warps an SM keeps in flight,
// follows the bin -- every row is a chain of dependent misses (offsets -> column ids -> labels) that only other rows
// can overlap.  One 1024-slot table per warp / 8192-slot table per CTA for all of M / C kept 24 warps per SM.
constexpr uint32_t CDLP_M1_MAX = 128, CDLP_C1_MAX = 1024, CDLP_C2_MAX = 2048;
constexpr uint32_t CDLP_PIECE = 4096;  // hub entries per CTA
constexpr uint32_t CDLP_CT = 8192;     // slots of the largest CTA-wide shared-memory table (64 KB): bin C3
constexpr uint32_t CDLP_HCT = 2 * CDLP_PIECE; // slots of a hub piece's shared-memory table
constexpr uint32_t CDLP_EPOCHS = 16, CDLP_CNT_MASK = 0x0FFFFFFFu; // 4-bit epoch tag above a 28-bit count

struct CdlpPlan {
    bool built = false;
    bool first_closed_form = false; // undirected, no repeated entries: iteration 1 is "smallest neighbour"
    Partition part; // row blocks balanced by entries (out + in)
    uint64_t nb[CDLP_BINS] = {0}; // rows per bin
    uint64_t nL = 0, n_ins = 0, slots = 0; // L = hub rows (bin H)
    uint32_t epoch = 0;            // of the last iteration that used the hub tables (1 .. CDLP_EPOCHS - 1; 0 = just cleared)
    DevBuf<uint32_t> list[CDLP_BINS - 1], listL;
    DevBuf<uint64_t> tab_off;      // nL + 1: first slot of each L row's table
    DevBuf<uint32_t> ins_row;      // insert chunks: index into listL
    DevBuf<uint8_t> ins_side;      // 0 = out adjacency, 1 = in adjacency
    DevBuf<uint64_t> ins_begin;    // first entry of the chunk
    DevBuf<uint2> gtab;            // global tables: slot = {label + 1, epoch << 28 | count}, one 8-byte word.  A slot whose
                                   // epoch is not the running iteration's is free: nothing is cleared between iterations
                                   // (a memset of 16 bytes per hub entry and, in sparse iterations, a scan of the active
                                   // rows' tables before); one memset every CDLP_EPOCHS - 1 iterations when the tag wraps
    DevBuf<unsigned long long> best; // nL arg-max accumulators
    // sparse iterations: the active rows of every bin and the pieces of the active hub rows, compacted per iteration
    DevBuf<uint32_t> alist[CDLP_BINS - 1], apieces;
    DevBuf<unsigned long long> acount; // CDLP_BINS counters (the last one: pieces)
};

__device__ __forceinline__ uint32_t hash32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85ebca6bu; h ^= h >> 13; h *= 0xc2b2ae35u; h ^= h >> 16;
    return h;
}

struct CdlpLists { uint32_t *l[CDLP_BINS]; };

__device__ __forceinline__ int cdlp_bin_of(uint64_t d)
{
    return d <= 4 ? 0 : d <= 8 ? 1 : d <= 16 ? 2 : d <= 32 ? 3 : d <= CDLP_M1_MAX ? 4 : d <= CDLP_M_MAX ? 5
         : d <= CDLP_C1_MAX ? 6 : d <= CDLP_C2_MAX ? 7 : d <= CDLP_C_MAX ? 8 : 9;
}

// (one atomic per bin per 256 rows: a per-row atomic on seven counters ran at 20 GB/s, 1 ms per pass at RMAT-22;
// the order inside a list is arbitrary here, the lists are sorted afterwards)
__global__ void __launch_bounds__(256)
k_cdlp_bin(const uint64_t *__restrict__ rp0, const uint64_t *__restrict__ rp1, uint64_t v0, uint64_t v1,
           CdlpLists lists, unsigned long long *__restrict__ counts, int write)
{
    __shared__ unsigned s_cnt[CDLP_BINS];
    __shared__ unsigned long long s_base[CDLP_BINS];
    const uint64_t span = v1 - v0, nround = (span + 255) & ~255ull;
    for (uint64_t base = (uint64_t)blockIdx.x * 256; base < nround; base += (uint64_t)gridDim.x * 256) {
        if (threadIdx.x < CDLP_BINS) s_cnt[threadIdx.x] = 0;
        __syncthreads();
        const uint64_t v = v0 + base + threadIdx.x;
        uint64_t d = 0;
        if (v < v1) {
            d = rp0[v + 1] - rp0[v];
            if (rp1) d += rp1[v + 1] - rp1[v];
        }
        const int b = cdlp_bin_of(d);
        unsigned pos = 0;
        if (d) pos = atomicAdd(&s_cnt[b], 1u);
        __syncthreads();
        if (threadIdx.x < CDLP_BINS && s_cnt[threadIdx.x]) s_base[threadIdx.x] = atomicAdd(&counts[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
        __syncthreads();
        if (d && write) lists.l[b][s_base[b] + pos] = (uint32_t)v;
        __syncthreads();
    }
}

__global__ void k_cdlp_init(uint32_t *__restrict__ a, uint32_t *__restrict__ b, uint64_t n)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) { a[v] = (uint32_t)v; b[v] = (uint32_t)v; }
}

// bins T4..T32: G lanes per row, one neighbour label per lane, no table.  K32 (n <= 2^28): the (group, label) match key
// fits 32 bits -- MATCH.ANY on a 64-bit key is the dearer instruction, and these kernels are all match + row bookkeeping.
template <int G, bool K32>
__global__ void __launch_bounds__(256)
k_cdlp_tiny(const uint32_t *__restrict__ list, uint64_t count, const unsigned long long *__restrict__ count_dev,
            const uint64_t *__restrict__ rp0,
            const uint32_t *__restrict__ col0, const uint64_t *__restrict__ rp1, const uint32_t *__restrict__ col1,
            const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt, const uint8_t *__restrict__ active,
            int *__restrict__ changed)
{
    const unsigned sub = threadIdx.x & (G - 1);
    uint64_t gi = ((uint64_t)blockIdx.x * 256 + threadIdx.x) / G;
    if (count_dev) count = *count_dev; // sparse iterations: `list` holds the active rows only, counted on the device
    const uint64_t ngrp = ((uint64_t)gridDim.x * 256) / G;
    const uint64_t trips = (count + ngrp - 1) / ngrp; // same for every lane: the warp-wide intrinsics stay convergent
    bool ch = false;
    for (uint64_t t = 0; t < trips; t++, gi += ngrp) {
        const bool live = gi < count;
        uint32_t v = 0, lab = 0;
        bool have = false, act = false;
        if (live) {
            v = list[gi];
            act = !active || active[v];
        }
        if (act) {
            const uint64_t a0 = rp0[v], d0 = rp0[v + 1] - a0;
            uint64_t a1 = 0, d = d0;
            if (rp1) { a1 = rp1[v]; d += rp1[v + 1] - a1; }
            have = sub < d;
            if (have) lab = cur[sub < d0 ? ld_stream(col0 + a0 + sub) : ld_stream(col1 + a1 + (sub - d0))];
        }
        if (!__any_sync(FULL, act)) { // late iterations: most warps hold no active row
            if (live && sub == 0) nxt[v] = cur[v];
            continue;
        }
        // lanes of one row with equal labels match each other; rows (groups) and idle lanes never do
        unsigned same;
        if (K32) {
            // group in bits 29-31, bit 28 clear; idle lanes: bit 28 set + the lane
            same = __match_any_sync(FULL, have ? (((lane_id() / G) << 29) | lab) : (0x10000000u | lane_id()));
        } else {
            const unsigned long long key = have ? (((unsigned long long)(lane_id() / G) << 32) | lab)
                                                : ((unsigned long long)(0x100u + lane_id()) << 32);
            same = __match_any_sync(FULL, key);
        }
        unsigned long long best = have ? (((unsigned long long)__popc(same) << 32) | (uint32_t)~lab) : 0ull;
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1) {
            const unsigned long long x = __shfl_xor_sync(FULL, best, o);
            best = x > best ? x : best;
        }
        if (live && sub == 0) {
            const uint32_t old = cur[v];
            const uint32_t nl = act ? ~(uint32_t)best : old; // a row without changed neighbours keeps its label
            nxt[v] = nl;
            if (nl != old) ch = true;
        }
    }
    if (ch) *changed = 1;
}

// add `n` occurrences of `lab` to an open-addressing table (power-of-two size) in shared memory; returns the label's
// count after the add
__device__ __forceinline__ uint32_t smem_insert(uint32_t *key, uint32_t *cnt, uint32_t mask, uint32_t lab, uint32_t n)
{
    uint32_t s = hash32(lab) & mask;
    for (;;) {
        uint32_t seen = ((volatile uint32_t *)key)[s];
        if (seen == EMPTY) seen = atomicCAS(&key[s], EMPTY, lab);
        if (seen == EMPTY || seen == lab) return atomicAdd(&cnt[s], n) + n;
        s = (s + 1) & mask;
    }
}

// lanes of a warp that hold the same label elect one lane to insert it with the multiplicity.  Returns the arg-max
// key (count << 32 | ~label) of the count that add produced, 0 on the other lanes: the add that completes a label's
// count returns the full count, so the largest key any add of a row has returned is the row's arg-max and the table
// never has to be scanned -- it is cleared with vector stores (the scan of 2-4 slots per entry cost as much as the
// inserts themselves).
__device__ __forceinline__ unsigned long long warp_insert(uint32_t *key, uint32_t *cnt, uint32_t mask, uint32_t lab, bool valid)
{
    const unsigned same = __match_any_sync(FULL, valid ? lab : EMPTY);
    if (valid && lane_id() == (unsigned)(__ffs(same) - 1))
        return ((unsigned long long)smem_insert(key, cnt, mask, lab, __popc(same)) << 32) | (uint32_t)~lab;
    return 0ull;
}

// slots [0, teff) of a table back to empty: `nthreads` threads (this one is `t`), 4 slots per store; teff is a multiple of 4
__device__ __forceinline__ void smem_table_clear(uint32_t *key, uint32_t *cnt, uint32_t teff, uint32_t t, uint32_t nthreads)
{
    for (uint32_t s = 4 * t; s < teff; s += 4 * nthreads) {
        *(uint4 *)(key + s) = make_uint4(EMPTY, EMPTY, EMPTY, EMPTY);
        *(uint4 *)(cnt + s) = make_uint4(0u, 0u, 0u, 0u);
    }
}

// Labels of the entries k0, k0 + step, ... (U of them) of a row whose first d0 entries come from (col0 + a0) and the
// rest from (col1 + a1): the U column ids and then the U labels are in flight together -- one trip at a time left a
// warp waiting on two dependent misses (~1700 cycles) per 32 entries, which, not the tables, bounded these kernels.
constexpr int CDLP_U = 4;
__device__ __forceinline__ void cdlp_labels(const uint32_t *__restrict__ col0, uint64_t a0, uint64_t d0,
                                            const uint32_t *__restrict__ col1, uint64_t a1, uint64_t d, uint64_t k0, uint64_t step,
                                            const uint32_t *__restrict__ cur, uint32_t (&lab)[CDLP_U])
{
    uint32_t c[CDLP_U];
#pragma unroll
    for (int j = 0; j < CDLP_U; j++) {
        const uint64_t k = k0 + (uint64_t)j * step;
        c[j] = k < d ? (k < d0 ? ld_stream(col0 + a0 + k) : ld_stream(col1 + a1 + (k - d0))) : EMPTY;
    }
#pragma unroll
    for (int j = 0; j < CDLP_U; j++) lab[j] = c[j] != EMPTY ? cur[c[j]] : EMPTY;
}

// bins M1 / M2: one warp per row, a WT-slot table per warp in shared memory (WT >= 2 x the bin's longest row)
template <uint32_t CDLP_WT>
__global__ void __launch_bounds__(256)
k_cdlp_warp_rows(const uint32_t *__restrict__ list, uint64_t count, const unsigned long long *__restrict__ count_dev,
                 const uint64_t *__restrict__ rp0,
                 const uint32_t *__restrict__ col0, const uint64_t *__restrict__ rp1, const uint32_t *__restrict__ col1,
                 const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt, const uint8_t *__restrict__ active,
                 int *__restrict__ changed)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    for (uint32_t i = threadIdx.x; i < 8 * CDLP_WT * 2; i += 256) s_tab[i] = (i < 8 * CDLP_WT) ? EMPTY : 0u;
    __syncthreads();
    const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
    uint32_t *key = s_tab + wib * CDLP_WT;
    uint32_t *cnt = s_tab + 8 * CDLP_WT + wib * CDLP_WT;
    if (count_dev) count = *count_dev;
    const uint64_t nwarp = (uint64_t)gridDim.x * 8;
    bool ch = false;
    for (uint64_t r = (uint64_t)blockIdx.x * 8 + wib; r < count; r += nwarp) {
        const uint32_t v = list[r];
        if (active && !active[v]) { if (lane == 0) nxt[v] = cur[v]; continue; }
        const uint64_t a0 = rp0[v], d0 = rp0[v + 1] - a0;
        uint64_t a1 = 0, d = d0;
        if (rp1) { a1 = rp1[v]; d += rp1[v + 1] - a1; }
        uint32_t teff = 64;
        while (teff < 2 * d && teff < CDLP_WT) teff <<= 1;
        const uint32_t mask = teff - 1;
        unsigned long long best = 0;
        for (uint64_t base = 0; base < d; base += 32 * CDLP_U) {
            uint32_t lab[CDLP_U];
            cdlp_labels(col0, a0, d0, col1, a1, d, base + lane, 32, cur, lab);
#pragma unroll
            for (int j = 0; j < CDLP_U; j++)
                if (base + 32u * j < d) {
                    const unsigned long long kk = warp_insert(key, cnt, mask, lab[j], lab[j] != EMPTY);
                    best = kk > best ? kk : best;
                }
        }
        __syncwarp();
        smem_table_clear(key, cnt, teff, lane, 32);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long x = __shfl_xor_sync(FULL, best, o);
            best = x > best ? x : best;
        }
        if (lane == 0) {
            const uint32_t nl = ~(uint32_t)best;
            nxt[v] = nl;
            if (nl != cur[v]) ch = true;
        }
        __syncwarp();
    }
    if (ch) *changed = 1;
}

// bins C1 / C2 / C3: one CTA per row, a CT-slot table in shared memory (CT >= 2 x the bin's longest row), NT threads
template <uint32_t CT, uint32_t NT>
__global__ void __launch_bounds__(NT)
k_cdlp_cta_rows(const uint32_t *__restrict__ list, uint64_t count, const unsigned long long *__restrict__ count_dev,
                const uint64_t *__restrict__ rp0,
                const uint32_t *__restrict__ col0, const uint64_t *__restrict__ rp1, const uint32_t *__restrict__ col1,
                const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt, const uint8_t *__restrict__ active,
                int *__restrict__ changed)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    uint32_t *key = s_tab, *cnt = s_tab + CT;
    __shared__ unsigned long long s_best[NT / 32];
    for (uint32_t i = threadIdx.x; i < CT; i += NT) { key[i] = EMPTY; cnt[i] = 0; }
    __syncthreads();
    if (count_dev) count = *count_dev;
    for (uint64_t r = blockIdx.x; r < count; r += gridDim.x) {
        const uint32_t v = list[r];
        if (active && !active[v]) { if (threadIdx.x == 0) nxt[v] = cur[v]; continue; } // uniform per CTA
        const uint64_t a0 = rp0[v], d0 = rp0[v + 1] - a0;
        uint64_t a1 = 0, d = d0;
        if (rp1) { a1 = rp1[v]; d += rp1[v + 1] - a1; }
        uint32_t teff = 1024;
        while (teff < 2 * d && teff < CT) teff <<= 1;
        const uint32_t mask = teff - 1;
        unsigned long long best = 0;
        for (uint64_t base = 0; base < d; base += NT * CDLP_U) {
            uint32_t lab[CDLP_U];
            cdlp_labels(col0, a0, d0, col1, a1, d, base + threadIdx.x, NT, cur, lab);
#pragma unroll
            for (int j = 0; j < CDLP_U; j++)
                if (base + NT * j + (threadIdx.x & ~31u) < d) { // warp-uniform
                    const unsigned long long kk = warp_insert(key, cnt, mask, lab[j], lab[j] != EMPTY);
                    best = kk > best ? kk : best;
                }
        }
        __syncthreads();
        smem_table_clear(key, cnt, teff, threadIdx.x, NT);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long x = __shfl_xor_sync(FULL, best, o);
            best = x > best ? x : best;
        }
        if (lane_id() == 0) s_best[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 1; i < (int)(NT / 32); i++) best = s_best[i] > best ? s_best[i] : best;
            const uint32_t nl = ~(uint32_t)best;
            nxt[v] = nl;
            if (nl != cur[v]) *changed = 1;
        }
        __syncthreads();
    }
}

// hub rows, pass 1: one CTA per CDLP_PIECE entries.  The piece is aggregated in shared memory
// first; only its distinct labels go to the row's global table (one atomicAdd of the count each).
constexpr uint32_t CDLP_HT = 512; // threads of a hub-piece CTA (3 CTAs of 64 KB per SM: 48 warps)
__global__ void __launch_bounds__(CDLP_HT)
k_cdlp_big_insert(const uint32_t *__restrict__ listL, const uint32_t *__restrict__ ins_row,
                  const uint8_t *__restrict__ ins_side, const uint64_t *__restrict__ ins_begin,
                  const uint64_t *__restrict__ tab_off, const uint64_t *__restrict__ rp0, const uint32_t *__restrict__ col0,
                  const uint64_t *__restrict__ rp1, const uint32_t *__restrict__ col1, const uint32_t *__restrict__ cur,
                  const uint32_t *__restrict__ piece_ids, const unsigned long long *__restrict__ npieces_dev, uint32_t npieces,
                  uint2 *__restrict__ gtab, uint32_t epoch, unsigned long long *__restrict__ best_out)
{
    extern __shared__ __align__(16) uint32_t s_tab[];
    uint32_t *key = s_tab, *cnt = s_tab + CDLP_HCT;
    __shared__ unsigned long long s_best[CDLP_HT / 32];
    // dense iterations: one CTA per piece (piece_ids == NULL); sparse ones: the pieces of the active hub rows, listed and
    // counted on the device, drawn by a persistent grid
    if (npieces_dev) npieces = (uint32_t)*npieces_dev;
    for (uint32_t ci = blockIdx.x; ci < npieces; ci += gridDim.x) {
        const uint32_t c = piece_ids ? piece_ids[ci] : ci;
        const uint32_t li = ins_row[c];
        const uint32_t v = listL[li];
        for (uint32_t i = threadIdx.x; i < CDLP_HCT; i += CDLP_HT) { key[i] = EMPTY; cnt[i] = 0; }
        __syncthreads();
        const bool side = ins_side[c] != 0;
        const uint64_t *rp = side ? rp1 : rp0;
        const uint32_t *col = side ? col1 : col0;
        const uint64_t b0 = ins_begin[c];
        const uint64_t row_end = rp[v + 1];
        const uint64_t e_end = (b0 + CDLP_PIECE < row_end) ? b0 + CDLP_PIECE : row_end;
        const uint64_t t0 = tab_off[li];
        const uint64_t tsize = tab_off[li + 1] - t0;
        for (uint64_t base = b0; base < e_end; base += CDLP_HT * CDLP_U) {
            uint32_t lab[CDLP_U];
            cdlp_labels(col, 0, e_end, nullptr, 0, e_end, base + threadIdx.x, CDLP_HT, cur, lab);
#pragma unroll
            for (int j = 0; j < CDLP_U; j++)
                if (base + CDLP_HT * j + (threadIdx.x & ~31u) < e_end) warp_insert(key, cnt, CDLP_HCT - 1, lab[j], lab[j] != EMPTY);
        }
        __syncthreads();
        // The add that completes a label's count returns that count, so the largest (count, ~label) any add of
        // the row has seen is the row's arg-max: one atomicMax per piece, no scan of the row's table.
        // Each warp owns an equal share of the table.  It first compacts its non-empty slots to the front of its share (in place: the
        // write position never passes the read position), so that every lane of every later trip has work -- with a few
        // hundred distinct labels per piece most trips of the strided walk waited two dependent L2-miss atomics for one or two
        // lanes.  Then the target lines of all its (label, count) pairs are prefetched into L2, and only then come the
        // atomics: CDLP_U claims in flight per lane, then their adds (same 8-byte slot, now an L2 hit).
        constexpr uint32_t SHARE = CDLP_HCT / (CDLP_HT / 32);
        const unsigned lane = lane_id();
        uint32_t *wkey = key + (threadIdx.x >> 5) * SHARE, *wcnt = cnt + (threadIdx.x >> 5) * SHARE;
        uint32_t nw = 0;
        for (uint32_t i = lane; i < SHARE; i += 32) {
            const uint32_t cc = wcnt[i], lab = wkey[i];
            const unsigned m = __ballot_sync(FULL, cc != 0);
            __syncwarp();
            if (cc) { const uint32_t d = nw + __popc(m & ((1u << lane) - 1u)); wkey[d] = lab; wcnt[d] = cc; }
            nw += __popc(m);
            __syncwarp();
        }
        for (uint32_t i = lane; i < nw; i += 32) {
            const uint64_t sl = ((uint64_t)hash32(wkey[i]) * tsize) >> 32;
            asm volatile("prefetch.global.L2 [%0];" ::"l"(gtab + t0 + sl));
        }
        // Claiming a slot: a slot tagged with this iteration's epoch is live (ours if it holds the label, else probe on);
        // any other tag means free, and it is claimed by a 64-bit compare-and-swap from the value just seen -- losing that
        // race makes the slot live, so it is simply looked at again.
        const uint32_t tag = epoch << 28;
        unsigned long long best = 0;
        for (uint32_t i0 = lane; i0 < nw; i0 += 32 * CDLP_U) {
            uint32_t cc[CDLP_U], lab[CDLP_U];
            uint64_t s[CDLP_U];
            uint2 seen[CDLP_U];
            bool claimed[CDLP_U];
#pragma unroll
            for (int j = 0; j < CDLP_U; j++) {
                claimed[j] = false;
                const bool have = i0 + 32 * j < nw;
                cc[j] = have ? wcnt[i0 + 32 * j] : 0u;
                lab[j] = have ? wkey[i0 + 32 * j] : EMPTY;
                s[j] = ((uint64_t)hash32(lab[j]) * tsize) >> 32;
            }
#pragma unroll
            for (int j = 0; j < CDLP_U; j++) seen[j] = cc[j] ? __ldcg(gtab + t0 + s[j]) : make_uint2(0u, 0u);
#pragma unroll
            for (int j = 0; j < CDLP_U; j++) {
                if (!cc[j]) continue;
                for (;;) {
                    if ((seen[j].y & ~CDLP_CNT_MASK) == tag) {
                        if (seen[j].x == lab[j] + 1u) break;
                        s[j] = (s[j] + 1 == tsize) ? 0 : s[j] + 1;
                        seen[j] = __ldcg(gtab + t0 + s[j]);
                        continue;
                    }
                    const unsigned long long was = ((unsigned long long)seen[j].y << 32) | seen[j].x;
                    // (the claim carries the piece's count: claiming and the first add are one operation)
                    const unsigned long long want = ((unsigned long long)(tag | cc[j]) << 32) | (lab[j] + 1u);
                    const unsigned long long got = atomicCAS((unsigned long long *)(gtab + t0 + s[j]), was, want);
                    if (got == was) { claimed[j] = true; break; }
                    seen[j] = make_uint2((uint32_t)got, (uint32_t)(got >> 32));
                }
            }
            unsigned long long now[CDLP_U];
#pragma unroll
            for (int j = 0; j < CDLP_U; j++)
                now[j] = claimed[j] ? cc[j]
                       : cc[j]      ? (unsigned long long)((atomicAdd(&gtab[t0 + s[j]].y, cc[j]) + cc[j]) & CDLP_CNT_MASK) : 0ull;
#pragma unroll
            for (int j = 0; j < CDLP_U; j++) {
                const unsigned long long kk = (now[j] << 32) | (uint32_t)~lab[j];
                if (cc[j] && kk > best) best = kk;
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long x = __shfl_xor_sync(FULL, best, o);
            best = x > best ? x : best;
        }
        if (lane_id() == 0) s_best[threadIdx.x >> 5] = best;
        __syncthreads();
        if (threadIdx.x == 0) {
#pragma unroll
            for (int i = 1; i < (int)(CDLP_HT / 32); i++) best = s_best[i] > best ? s_best[i] : best;
            if (best) atomicMax(&best_out[li], best);
        }
        __syncthreads(); // the table and s_best are reused by the next piece
    }
}

__global__ void k_cdlp_big_final(const uint32_t *__restrict__ listL, uint64_t nL, unsigned long long *__restrict__ best,
                                 const uint32_t *__restrict__ cur, uint32_t *__restrict__ nxt,
                                 const uint8_t *__restrict__ active, int *__restrict__ changed)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nL) return;
    const uint32_t v = listL[i];
    if (active && !active[v]) { nxt[v] = cur[v]; return; }
    const uint32_t nl = ~(uint32_t)best[i];
    best[i] = 0;
    nxt[v] = nl;
    if (nl != cur[v]) *changed = 1;
}

// ---- first iteration in closed form --------------------------------------------------------------
// L0(v) = v, so in iteration 1 every stored entry of a row carries a different label unless the row
// repeats an entry: all counts are 1 and the smallest label -- the first entry of the sorted row --
// wins.  Holds for undirected graphs (one adjacency) without repeated entries, which k_cdlp_has_dup
// checks once per graph: rows are sorted, so a repeat is an equal neighbour pair that does not
// straddle a row border.
__global__ void k_cdlp_has_dup(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t n, uint64_t m,
                               int *__restrict__ dup)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x + 1;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) {
        if (ld_stream(col + e) != ld_stream(col + e - 1)) continue;
        uint64_t lo = 0, hi = n; // is there a row r with rowptr[r] == e ?
        while (lo < hi) {
            const uint64_t mid = (lo + hi) >> 1;
            if (rowptr[mid] < e) lo = mid + 1; else hi = mid;
        }
        if (lo >= n || rowptr[lo] != e) *dup = 1;
    }
}

__global__ void k_cdlp_first(const uint64_t *__restrict__ rowptr, const uint32_t *__restrict__ col, uint64_t v0, uint64_t v1,
                             uint32_t *__restrict__ nxt, int *__restrict__ changed)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    bool ch = false;
    for (; v < v1; v += stride) {
        const uint64_t a = rowptr[v];
        const uint32_t nl = rowptr[v + 1] > a ? col[a] : (uint32_t)v;
        nxt[v] = nl;
        ch |= nl != (uint32_t)v;
    }
    if (ch) *changed = 1;
}

// ---- active rows -------------------------------------------------------------------------------
// Sparse iterations recompute 5 % of the rows: walking every bin's whole list for them (two dependent loads per row,
// list entry -> active byte) cost more than the rows themselves, and the hub kernel launched 35 K CTAs to find a few
// active pieces.  The active rows of every bin (blockIdx.y < CDLP_BINS - 1) and the pieces of the active hub rows
// (blockIdx.y == CDLP_BINS - 1) are compacted first; order inside a 256-row block is kept, blocks land in any order.
struct CdlpCompact {
    const uint32_t *src[CDLP_BINS]; // bin lists; the last entry is unused (pieces are numbered 0 .. n - 1)
    uint32_t *dst[CDLP_BINS];
    uint64_t n[CDLP_BINS];
};
__global__ void __launch_bounds__(256)
k_cdlp_compact_active(CdlpCompact a, const uint32_t *__restrict__ listL, const uint32_t *__restrict__ ins_row,
                      const uint8_t *__restrict__ active, unsigned long long *__restrict__ counts)
{
    __shared__ unsigned s_cnt[8];
    __shared__ unsigned long long s_base;
    const int b = blockIdx.y;
    const uint64_t n = a.n[b], nround = (n + 255) & ~255ull;
    const unsigned lane = lane_id(), wib = threadIdx.x >> 5;
    for (uint64_t base = (uint64_t)blockIdx.x * 256; base < nround; base += (uint64_t)gridDim.x * 256) {
        const uint64_t i = base + threadIdx.x;
        uint32_t item = 0;
        bool in = false;
        if (i < n) {
            item = b == CDLP_BINS - 1 ? (uint32_t)i : a.src[b][i];
            in = active[b == CDLP_BINS - 1 ? listL[ins_row[i]] : item] != 0;
        }
        const unsigned mask = __ballot_sync(FULL, in);
        if (lane == 0) s_cnt[wib] = __popc(mask);
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned tot = 0;
#pragma unroll
            for (int w = 0; w < 8; w++) { const unsigned c = s_cnt[w]; s_cnt[w] = tot; tot += c; }
            s_base = tot ? atomicAdd(&counts[b], (unsigned long long)tot) : 0ull;
        }
        __syncthreads();
        if (in) a.dst[b][s_base + s_cnt[wib] + __popc(mask & ((1u << lane) - 1u))] = item;
        __syncthreads();
    }
}

// entries of the rows [v0, v1) whose label just changed (what marking their neighbours would cost)
__global__ void __launch_bounds__(256)
k_cdlp_stat(const uint64_t *__restrict__ rp0, const uint64_t *__restrict__ rp1, uint64_t v0, uint64_t v1,
            const uint32_t *__restrict__ cur, const uint32_t *__restrict__ nxt, unsigned long long *__restrict__ degsum)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long s = 0;
    for (; v < v1; v += stride)
        if (cur[v] != nxt[v]) { s += rp0[v + 1] - rp0[v]; if (rp1) s += rp1[v + 1] - rp1[v]; }
    s = warp_sum(s);
    if (lane_id() == 0 && s) atomicAdd(degsum, s);
}

// rows of [v0, v1) up to CDLP_C_MAX entries (bins T..C): a warp looks at 32 consecutive vertices at a time (coalesced;
// an 8-lane group per vertex walked all n vertices at two dependent loads per trip: 140 us at RMAT-24 for a few
// thousand changed rows), then its four 8-lane groups take the changed ones four at a time and mark their neighbours
// (for directed graphs on both adjacencies: v counts towards its out- AND in-neighbours)
__global__ void __launch_bounds__(256)
k_cdlp_mark_rows(const uint64_t *__restrict__ rp0, const uint32_t *__restrict__ col0, const uint64_t *__restrict__ rp1,
                 const uint32_t *__restrict__ col1, uint64_t v0, uint64_t v1, const uint32_t *__restrict__ cur,
                 const uint32_t *__restrict__ nxt, uint8_t *__restrict__ active)
{
    const unsigned lane = lane_id(), grp = lane >> 3, sub = lane & 7u;
    const uint64_t wid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t nw = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    for (uint64_t b = v0 + wid * 32; b < v1; b += nw * 32) {
        const uint64_t mine = b + lane;
        unsigned m = __ballot_sync(FULL, mine < v1 && cur[mine] != nxt[mine]);
        while (m) {
            const unsigned bit = __fns(m, 0, grp + 1); // this group's changed vertex of the round, if any
#pragma unroll
            for (int k = 0; k < 4; k++) m &= m - 1;    // (m - 1 wraps harmlessly once m is 0: 0 & x = 0)
            if (bit == 0xFFFFFFFFu) continue;
            const uint64_t v = b + bit;
            const uint64_t a0 = rp0[v], e0 = rp0[v + 1];
            uint64_t a1 = 0, e1 = 0;
            if (rp1) { a1 = rp1[v]; e1 = rp1[v + 1]; }
            if ((e0 - a0) + (e1 - a1) > CDLP_C_MAX) continue; // hub: k_cdlp_mark_pieces
            for (uint64_t e = a0 + sub; e < e0; e += 8) active[ld_stream(col0 + e)] = 1;
            for (uint64_t e = a1 + sub; e < e1; e += 8) active[ld_stream(col1 + e)] = 1;
        }
    }
}

// hub rows: one CTA per CDLP_PIECE entries (the insert pieces of the plan)
__global__ void __launch_bounds__(256)
k_cdlp_mark_pieces(const uint32_t *__restrict__ listL, const uint32_t *__restrict__ ins_row, const uint8_t *__restrict__ ins_side,
                   const uint64_t *__restrict__ ins_begin, const uint64_t *__restrict__ rp0, const uint32_t *__restrict__ col0,
                   const uint64_t *__restrict__ rp1, const uint32_t *__restrict__ col1, const uint32_t *__restrict__ cur,
                   const uint32_t *__restrict__ nxt, uint8_t *__restrict__ active)
{
    const uint32_t c = blockIdx.x;
    const uint32_t v = listL[ins_row[c]];
    if (cur[v] == nxt[v]) return;
    const bool side = ins_side[c] != 0;
    const uint64_t *rp = side ? rp1 : rp0;
    const uint32_t *col = side ? col1 : col0;
    const uint64_t b0 = ins_begin[c];
    const uint64_t row_end = rp[v + 1];
    const uint64_t e_end = (b0 + CDLP_PIECE < row_end) ? b0 + CDLP_PIECE : row_end;
    for (uint64_t e = b0 + threadIdx.x; e < e_end; e += 256) active[ld_stream(col + e)] = 1;
}

// entries of the marked rows of [v0, v1): what the next iteration reads
__global__ void __launch_bounds__(256)
k_cdlp_active_entries(const uint64_t *__restrict__ rp0, const uint64_t *__restrict__ rp1, uint64_t v0, uint64_t v1,
                      const uint8_t *__restrict__ active, unsigned long long *__restrict__ total)
{
    uint64_t v = v0 + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    unsigned long long s = 0;
    for (; v < v1; v += stride)
        if (active[v]) { s += rp0[v + 1] - rp0[v]; if (rp1) s += rp1[v + 1] - rp1[v]; }
    s = warp_sum(s);
    if (lane_id() == 0 && s) atomicAdd(total, s);
}

__global__ void k_widen_u32_cdlp(const uint32_t *__restrict__ in, uint64_t n, uint64_t *__restrict__ out)
{
    uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; v < n; v += stride) out[v] = in[v];
}

// {first, last+1} entry of every hub row on both adjacencies
__global__ void k_cdlp_hub_extents(const uint32_t *__restrict__ listL, uint64_t nL, const uint64_t *__restrict__ rp0,
                                   const uint64_t *__restrict__ rp1, uint64_t *__restrict__ ext)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nL) return;
    const uint32_t v = listL[i];
    ext[4 * i] = rp0[v]; ext[4 * i + 1] = rp0[v + 1];
    ext[4 * i + 2] = rp1 ? rp1[v] : 0; ext[4 * i + 3] = rp1 ? rp1[v + 1] : 0;
}

static CdlpPlan *build_cdlp_plan(gx_graph *g)
{
    CdlpPlan *p = new CdlpPlan();
    const uint64_t n = g->n;
    const uint64_t *rp0 = g->out.rowptr.p;
    const uint64_t *rp1 = g->directed ? g->in.rowptr.p : nullptr;
    p->part = make_partition(rp0, rp1, n);
    const uint64_t v0 = p->part.lo, v1 = p->part.hi;
    DevBuf<unsigned long long> counts(CDLP_BINS);
    counts.zero();
    DevBuf<uint32_t> dummy(1);
    CdlpLists lists;
    for (int b = 0; b < CDLP_BINS; b++) lists.l[b] = dummy.p;
    GX_LAUNCH(k_cdlp_bin, grid_persistent(8), 256, 0, rp0, rp1, v0, v1, lists, counts.p, 0);
    unsigned long long h[CDLP_BINS];
    read_back(h, counts.p, sizeof(h));
    for (int b = 0; b < CDLP_BINS; b++) p->nb[b] = h[b];
    p->nL = h[CDLP_BINS - 1];
    for (int b = 0; b < CDLP_BINS - 1; b++) { p->list[b].alloc(h[b] ? h[b] : 1); lists.l[b] = p->list[b].p; p->alist[b].alloc(h[b] ? h[b] : 1); }
    p->acount.alloc(CDLP_BINS);
    p->listL.alloc(p->nL ? p->nL : 1);
    lists.l[CDLP_BINS - 1] = p->listL.p;
    counts.zero();
    GX_LAUNCH(k_cdlp_bin, grid_persistent(8), 256, 0, rp0, rp1, v0, v1, lists, counts.p, 1);
    // ascending vertex ids: neighbouring list entries then read neighbouring offsets, labels and entries
    for (int b = 0; b < CDLP_BINS - 1; b++) sort_keys32(p->list[b], h[b], bits_for(n));
    if (p->nL) {
        // the hub rows' extents come back from the device (a few thousand rows: the offsets of all n rows used to be
        // copied to the host for this, 19 ms of a 41 ms plan at RMAT-24); the piece lists are laid out on the host
        cudaStream_t s = ctx().stream;
        sort_keys32(p->listL, p->nL, bits_for(n));
        DevBuf<uint64_t> ext(4 * p->nL);
        GX_LAUNCH(k_cdlp_hub_extents, grid_for(p->nL, 256), 256, 0, p->listL.p, p->nL, rp0, rp1, ext.p);
        std::vector<uint64_t> hx(4 * p->nL);
        GX_CUDA(cudaMemcpyAsync(hx.data(), ext.p, 4 * p->nL * sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        GX_CUDA(cudaStreamSynchronize(s));
        std::vector<uint64_t> tab_off(p->nL + 1), ins_begin;
        std::vector<uint32_t> ins_row;
        std::vector<uint8_t> ins_side;
        uint64_t off = 0;
        for (uint64_t i = 0; i < p->nL; i++) {
            const uint64_t a0 = hx[4 * i], b0 = hx[4 * i + 1], a1 = hx[4 * i + 2], b1 = hx[4 * i + 3];
            uint64_t d = b0 - a0;
            for (uint64_t b = a0; b < b0; b += CDLP_PIECE) { ins_row.push_back((uint32_t)i); ins_side.push_back(0); ins_begin.push_back(b); }
            if (rp1) {
                d += b1 - a1;
                for (uint64_t b = a1; b < b1; b += CDLP_PIECE) { ins_row.push_back((uint32_t)i); ins_side.push_back(1); ins_begin.push_back(b); }
            }
            GX_REQUIRE(d <= CDLP_CNT_MASK, "CDLP: a row with 2^28 or more entries (the hub tables count in 28 bits)");
            tab_off[i] = off;
            off += 2 * d;
        }
        tab_off[p->nL] = off;
        p->slots = off;
        p->n_ins = ins_row.size();
        p->tab_off.alloc(p->nL + 1);
        p->ins_row.alloc(p->n_ins); p->ins_side.alloc(p->n_ins); p->ins_begin.alloc(p->n_ins);
        p->gtab.alloc(off); p->best.alloc(p->nL); p->apieces.alloc(p->n_ins);
        GX_CUDA(cudaMemcpyAsync(p->tab_off.p, tab_off.data(), (p->nL + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->ins_row.p, ins_row.data(), p->n_ins * sizeof(uint32_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->ins_side.p, ins_side.data(), p->n_ins * sizeof(uint8_t), cudaMemcpyHostToDevice, s));
        GX_CUDA(cudaMemcpyAsync(p->ins_begin.p, ins_begin.data(), p->n_ins * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
        p->gtab.zero();
        p->best.zero();
        GX_CUDA(cudaStreamSynchronize(s));
    }
    if (!g->directed && g->m) {
        DevBuf<int> dup(1);
        dup.zero();
        GX_LAUNCH(k_cdlp_has_dup, grid_persistent(8), 256, 0, rp0, g->out.col.p, n, g->m, dup.p);
        int hd = 0;
        read_back(&hd, dup.p, sizeof(hd));
        p->first_closed_form = !hd;
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
        ensure_in_full(g); // several GPUs: the row blocks are balanced over out- and in-entries together, not the in-edge blocks
        if (!g->cdlp_plan) {
            PhaseTimer tb(&c.timing.build_ms);
            g->cdlp_plan = build_cdlp_plan(g);
        }
        CdlpPlan &p = *(CdlpPlan *)g->cdlp_plan;
        const uint64_t *rp0 = g->out.rowptr.p, *rp1 = g->directed ? g->in.rowptr.p : nullptr;
        const uint32_t *col0 = g->out.col.p, *col1 = g->directed ? g->in.col.p : nullptr;
        constexpr size_t SMEM_C = (size_t)CDLP_CT * 8;
        GX_CUDA(cudaFuncSetAttribute(k_cdlp_warp_rows<1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 1024 * 8));
        GX_CUDA(cudaFuncSetAttribute(k_cdlp_cta_rows<CDLP_CT, 512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_C));
        GX_CUDA(cudaFuncSetAttribute(k_cdlp_big_insert, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(CDLP_HCT * 8)));
        g->res_u64.alloc(n);
        const uint64_t m_eff = g->directed ? 2 * g->m : g->m;
        const char *ae = getenv("GX_CDLP_ACTIVE"); // GX_CDLP_ACTIVE=0: recompute every row every iteration
        const bool use_active = !(ae && ae[0] == '0');
        const bool k32 = n <= (1ull << 28); // 32-bit match keys in the register kernels
        DevBuf<uint32_t> la(n), lb(n);
        DevBuf<uint8_t> active(use_active ? n : 0);
        DevBuf<unsigned long long> stats(3); // [0] changed flag, [1] entries of the changed rows, [2] entries of marked rows (all sparse iterations)
        int *changed = (int *)stats.p;
        uint32_t iters = 0, sparse_iters = 0;
        uint64_t inspected = 0;
        {
            PhaseTimer tk(&c.timing.kernel_ms);
            GX_LAUNCH(k_cdlp_init, grid_persistent(8), 256, 0, la.p, lb.p, n);
            stats.zero();
            uint32_t *cur = la.p, *nxt = lb.p;
            const uint8_t *act = nullptr; // nullptr: every row is recomputed
            const char *fe = getenv("GX_CDLP_FIRST"); // GX_CDLP_FIRST=0: iteration 1 through the general kernels
            const bool first_cf = p.first_closed_form && !(fe && fe[0] == '0');
            for (int it = 0; it < itermax; it++) {
                GX_CUDA(cudaMemsetAsync(stats.p, 0, 2 * sizeof(unsigned long long), c.stream));
                if (it == 0 && first_cf) {
                    GX_LAUNCH(k_cdlp_first, grid_persistent(8), 256, 0, rp0, col0, p.part.lo, p.part.hi, nxt, changed);
                    inspected += n;
                } else {
                // sparse iteration: only the marked rows are recomputed -- they are compacted per bin first (counts stay
                // on the device), every other row keeps its label by one copy of the label array
                const bool sparse = act != nullptr;
                const uint32_t *lst[CDLP_BINS - 1];
                const unsigned long long *cnt_dev[CDLP_BINS - 1];
                for (int b = 0; b < CDLP_BINS - 1; b++) { lst[b] = p.list[b].p; cnt_dev[b] = nullptr; }
                if (sparse) {
                    CdlpCompact ca{};
                    for (int b = 0; b < CDLP_BINS - 1; b++) {
                        ca.src[b] = p.list[b].p; ca.dst[b] = p.alist[b].p; ca.n[b] = p.nb[b];
                        lst[b] = p.alist[b].p; cnt_dev[b] = p.acount.p + b;
                    }
                    ca.src[CDLP_BINS - 1] = nullptr; ca.dst[CDLP_BINS - 1] = p.apieces.p; ca.n[CDLP_BINS - 1] = p.nL ? p.n_ins : 0;
                    p.acount.zero();
                    GX_LAUNCH(k_cdlp_compact_active, dim3(8 * (unsigned)c.num_sms, CDLP_BINS), 256, 0, ca, p.listL.p, p.ins_row.p, act, p.acount.p);
                    GX_CUDA(cudaMemcpyAsync(nxt, cur, n * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c.stream));
                }
                const uint8_t *no_map = nullptr; // the row kernels get compact lists, not the map
                if (p.nL) {
                    if (++p.epoch == CDLP_EPOCHS) { // the tag wraps: slots of 15 iterations ago would look live again
                        p.gtab.zero();
                        p.epoch = 1;
                    }
                    GX_LAUNCH(k_cdlp_big_insert, sparse ? grid_persistent(3) : (unsigned)p.n_ins, CDLP_HT, CDLP_HCT * 8, p.listL.p, p.ins_row.p,
                              p.ins_side.p, p.ins_begin.p, p.tab_off.p, rp0, col0, rp1, col1, cur, sparse ? p.apieces.p : nullptr,
                              sparse ? p.acount.p + (CDLP_BINS - 1) : nullptr, (uint32_t)p.n_ins, p.gtab.p, p.epoch, p.best.p);
                    GX_LAUNCH(k_cdlp_big_final, grid_for(p.nL, 256), 256, 0, p.listL.p, p.nL, p.best.p, cur, nxt, act, changed);
                }
#define CDLP_ROWS(kern, b, ctas, threads, smem)                                                                                     \
    do {                                                                                                                            \
        if (p.nb[b])                                                                                                                \
            GX_LAUNCH(kern, grid_persistent(ctas), threads, smem, lst[b], p.nb[b], cnt_dev[b], rp0, col0, rp1, col1, cur, nxt, no_map, changed); \
    } while (0)
                CDLP_ROWS((k_cdlp_cta_rows<8192, 512>), 8, 3, 512, 8192 * 8);
                CDLP_ROWS((k_cdlp_cta_rows<4096, 256>), 7, 6, 256, 4096 * 8);
                CDLP_ROWS((k_cdlp_cta_rows<2048, 256>), 6, 8, 256, 2048 * 8);
                CDLP_ROWS(k_cdlp_warp_rows<1024>, 5, 3, 256, 8 * 1024 * 8);
                CDLP_ROWS(k_cdlp_warp_rows<256>, 4, 8, 256, 8 * 256 * 8);
                if (k32) {
                    CDLP_ROWS((k_cdlp_tiny<32, true>), 3, 8, 256, 0);
                    CDLP_ROWS((k_cdlp_tiny<16, true>), 2, 8, 256, 0);
                    CDLP_ROWS((k_cdlp_tiny<8, true>), 1, 8, 256, 0);
                    CDLP_ROWS((k_cdlp_tiny<4, true>), 0, 8, 256, 0);
                } else {
                    CDLP_ROWS((k_cdlp_tiny<32, false>), 3, 8, 256, 0);
                    CDLP_ROWS((k_cdlp_tiny<16, false>), 2, 8, 256, 0);
                    CDLP_ROWS((k_cdlp_tiny<8, false>), 1, 8, 256, 0);
                    CDLP_ROWS((k_cdlp_tiny<4, false>), 0, 8, 256, 0);
                }
#undef CDLP_ROWS
                if (!act) inspected += m_eff; else sparse_iters++;
                }
                const bool more = it + 1 < itermax;
                if (use_active && more)
                    GX_LAUNCH(k_cdlp_stat, grid_persistent(8), 256, 0, rp0, rp1, p.part.lo, p.part.hi, cur, nxt, stats.p + 1);
                if (multi()) {
                    allgatherv(nxt, Dt::U32, p.part);       // owners publish their new labels
                    allreduce(stats.p, 2, Dt::U64, Red::Sum);
                }
                iters++;
                unsigned long long h[2] = {0, 0};
                read_back(h, stats.p, sizeof(h));
                if (!h[0]) { uint32_t *t = cur; cur = nxt; nxt = t; break; }
                if (use_active && more) {
                    if (h[1] <= m_eff / 8) {
                        // few rows changed: they mark their neighbours, only those are recomputed next
                        active.zero();
                        GX_LAUNCH(k_cdlp_mark_rows, grid_persistent(8), 256, 0, rp0, col0, rp1, col1, p.part.lo, p.part.hi, cur, nxt,
                                  active.p);
                        if (p.nL)
                            GX_LAUNCH(k_cdlp_mark_pieces, (unsigned)p.n_ins, 256, 0, p.listL.p, p.ins_row.p, p.ins_side.p,
                                      p.ins_begin.p, rp0, col0, rp1, col1, cur, nxt, active.p);
                        // a rank marks from the changed rows of its own block; the maps are OR-ed
                        if (multi()) allreduce(active.p, n, Dt::U8, Red::Max);
                        GX_LAUNCH(k_cdlp_active_entries, grid_persistent(8), 256, 0, rp0, rp1, p.part.lo, p.part.hi, active.p, stats.p + 2);
                        act = active.p;
                    } else {
                        act = nullptr;
                    }
                }
                uint32_t *t = cur; cur = nxt; nxt = t;
            }
            if (sparse_iters) {
                if (multi()) allreduce(stats.p + 2, 1, Dt::U64, Red::Sum);
                unsigned long long h = 0;
                read_back(&h, stats.p + 2, sizeof(h));
                inspected += h;
            }
            GX_LAUNCH(k_widen_u32_cdlp, grid_persistent(8), 256, 0, cur, n, g->res_u64.p);
        }
        c.timing.iterations = iters;
        c.timing.edges_inspected = inspected;
        c.timing.algorithmic_bytes = (uint64_t)iters * (4 * m_eff + (g->directed ? 2 : 1) * 8 * (n + 1) + 8 * n);
        if (label_host) {
            PhaseTimer td(&c.timing.d2h_ms);
            GX_CUDA(cudaMemcpyAsync(label_host, g->res_u64.p, n * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        }
        GX_CUDA(cudaStreamSynchronize(c.stream));
    });
}
