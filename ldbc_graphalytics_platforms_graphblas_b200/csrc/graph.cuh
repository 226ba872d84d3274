// graph.cuh -- device-resident graph: CSR (out-edges), cached CSC (in-edges),
// cached degree-oriented union graph for LCC, result buffers.
//
// HBM layout (all structure-of-arrays, 256-byte aligned by the allocator):
//   rowptr  uint64[n+1]   8-byte offsets (nnz of a symmetric RMAT-26 exceeds 2^31)
//   col     uint32[m]     4-byte vertex ids, sorted inside each row
//   w       double[m]     only for weighted graphs (SSSP)
// The reference keeps 8-byte indices everywhere (GrB_Index, cdlp_kernel.cu);
// halving the index width halves the dominant adjacency stream.
#pragma once

#include "comm.cuh"
#include "common.cuh"

namespace gx {

// Degree-skew plan of one adjacency: rows longer than ROW_SPLIT entries are cut
// into CHUNK-entry pieces that one CTA each streams; short rows go to sub-warp
// groups.  RMAT hubs (10^5..10^6 entries) would otherwise serialise on one warp.
constexpr uint32_t ROW_SPLIT = 256;  // rows with more entries take the chunked path
constexpr uint32_t CHUNK = 2048;     // entries per CTA on the chunked path

struct RowPlan {
    bool built = false;
    Partition part;                    // row blocks of the ranks (whole range on one GPU)
    uint64_t n_long = 0, n_chunks = 0; // long rows / chunks inside this rank's block
    DevBuf<uint32_t> long_rows;        // n_long, ascending vertex ids
    DevBuf<uint32_t> long_first_chunk; // n_long + 1
    DevBuf<uint32_t> chunk_row;        // n_chunks: vertex id
    DevBuf<uint64_t> chunk_begin;      // n_chunks: first entry offset
};

struct Adj {
    DevBuf<uint64_t> rowptr;
    DevBuf<uint32_t> col;
    DevBuf<double> w;
    RowPlan plan;
};

constexpr uint32_t LCC_MULT_BIT = 0x80000000u; // oriented col entry: bit31 = reciprocal pair
constexpr uint32_t LCC_TAB_MIN = 16;           // oriented rows with at least this many entries get a membership table
#ifdef __CUDACC__
__host__ __device__ __forceinline__ uint64_t lcc_tab_hash(uint32_t id, uint64_t size_pow2)
{
    return ((uint64_t)(id * 2654435761u) * size_pow2) >> 32; // multiplicative hash, top bits
}
#endif

// builders shared by graph.cu / rmat.cu
void expand_row_ids(const uint64_t *rowptr, uint64_t n, uint64_t m, uint32_t *row_of_edge);
void rowptr_from_sorted_rows(const uint32_t *sorted_rows, uint64_t m, uint64_t n, uint64_t *rowptr);
void rowptr_from_sorted_keys(const uint64_t *sorted_keys, uint64_t m, uint64_t n, uint64_t *rowptr);
void sort_keys64(DevBuf<uint64_t> &keys, uint64_t count, int end_bit);
void sort_keys32(DevBuf<uint32_t> &keys, uint64_t count, int end_bit);
void sort_pairs64_f64(DevBuf<uint64_t> &keys, DevBuf<double> &vals, uint64_t count, int end_bit);
int bits_for(uint64_t n);
uint64_t select_flagged(const uint64_t *in, const uint8_t *flags, uint64_t count, DevBuf<uint64_t> &out);

} // namespace gx

struct gx_graph {
    uint64_t n = 0, m = 0;
    bool directed = false, weighted = false;
    double mean_weight = 1.0;         // of the stored FP64 values (SSSP bucket width)
    bool have_mean_weight = false;
    gx::Adj out;          // CSR: row v = out-neighbours of v
    gx::Adj in;           // CSC: row v = in-neighbours of v (directed only, built lazily)
    bool have_in = false;
    // several GPUs: after the block-local transposition a rank holds in.rowptr whole but only the in.col entries of its
    // own row block (in.plan.part) -- all PageRank and the pull levels of BFS ever read; in_block_entries[r] is the first
    // entry of rank r's block.  ensure_in_full() all-gathers the other blocks for the algorithms that need them.
    bool in_block_only = false;
    std::vector<uint64_t> in_block_entries;

    // LCC cache: U = A v A' without self-loops, oriented low -> high (degree, id)
    bool have_lcc = false;
    uint64_t om = 0;                  // oriented entries
    gx::DevBuf<uint64_t> orowptr;     // n+1
    gx::DevBuf<uint32_t> ocol;        // om, sorted; bit31 = both directions present in A
    gx::DevBuf<uint32_t> orow;        // om, source vertex of each oriented entry
    gx::DevBuf<uint32_t> udeg;        // n, degree in U
    uint64_t lcc_list_bytes = 0;      // 4 * sum over oriented edges of (d+(u) + d+(v))
    // membership tables of the longer oriented rows (open addressing, 4 slots per entry): one or two
    // probes instead of a binary search when a short list is intersected with a hub's list
    gx::DevBuf<uint64_t> ltab_off;    // n+1: first slot of row v's table (empty range: no table)
    gx::DevBuf<uint32_t> ltab;        // slots hold ocol values (id | multiplicity bit), 0xFFFFFFFF = empty
    // the oriented entries once more, ordered by the owner of the LONGER list of each intersection: the groups
    // running at any time then probe a handful of tables (L1/L2 hits) instead of thousands (DRAM sectors)
    gx::DevBuf<uint32_t> lcc_eu, lcc_ev; // om each: source, target | multiplicity bit
    gx::DevBuf<uint32_t> lcc_owner;      // om: that owner (the sort key), ascending -- consecutive entries of one owner form a segment

    void *cdlp_plan = nullptr;        // gx::CdlpPlan (algo_cdlp.cu), degree bins + spill tables
    void *pr_cache = nullptr;         // gx::PrTiles (algo_pr.cu), tiling of the in-edge entries
    void *sssp_cache = nullptr;       // gx::SsspCache (algo_sssp.cu), rows partitioned into light | heavy entries

    // results of the last run of each algorithm stay on the device
    gx::DevBuf<int64_t> res_i64;
    gx::DevBuf<uint64_t> res_u64;
    gx::DevBuf<double> res_f64;

    const gx::Adj &in_adj() const { return directed ? in : out; }
    gx::Adj &in_adj() { return directed ? in : out; }
    ~gx_graph();
};

namespace gx {
void ensure_in_adj(gx_graph *g);   // LAGraph_Cached_AT analogue
void ensure_in_full(gx_graph *g);  // several GPUs: every rank gets the in.col entries of all row blocks
void ensure_lcc_cache(gx_graph *g);
void finish_graph(gx_graph *g);    // validation + row sorting after upload
void ensure_plan(Adj &a, uint64_t n); // long-row chunk plan of one adjacency
} // namespace gx
