// comm.cuh -- multi-GPU plumbing: one process per GPU, NCCL over NVLink.
//
// Sharding model: every rank keeps the whole adjacency in its own HBM
// (RMAT-26 symmetric + FP64 weights is ~26 GB of 180 GB), the ROWS are split into contiguous
// blocks balanced by entry count, and the dense per-vertex state is replicated: after each
// bulk-synchronous step the owned slices are all-gathered (PR ranks, CDLP labels, BFS frontier
// bitmap) or the replicas min/sum-reduced (WCC parents, SSSP distances, LCC counts).
// Getting the graph in follows the same idea in reverse: every rank is handed the same host arrays,
// uploads 1/nranks of them over PCIe and receives the rest by all-gather over NVLink (graph.cu).
#pragma once

#include <vector>

#include "common.cuh"

namespace gx {

struct Partition {
    std::vector<uint64_t> b; // nranks + 1 boundaries, b[0] = 0, b[nranks] = n, multiples of 32 in between
    uint64_t lo = 0, hi = 0; // this rank's block [lo, hi)
};

// Row blocks balanced by entries of rp0 (+ rp1 when given); device arrays of n+1 offsets.
Partition make_partition(const uint64_t *rp0, const uint64_t *rp1, uint64_t n);
// Even split of [0, count) (work lists that are already balanced per item).
Partition make_even_partition(uint64_t count, uint64_t align = 1);

// A buffer every rank allocates with the same size and maps into all other ranks (CUDA IPC over
// NVLink peer access): kernels store their results straight into the peers' copies, so a
// compute step and the exchange of its output are one kernel ("fused collective").
constexpr int MAX_PEERS = 8;
struct PeerBuf {
    void *local = nullptr;
    void *peer[MAX_PEERS] = {nullptr}; // peer[r] = rank r's buffer as seen from here (peer[rank] == local)
    size_t bytes = 0, capacity = 0;    // requested / allocated
    bool shared = false;               // false: single rank or peer mapping unavailable
    bool pooled = false;               // allocated from the stream-ordered pool (not exportable)
};
void peer_alloc(PeerBuf &b, size_t bytes);
void peer_free(PeerBuf &b);
void peer_cache_clear(); // releases the buffers parked for reuse (before the communicator goes away)

enum class Red { Sum, Min, Max };
enum class Dt { U32, I32, U64, F64, U8 };

inline bool multi() { return ctx().nranks > 1; }
// in place: rank r contributes elements [p.b[r] / div, p.b[r+1] / div) of buf (div = 32 for bitmap words)
void allgatherv(void *buf, Dt dt, const Partition &p, uint64_t div = 1, uint64_t total = 0);
// in place, equal blocks: rank r contributes elements [r * count_per_rank, (r+1) * count_per_rank) of buf
void allgather_equal(void *buf, Dt dt, uint64_t count_per_rank);
void allreduce(void *buf, uint64_t count, Dt dt, Red op);

} // namespace gx
