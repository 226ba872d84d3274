// comm.cuh -- multi-GPU plumbing: one process per GPU, NCCL over NVLink.
//
// Sharding model: the ROWS are split into contiguous blocks balanced by entry count and a rank
// computes the rows of its block.  The out-adjacency is on every rank (every process is handed the same host arrays,
// uploads 1/nranks of them over PCIe and receives the rest by all-gather over NVLink, graph.cu); the in-adjacency is
// row-partitioned (block-local transposition: a rank sorts and keeps the in-edge rows of its block only).  Per-vertex
// state is exchanged after each bulk-synchronous step in the cheapest form that works for the algorithm: PageRank
// pushes only the slots a consumer gathers into that consumer's compact vector (peer stores), SSSP forwards
// improvements to the owner's distance by peer atomics, BFS all-reduces next-frontier words, CDLP all-gathers labels,
// WCC / LCC reduce replicas; the small per-round scalars go through peer mailboxes (below) instead of NCCL.
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

// Latency-critical tiny exchanges (a queue size and a minimum per round of a frontier algorithm) go through
// peer-mapped MAILBOXES instead of NCCL + a device-to-host copy: a rank stores its words into slot [rank] of every
// peer's mailbox over NVLink, spins until all slots of its own mailbox carry the round's sequence number, and
// leaves the reduced result in host-mapped memory where the host polls it.  Like an all-reduce it is also a barrier:
// a rank's words are stored after its earlier kernels on the stream have completed.  Two slot sets alternate by the
// parity of the sequence number (a rank can run at most one exchange ahead of a peer).
constexpr int MAIL_WORDS = 4; // a, b, sequence number, (pad)
struct MailTable {
    unsigned long long *peer[MAX_PEERS]; // rank r's mailbox as seen from here
    int nranks, rank;
};
struct PeerMail {
    PeerBuf buf;                          // 2 * MAX_PEERS * MAIL_WORDS words
    MailTable table{};
    unsigned long long *host = nullptr;   // pinned, mapped: [parity][4] = max a, max b, local extra, sequence number
    unsigned long long *host_dev = nullptr;
    unsigned long long seq = 0;
    bool ok = false;
};
void peer_mail_open(PeerMail &m);   // collective; m.ok = false when the peer mapping is unavailable
void peer_mail_close(PeerMail &m);
// the communicator's own mailbox: opened by the first caller (collective: all ranks reach it in the same call),
// reused by every later exchange (sequence numbers keep counting), closed with the communicator
PeerMail &context_mail();
// waits (host side) for the result of the exchange with sequence number `seq`: out = {max a, max b, extra}
void peer_mail_wait(PeerMail &m, unsigned long long seq, unsigned long long out[3]);

#ifdef __CUDACC__
// one warp (threadIdx.x < 32 of one CTA); returns on every lane; the result goes to host memory
__device__ __forceinline__ void peer_mail_exchange(const MailTable &t, unsigned long long a, unsigned long long b,
                                                   unsigned long long extra, unsigned long long seq,
                                                   unsigned long long *__restrict__ host_out)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned par = (unsigned)(seq & 1ull);
    if ((int)lane < t.nranks) {
        volatile unsigned long long *slot = t.peer[lane] + ((size_t)par * MAX_PEERS + t.rank) * MAIL_WORDS;
        slot[0] = a;
        slot[1] = b;
        __threadfence_system();
        slot[2] = seq;
    }
    unsigned long long ma = 0, mb = 0;
    bool bad = false;
    if ((int)lane < t.nranks) {
        volatile unsigned long long *mine = t.peer[t.rank] + ((size_t)par * MAX_PEERS + lane) * MAIL_WORDS;
        const long long t0 = clock64();
        while (mine[2] != seq)
            if (clock64() - t0 > 8000000000ll) { bad = true; break; } // ~4 s: a peer died; the host raises an error
        __threadfence_system();
        ma = mine[0];
        mb = mine[1];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long xa = __shfl_xor_sync(0xffffffffu, ma, o), xb = __shfl_xor_sync(0xffffffffu, mb, o);
        ma = xa > ma ? xa : ma;
        mb = xb > mb ? xb : mb;
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        volatile unsigned long long *h = host_out + (size_t)par * 4;
        h[0] = ma;
        h[1] = mb;
        h[2] = extra;
        __threadfence_system();
        h[3] = bad ? ~0ull : seq;
    }
}

// Same mailboxes, result kept on the device: every rank's double summed in rank order -- the same bits on every rank --
// and, like the exchange above, a barrier (a peer's word arrives after that peer's earlier kernels on its stream, i.e.
// after all its peer stores).  No host involvement: an iteration loop can queue these back to back.  One warp.
__device__ __forceinline__ double peer_mail_sum_f64(const MailTable &t, double x, unsigned long long seq)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned par = (unsigned)(seq & 1ull);
    if ((int)lane < t.nranks) {
        volatile unsigned long long *slot = t.peer[lane] + ((size_t)par * MAX_PEERS + t.rank) * MAIL_WORDS;
        slot[0] = (unsigned long long)__double_as_longlong(x);
        __threadfence_system();
        slot[2] = seq;
    }
    double v = 0.0;
    bool bad = false;
    if ((int)lane < t.nranks) {
        volatile unsigned long long *mine = t.peer[t.rank] + ((size_t)par * MAX_PEERS + lane) * MAIL_WORDS;
        const long long t0 = clock64();
        while (mine[2] != seq)
            if (clock64() - t0 > 8000000000ll) { bad = true; break; } // ~4 s: a peer died
        __threadfence_system();
        v = __longlong_as_double((long long)mine[0]);
    }
    if (__any_sync(0xffffffffu, bad)) __trap(); // surfaces as a CUDA error at the caller's next synchronisation
    double sum = 0.0;
    for (int r = 0; r < t.nranks; r++) sum += __shfl_sync(0xffffffffu, v, r);
    return sum;
}
#endif

enum class Red { Sum, Min, Max };
enum class Dt { U32, I32, U64, F64, U8 };

inline bool multi() { return ctx().nranks > 1; }
// in place: rank r contributes elements [p.b[r] / div, p.b[r+1] / div) of buf (div = 32 for bitmap words)
void allgatherv(void *buf, Dt dt, const Partition &p, uint64_t div = 1, uint64_t total = 0);
// in place, equal blocks: rank r contributes elements [r * count_per_rank, (r+1) * count_per_rank) of buf
void allgather_equal(void *buf, Dt dt, uint64_t count_per_rank);
void allreduce(void *buf, uint64_t count, Dt dt, Red op);

} // namespace gx
