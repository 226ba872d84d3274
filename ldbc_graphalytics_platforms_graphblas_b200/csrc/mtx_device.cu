// mtx_device.cu -- graph.mtx -> device CSR with the text tokenised on the GPU.
//
// Replaces ReadMatrixMarket (graphio.cpp:10-24: LAGraph_MMRead of `graph.mtx`) + the upload for text inputs: the
// file's bytes go to HBM as they are (memory-mapped, copied in 64 MB pieces on a few host threads), and the device
// finds the entries, parses them and builds the CSR -- the host never looks at the body.  The host-threaded parser
// (host/graphio.cpp, ReadMtxFile) reads ~0.12 GB/s end to end; an RMAT-26 edge list is ~30 GB of text.
//
//   k_mtx_count   a 256-thread CTA per 4 KB of text, 16 bytes per thread: entry starts (first non-blank byte of a
//                 line) counted per block
//   (scan)        block offsets -> index of every entry; the total must equal the size line's nnz
//   k_mtx_parse   same decomposition; the thread that owns an entry's first byte parses `i j [value]`: decimal
//                 integers, and FP64 values correctly rounded on the device (decimal_to_double.cuh: Clinger's fast
//                 path, then Eisel-Lemire with a 128-bit power-of-five table); the few values it cannot decide
//                 (> 19 digits, subnormal range, "inf") are listed and settled by the host's strtod on the mapped file
//   (sort)        (row << 32 | col) keys -- both orientations for `symmetric` files, self-loops dropped -- radix-sorted
//                 (with the values as payload), duplicates removed keeping the smallest weight: the same cleaning
//                 ReadMtxFile applies, so both loaders produce identical arrays (tests/test_gpu_parity.py)
// The result is a gx_graph like one made by gx_graph_create_csr32 from ReadMtxFile's arrays.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <cub/cub.cuh>

#include "decimal_to_double.cuh"
#include "graph.cuh"

namespace gx {

static const uint64_t POW5_HOST[] = {
#include "pow5_table.inc"
};

constexpr int MTX_BLOCK_BYTES = 4096; // per CTA: 256 threads x 16 bytes
constexpr int MTX_FRONT_PAD = 16;     // '\n' bytes before the body (text[p - 1] is always readable, 16-byte alignment)
constexpr int MTX_TAIL_PAD = 64;      // '\n' bytes after it (a last line without a newline still ends)

struct MtxFlags { int malformed, out_of_range; unsigned long long n_fallback; };

__device__ __forceinline__ bool mtx_blank(unsigned char c) { return c == ' ' || c == '\t' || c == '\r'; }

// bit i: byte base + i is the first non-blank byte of a line.  `text` points at the padded buffer.
__device__ __forceinline__ unsigned mtx_start_mask(const unsigned char *__restrict__ text, uint64_t base, uint64_t end)
{
    const uint4 v = *(const uint4 *)(text + base);
    const uint32_t wd[4] = {v.x, v.y, v.z, v.w};
    unsigned char prev = text[base - 1];
    unsigned mask = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const unsigned char c = (unsigned char)(wd[i >> 2] >> ((i & 3) * 8));
        if (base + i < end && c != '\n' && !mtx_blank(c)) {
            bool st = prev == '\n';
            if (!st && mtx_blank(prev)) { // leading blanks of a line (rare): walk back to what precedes them
                uint64_t q = base + i - 1;
                while (mtx_blank(text[q])) q--;
                st = text[q] == '\n';
            }
            if (st) mask |= 1u << i;
        }
        prev = c;
    }
    return mask;
}

__global__ void __launch_bounds__(256)
k_mtx_count(const unsigned char *__restrict__ text, uint64_t begin, uint64_t end, uint64_t *__restrict__ block_count)
{
    const uint64_t nblocks = (end - begin + MTX_BLOCK_BYTES - 1) / MTX_BLOCK_BYTES;
    for (uint64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const uint64_t base = begin + b * MTX_BLOCK_BYTES + 16ull * threadIdx.x;
        unsigned cnt = base < end ? __popc(mtx_start_mask(text, base, end)) : 0u;
        cnt = warp_sum(cnt);
        __shared__ unsigned s[8];
        if (lane_id() == 0) s[threadIdx.x >> 5] = cnt;
        __syncthreads();
        if (threadIdx.x == 0) block_count[b] = (uint64_t)s[0] + s[1] + s[2] + s[3] + s[4] + s[5] + s[6] + s[7];
        __syncthreads();
    }
}

struct DevBytes {
    const unsigned char *p;
    __device__ __forceinline__ unsigned char operator[](uint64_t i) const { return p[i]; }
};

// decimal integer at text[p]; p advances past it; ok = false without digits or beyond 19 of them
__device__ __forceinline__ uint64_t mtx_parse_u64(const unsigned char *__restrict__ text, uint64_t &p, bool &ok)
{
    uint64_t v = 0;
    int nd = 0;
    unsigned char c = text[p];
    while (c >= '0' && c <= '9') { v = v * 10 + (uint64_t)(c - '0'); nd++; c = text[++p]; }
    if (nd == 0 || nd > 19) ok = false;
    return v;
}

// MODE bit 0: the file is `symmetric` (both orientations are emitted), bit 1: entries carry a value token,
// bit 2: the values are kept (FP64 weights).  Keys: (row << 32) | col, 0-based; ~0 for what is dropped.
template <int MODE>
__global__ void __launch_bounds__(256)
k_mtx_parse(const unsigned char *__restrict__ text, uint64_t begin, uint64_t end, const uint64_t *__restrict__ block_off,
            uint64_t n, uint64_t *__restrict__ keys, double *__restrict__ vals, const double *__restrict__ p10,
            const uint64_t *__restrict__ pow5, uint64_t *__restrict__ fb_entry, uint64_t *__restrict__ fb_offset,
            uint64_t fb_cap, MtxFlags *__restrict__ flags)
{
    constexpr bool SYM = MODE & 1, HAS_VALUE = (MODE & 2) != 0, KEEP = (MODE & 4) != 0;
    const uint64_t nblocks = (end - begin + MTX_BLOCK_BYTES - 1) / MTX_BLOCK_BYTES;
    __shared__ unsigned s_warp[8];
    for (uint64_t b = blockIdx.x; b < nblocks; b += gridDim.x) {
        const uint64_t base = begin + b * MTX_BLOCK_BYTES + 16ull * threadIdx.x;
        const unsigned mask = base < end ? mtx_start_mask(text, base, end) : 0u;
        // exclusive rank of the thread's first entry inside the block
        const unsigned mine = __popc(mask);
        unsigned incl = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const unsigned up = __shfl_up_sync(FULL, incl, d);
            if (lane_id() >= (unsigned)d) incl += up;
        }
        if (lane_id() == 31) s_warp[threadIdx.x >> 5] = incl;
        __syncthreads();
        unsigned before = incl - mine;
        for (unsigned w = 0; w < (threadIdx.x >> 5); w++) before += s_warp[w];
        uint64_t idx = block_off[b] + before;
        for (unsigned m = mask; m; m &= m - 1, idx++) {
            uint64_t p = base + (uint64_t)(__ffs(m) - 1);
            bool ok = true;
            const uint64_t i = mtx_parse_u64(text, p, ok);
            while (mtx_blank(text[p])) p++;
            const uint64_t j = mtx_parse_u64(text, p, ok);
            double x = 1.0;
            if (HAS_VALUE) {
                if (!mtx_blank(text[p])) ok = false; // the column index must end at a blank
                while (mtx_blank(text[p])) p++;
                if (text[p] == '\n') ok = false;     // entry without a value
                else if (KEEP) {
                    const uint64_t tok = p;
                    const int st = parse_double_token(DevBytes{text}, p, p10, pow5, &x);
                    if (st != 0) { // undecided: the host settles it from the same bytes
                        const unsigned long long k = atomicAdd(&flags->n_fallback, 1ull);
                        if (k < fb_cap) { fb_entry[k] = idx; fb_offset[k] = tok; }
                        x = 0.0;
                    }
                }
            } else if (!(mtx_blank(text[p]) || text[p] == '\n')) ok = false;
            if (!ok) { flags->malformed = 1; continue; }
            if (i < 1 || j < 1 || i > n || j > n) { flags->out_of_range = 1; continue; }
            const uint64_t r = i - 1, c = j - 1;
            const uint64_t fwd = r == c ? ~0ull : ((r << 32) | c);
            if (SYM) {
                keys[2 * idx] = fwd;
                keys[2 * idx + 1] = r == c ? ~0ull : ((c << 32) | r);
                if (KEEP) { vals[2 * idx] = x; vals[2 * idx + 1] = x; }
            } else {
                keys[idx] = fwd;
                if (KEEP) vals[idx] = x;
            }
        }
        __syncthreads();
    }
}

// values the host decided: entry index -> value (both orientations of a symmetric file)
__global__ void k_mtx_patch(const uint64_t *__restrict__ entry, const double *__restrict__ value, uint64_t count, int sym,
                            double *__restrict__ vals)
{
    uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; k < count; k += stride) {
        if (sym) { vals[2 * entry[k]] = value[k]; vals[2 * entry[k] + 1] = value[k]; }
        else vals[entry[k]] = value[k];
    }
}

// sorted (key, value) pairs: head = first of a run of equal keys (dropped keys are ~0); the head takes the run's
// smallest value (ReadMtxFile sorts a row by (column, weight) and keeps the first of equal columns)
__global__ void k_mtx_heads(const uint64_t *__restrict__ keys, double *__restrict__ vals, uint64_t cnt, uint8_t *__restrict__ head)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; i < cnt; i += stride) {
        const uint64_t k = keys[i];
        const bool h = k != ~0ull && (i == 0 || keys[i - 1] != k);
        head[i] = h ? 1 : 0;
        if (h && vals) {
            double m = vals[i];
            for (uint64_t e = i + 1; e < cnt && keys[e] == k; e++) m = vals[e] < m ? vals[e] : m;
            vals[i] = m; // (only heads are read afterwards; a run's other slots are never heads)
        }
    }
}

__global__ void k_mtx_low32(const uint64_t *__restrict__ keys, uint64_t m, uint32_t *__restrict__ out)
{
    uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (; e < m; e += stride) out[e] = (uint32_t)keys[e];
}

namespace {

struct MappedFile {
    int fd = -1;
    const char *data = nullptr;
    size_t size = 0;
    explicit MappedFile(const std::string &path)
    {
        fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) throw Error(GX_ERR_IO, "Cannot open file: " + path);
        struct stat st;
        if (fstat(fd, &st) != 0) { close(fd); throw Error(GX_ERR_IO, "Cannot stat file: " + path); }
        size = (size_t)st.st_size;
        if (size) {
            void *p = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
            if (p == MAP_FAILED) { close(fd); throw Error(GX_ERR_IO, "Cannot map file: " + path); }
            data = (const char *)p;
            madvise(p, size, MADV_SEQUENTIAL);
        }
    }
    MappedFile(const MappedFile &) = delete;
    MappedFile &operator=(const MappedFile &) = delete;
    ~MappedFile()
    {
        if (data) munmap((void *)data, size);
        if (fd >= 0) close(fd);
    }
};

struct MtxHeader { bool symmetric = false, pattern = false, weighted = false; uint64_t n = 0, nnz = 0; size_t body = 0; };

// banner, comment lines and the size line (the only part of the file the host reads)
MtxHeader parse_header(const MappedFile &f, const std::string &path)
{
    MtxHeader h;
    const char *p = f.data, *end = f.data + f.size;
    if (f.size < 14 || std::strncmp(p, "%%MatrixMarket", 14) != 0) throw Error(GX_ERR_IO, "Not a Matrix Market file: " + path);
    auto line_end = [&](const char *q) { while (q < end && *q != '\n') q++; return q; };
    const char *e = line_end(p);
    std::string banner(p, e);
    std::transform(banner.begin(), banner.end(), banner.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (banner.find("coordinate") == std::string::npos) throw Error(GX_ERR_IO, "Only coordinate Matrix Market files are supported");
    h.symmetric = banner.find("symmetric") != std::string::npos;
    h.pattern = banner.find("pattern") != std::string::npos;
    h.weighted = banner.find("real") != std::string::npos || banner.find("double") != std::string::npos;
    p = e < end ? e + 1 : end;
    while (p < end && *p == '%') { e = line_end(p); p = e < end ? e + 1 : end; }
    e = line_end(p);
    std::string size_line(p, e);
    unsigned long long a = 0, b = 0, c = 0;
    if (sscanf(size_line.c_str(), "%llu %llu %llu", &a, &b, &c) != 3) throw Error(GX_ERR_IO, "Matrix Market size line expected: " + path);
    if (a != b) throw Error(GX_ERR_IO, "Adjacency matrix must be square");
    if (a >= 0xFFFFFFFEull) throw Error(GX_ERR_IO, "More than 2^32 - 2 vertices are not supported");
    h.n = a;
    h.nnz = c;
    h.body = (size_t)((e < end ? e + 1 : end) - f.data);
    return h;
}

// the body's bytes to the device: 64 MB pieces of the mapping, a few host threads with a stream each (the copies out
// of pageable memory are staged by the driver; several in flight keep the page-cache reads and the bus busy)
void upload_text(unsigned char *dst, const char *src, size_t bytes)
{
    const size_t PIECE = 64u << 20;
    const size_t pieces = (bytes + PIECE - 1) / PIECE;
    unsigned T = std::thread::hardware_concurrency();
    if (const char *e = getenv("GX_LOADER_THREADS")) T = (unsigned)atoi(e);
    T = std::max(1u, std::min({T, 4u, (unsigned)pieces}));
    int dev = 0;
    GX_CUDA(cudaGetDevice(&dev));
    std::vector<cudaError_t> err(T, cudaSuccess);
    auto body = [&](unsigned t) {
        cudaSetDevice(dev);
        cudaStream_t s = nullptr;
        if ((err[t] = cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking)) != cudaSuccess) return;
        for (size_t k = t; k < pieces && err[t] == cudaSuccess; k += T) {
            const size_t o = k * PIECE, len = std::min(PIECE, bytes - o);
            err[t] = cudaMemcpyAsync(dst + o, src + o, len, cudaMemcpyHostToDevice, s);
        }
        const cudaError_t e2 = cudaStreamSynchronize(s);
        if (err[t] == cudaSuccess) err[t] = e2;
        cudaStreamDestroy(s);
    };
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < T; t++) pool.emplace_back(body, t);
    body(0);
    for (auto &th : pool) th.join();
    for (cudaError_t e : err) GX_CUDA(e);
}

} // namespace

// GX_TIMING_DEBUG=1: wall-clock per phase of the load on stderr
struct LoadLog {
    bool on = getenv("GX_TIMING_DEBUG") != nullptr && ctx().rank == 0;
    std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    void mark(const char *what)
    {
        if (!on) return;
        cudaStreamSynchronize(ctx().stream);
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[gx load] %-34s %8.3f ms\n", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

static gx_graph *load_mtx_device(const std::string &path, int directed)
{
    Context &c = ctx();
    LoadLog log;
    MappedFile file(path);
    const MtxHeader h = parse_header(file, path);
    const uint64_t n = h.n;
    const size_t body = file.size - h.body;
    // ---- text to HBM, padded with newlines on both sides
    GX_CUDA(cudaStreamSynchronize(c.stream));
    DevBuf<unsigned char> text(MTX_FRONT_PAD + body + MTX_TAIL_PAD);
    GX_CUDA(cudaMemsetAsync(text.p, '\n', MTX_FRONT_PAD, c.stream));
    GX_CUDA(cudaMemsetAsync(text.p + MTX_FRONT_PAD + body, '\n', MTX_TAIL_PAD, c.stream));
    GX_CUDA(cudaStreamSynchronize(c.stream)); // the allocation is stream-ordered; the upload uses streams of its own
    {
        PhaseTimer t(&c.timing.h2d_ms);
        if (body) upload_text(text.p + MTX_FRONT_PAD, file.data + h.body, body);
    }
    log.mark("text to HBM");
    PhaseTimer tb(&c.timing.build_ms);
    const uint64_t begin = MTX_FRONT_PAD, end = MTX_FRONT_PAD + body;
    const uint64_t nblocks = (body + MTX_BLOCK_BYTES - 1) / MTX_BLOCK_BYTES;
    uint64_t entries = 0;
    DevBuf<uint64_t> block_off(nblocks + 1);
    if (nblocks) {
        DevBuf<uint64_t> block_count(nblocks + 1);
        GX_CUDA(cudaMemsetAsync(block_count.p + nblocks, 0, sizeof(uint64_t), c.stream));
        const unsigned grid = (unsigned)std::min<uint64_t>(nblocks, grid_persistent(8));
        GX_LAUNCH(k_mtx_count, grid, 256, 0, text.p, begin, end, block_count.p);
        size_t tbytes = 0;
        GX_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tbytes, block_count.p, block_off.p, (int64_t)(nblocks + 1), c.stream));
        DevBuf<char> tmp(tbytes);
        GX_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, tbytes, block_count.p, block_off.p, (int64_t)(nblocks + 1), c.stream));
        count_launch();
        read_back(&entries, block_off.p + nblocks, sizeof(uint64_t));
    }
    log.mark("count entries + scan");
    if (entries < h.nnz) throw Error(GX_ERR_IO, "Matrix Market file ends before nnz entries were read");
    if (entries > h.nnz) throw Error(GX_ERR_IO, "Matrix Market file holds more entries than its size line announces");
    // ---- parse
    const uint64_t cnt = h.symmetric ? 2 * entries : entries;
    DevBuf<uint64_t> keys(cnt ? cnt : 1);
    DevBuf<double> vals(h.weighted && cnt ? cnt : 1);
    const uint64_t fb_cap = entries ? entries : 1;
    DevBuf<uint64_t> fb_entry, fb_offset;
    if (h.weighted) { fb_entry.alloc(fb_cap); fb_offset.alloc(fb_cap); }
    DevBuf<MtxFlags> flags(1);
    flags.zero();
    DevBuf<double> p10(23);
    DevBuf<uint64_t> pow5(sizeof(POW5_HOST) / sizeof(uint64_t));
    {
        double hp[23];
        hp[0] = 1.0;
        for (int i = 1; i < 23; i++) hp[i] = hp[i - 1] * 10.0; // exact up to 1e22
        GX_CUDA(cudaMemcpyAsync(p10.p, hp, sizeof(hp), cudaMemcpyHostToDevice, c.stream));
        GX_CUDA(cudaMemcpyAsync(pow5.p, POW5_HOST, sizeof(POW5_HOST), cudaMemcpyHostToDevice, c.stream));
        GX_CUDA(cudaStreamSynchronize(c.stream)); // hp is on the stack
    }
    if (nblocks) {
        const unsigned grid = (unsigned)std::min<uint64_t>(nblocks, grid_persistent(8));
        const int mode = (h.symmetric ? 1 : 0) | (h.pattern ? 0 : 2) | (h.weighted ? 4 : 0);
#define GX_MTX_PARSE(M)                                                                                                        \
    GX_LAUNCH(k_mtx_parse<M>, grid, 256, 0, text.p, begin, end, block_off.p, n, keys.p, vals.p, p10.p, pow5.p, fb_entry.p,     \
              fb_offset.p, fb_cap, flags.p)
        switch (mode) {
        case 0: GX_MTX_PARSE(0); break;
        case 1: GX_MTX_PARSE(1); break;
        case 2: GX_MTX_PARSE(2); break;
        case 3: GX_MTX_PARSE(3); break;
        case 6: GX_MTX_PARSE(6); break;
        case 7: GX_MTX_PARSE(7); break;
        default: throw Error(GX_ERR_IO, "Matrix Market banner: `pattern` files cannot carry real values");
        }
#undef GX_MTX_PARSE
    }
    MtxFlags hf;
    read_back(&hf, flags.p, sizeof(hf));
    log.mark("parse");
    c.timing.edges_inspected = entries;
    c.timing.iterations = (uint32_t)std::min<unsigned long long>(hf.n_fallback, 0xFFFFFFFFull); // values left to the host
    if (hf.malformed) throw Error(GX_ERR_IO, "Malformed Matrix Market entry (expected `row column [value]`)");
    if (hf.out_of_range) throw Error(GX_ERR_IO, "Matrix Market entry out of range");
    if (hf.n_fallback) {
        // values the device left undecided: strtod on the mapped bytes, patched in by entry index
        const uint64_t k = hf.n_fallback;
        std::vector<uint64_t> he(k), ho(k);
        std::vector<double> hv(k);
        GX_CUDA(cudaMemcpyAsync(he.data(), fb_entry.p, k * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        GX_CUDA(cudaMemcpyAsync(ho.data(), fb_offset.p, k * sizeof(uint64_t), cudaMemcpyDeviceToHost, c.stream));
        GX_CUDA(cudaStreamSynchronize(c.stream));
        unsigned T = std::max(1u, std::min(std::thread::hardware_concurrency(), 64u));
        if (k < (1u << 16)) T = 1;
        std::vector<int> bad(T, 0);
        auto work = [&](unsigned t) {
            char tok[512];
            for (uint64_t x = k * t / T; x < k * (t + 1) / T; x++) {
                const char *q = file.data + h.body + (ho[x] - MTX_FRONT_PAD), *fe = file.data + file.size;
                size_t len = 0;
                while (q + len < fe && len + 1 < sizeof(tok) && q[len] != '\n' && q[len] != ' ' && q[len] != '\t' && q[len] != '\r') len++;
                memcpy(tok, q, len);
                tok[len] = '\0';
                char *e = nullptr;
                hv[x] = std::strtod(tok, &e);
                if (e == tok) bad[t] = 1;
            }
        };
        std::vector<std::thread> pool;
        for (unsigned t = 1; t < T; t++) pool.emplace_back(work, t);
        work(0);
        for (auto &th : pool) th.join();
        for (int b : bad)
            if (b) throw Error(GX_ERR_IO, "Matrix Market entry without a value");
        DevBuf<double> dv(k);
        GX_CUDA(cudaMemcpyAsync(dv.p, hv.data(), k * sizeof(double), cudaMemcpyHostToDevice, c.stream));
        GX_LAUNCH(k_mtx_patch, grid_persistent(4), 256, 0, fb_entry.p, dv.p, k, h.symmetric ? 1 : 0, vals.p);
        GX_CUDA(cudaStreamSynchronize(c.stream)); // hv
    }
    if (hf.n_fallback) log.mark("host strtod of undecided values");
    text.release();
    // ---- COO -> CSR: sorted rows, no self-loops, no duplicates (the smallest weight of a repeated entry)
    gx_graph *g = new gx_graph();
    try {
        g->n = n;
        g->directed = directed != 0;
        g->weighted = h.weighted;
        uint64_t m = 0;
        DevBuf<uint64_t> ukeys(cnt ? cnt : 1);
        if (cnt) {
            const int end_bit = 32 + bits_for(n);
            if (h.weighted) sort_pairs64_f64(keys, vals, cnt, end_bit);
            else sort_keys64(keys, cnt, end_bit);
            DevBuf<uint8_t> head(cnt);
            GX_LAUNCH(k_mtx_heads, grid_persistent(8), 256, 0, keys.p, h.weighted ? vals.p : nullptr, cnt, head.p);
            m = select_flagged(keys.p, head.p, cnt, ukeys);
            if (h.weighted) {
                DevBuf<uint64_t> uvals(cnt);
                const uint64_t m2 = select_flagged((const uint64_t *)vals.p, head.p, cnt, uvals);
                GX_REQUIRE(m2 == m, "internal: key / value selection disagree");
                g->out.w.alloc(m);
                GX_CUDA(cudaMemcpyAsync(g->out.w.p, uvals.p, m * sizeof(double), cudaMemcpyDeviceToDevice, c.stream));
                GX_CUDA(cudaStreamSynchronize(c.stream)); // uvals goes out of scope
            }
        }
        g->m = m;
        g->out.rowptr.alloc(n + 1);
        g->out.col.alloc(m);
        if (n && m) rowptr_from_sorted_keys(ukeys.p, m, n, g->out.rowptr.p);
        else g->out.rowptr.zero();
        if (m) GX_LAUNCH(k_mtx_low32, grid_persistent(8), 256, 0, ukeys.p, m, g->out.col.p);
        GX_CUDA(cudaStreamSynchronize(c.stream));
        log.mark("sort + dedupe + CSR");
    } catch (...) {
        delete g;
        throw;
    }
    return g;
}

} // namespace gx

using namespace gx;

extern "C" int gx_graph_load_mtx(gx_graph **out, const char *path, int directed, unsigned cache)
{
    return guarded([&] {
        require_ready();
        GX_REQUIRE(out != nullptr && path != nullptr, "graph handle or path is NULL");
        ctx().timing = gx_timing{};
        gx_graph *g = load_mtx_device(path, directed);
        try {
            if ((cache & GX_CACHE_AT) && g->directed) ensure_in_adj(g); // (times itself into build_ms)
            if (cache & GX_CACHE_LCC) {
                PhaseTimer t(&ctx().timing.build_ms);
                ensure_lcc_cache(g);
            }
            GX_CUDA(cudaStreamSynchronize(ctx().stream));
        } catch (...) {
            delete g;
            throw;
        }
        *out = g;
    });
}
