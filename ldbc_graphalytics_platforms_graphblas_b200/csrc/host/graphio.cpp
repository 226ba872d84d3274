#include "graphio.h"

#include <algorithm>
#include <cinttypes>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <atomic>
#include <exception>
#include <numeric>
#include <thread>

#include "computation_timer.hpp"

namespace {

constexpr size_t GRB_HEADER_LEN = 512; // LAGRAPH_BIN_HEADER

struct FileBytes {
    std::vector<char> data;
    explicit FileBytes(const std::string &path)
    {
        FILE *f = fopen(path.c_str(), "rb");
        if (!f) throw std::runtime_error("Cannot open file: " + path);
        fseek(f, 0, SEEK_END);
        long sz = ftell(f);
        fseek(f, 0, SEEK_SET);
        data.resize((size_t)sz + 1);
        if (sz && fread(data.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); throw std::runtime_error("Short read: " + path); }
        data[(size_t)sz] = '\0';
        fclose(f);
    }
};

inline const char *skip_ws(const char *p) { while (*p == ' ' || *p == '\t' || *p == '\r') p++; return p; }
inline const char *next_line(const char *p) { while (*p && *p != '\n') p++; return *p ? p + 1 : p; }

inline uint64_t parse_u64(const char *&p)
{
    p = skip_ws(p);
    uint64_t v = 0;
    if (*p < '0' || *p > '9') throw std::runtime_error("Matrix Market file: expected an integer");
    while (*p >= '0' && *p <= '9') v = v * 10 + (uint64_t)(*p++ - '0');
    return v;
}

// Worker threads of the loader: the text of an RMAT-26 graph is tens of GB, and both the parse and the
// CSR construction split cleanly (the reference's LAGraph_MMRead is single-threaded, graphio.cpp:14).
unsigned loader_threads(size_t work_items, size_t min_per_thread)
{
    unsigned t = std::thread::hardware_concurrency();
    if (const char *e = std::getenv("GX_LOADER_THREADS")) t = (unsigned)std::atoi(e);
    if (t < 1) t = 1;
    if (t > 64) t = 64;
    const size_t by_work = work_items / (min_per_thread ? min_per_thread : 1);
    if (by_work < t) t = by_work ? (unsigned)by_work : 1;
    return t;
}

template <class F>
void run_parallel(unsigned nthreads, F &&body)
{
    std::vector<std::exception_ptr> err(nthreads);
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < nthreads; t++)
        pool.emplace_back([&, t] { try { body(t); } catch (...) { err[t] = std::current_exception(); } });
    try { body(0); } catch (...) { err[0] = std::current_exception(); }
    for (auto &th : pool) th.join();
    for (auto &e : err)
        if (e) std::rethrow_exception(e);
}

// COO (0-based, already mirrored for symmetric files) -> CSR with sorted, duplicate-free rows.
// Self-loops are dropped and duplicates keep the smallest weight: the algorithms assume a
// loop-free simple graph (LAGraph_cdlp.c: "assume ... no self edges").  Counting, scattering and the
// per-row sort + dedupe run on all host threads.
HostMatrix coo_to_csr(GrB_Index n, std::vector<uint32_t> &src, std::vector<uint32_t> &dst, std::vector<double> &val, bool weighted)
{
    HostMatrix A;
    A.nrows = n;
    A.iso = !weighted;
    const size_t nz = src.size();
    const unsigned T = loader_threads(nz, 1u << 18);
    std::vector<GrB_Index> cnt(n + 1, 0);
    run_parallel(T, [&](unsigned t) {
        for (size_t k = nz * t / T; k < nz * (t + 1) / T; k++)
            if (src[k] != dst[k]) __atomic_fetch_add(&cnt[src[k] + 1], 1, __ATOMIC_RELAXED);
    });
    for (GrB_Index i = 0; i < n; i++) cnt[i + 1] += cnt[i];
    std::vector<uint32_t> col(cnt[n]);
    std::vector<double> w(weighted ? cnt[n] : 0);
    {
        std::vector<GrB_Index> cur(cnt.begin(), cnt.end() - 1);
        run_parallel(T, [&](unsigned t) {
            for (size_t k = nz * t / T; k < nz * (t + 1) / T; k++) {
                if (src[k] == dst[k]) continue;
                const GrB_Index p = __atomic_fetch_add(&cur[src[k]], 1, __ATOMIC_RELAXED); // any order: rows are sorted below
                col[p] = dst[k];
                if (weighted) w[p] = val[k];
            }
        });
    }
    { std::vector<uint32_t>().swap(src); std::vector<uint32_t>().swap(dst); std::vector<double>().swap(val); }
    // rows sorted by (column, weight) and deduplicated in place, in blocks of rows drawn from a counter
    std::vector<GrB_Index> kept(n + 1, 0);
    const unsigned T2 = loader_threads((size_t)n, 1u << 12);
    std::atomic<GrB_Index> next_block{0};
    const GrB_Index BLOCK = 4096;
    run_parallel(T2, [&](unsigned) {
        std::vector<std::pair<uint32_t, double>> tmp;
        for (;;) {
            const GrB_Index r0 = next_block.fetch_add(BLOCK);
            if (r0 >= n) break;
            const GrB_Index r1 = std::min<GrB_Index>(n, r0 + BLOCK);
            for (GrB_Index i = r0; i < r1; i++) {
                const GrB_Index a = cnt[i], b = cnt[i + 1];
                GrB_Index out = a;
                if (weighted) {
                    tmp.resize(b - a);
                    for (GrB_Index k = a; k < b; k++) tmp[k - a] = {col[k], w[k]};
                    std::sort(tmp.begin(), tmp.end());
                    for (size_t k = 0; k < tmp.size(); k++) {
                        if (k && tmp[k].first == tmp[k - 1].first) continue; // the smallest weight came first
                        col[out] = tmp[k].first;
                        w[out++] = tmp[k].second;
                    }
                } else {
                    std::sort(col.begin() + a, col.begin() + b);
                    for (GrB_Index k = a; k < b; k++) {
                        if (k > a && col[k] == col[k - 1]) continue;
                        col[out++] = col[k];
                    }
                }
                kept[i + 1] = out - a;
            }
        }
    });
    A.Ap.assign(n + 1, 0);
    for (GrB_Index i = 0; i < n; i++) A.Ap[i + 1] = A.Ap[i] + kept[i + 1];
    A.nvals = A.Ap[n];
    A.Aj.resize(A.nvals);
    if (weighted) A.Ax.resize(A.nvals);
    run_parallel(T2, [&](unsigned t) {
        for (GrB_Index i = n * t / T2; i < n * (t + 1) / T2; i++) {
            const GrB_Index len = kept[i + 1];
            std::copy(col.begin() + cnt[i], col.begin() + cnt[i] + len, A.Aj.begin() + A.Ap[i]);
            if (weighted) std::copy(w.begin() + cnt[i], w.begin() + cnt[i] + len, A.Ax.begin() + A.Ap[i]);
        }
    });
    return A;
}

} // namespace

// ------------------------------------------------------------------------- .mtx
// Format as written by bin/py/relabel.py:64-79: banner `%%MatrixMarket matrix coordinate
// integer|real general|symmetric`, `%%GraphBLAS GrB_BOOL|GrB_FP64`, `n n nnz`, then 1-based
// `src dst val` in no particular order.  Symmetric files list each edge once in either
// triangle; every off-diagonal entry is mirrored (what LAGraph_MMRead does).
HostMatrix ReadMtxFile(const std::string &path)
{
    FileBytes file(path);
    const char *p = file.data.data();
    if (std::strncmp(p, "%%MatrixMarket", 14) != 0) throw std::runtime_error("Not a Matrix Market file: " + path);
    std::string banner(p, next_line(p) - p);
    std::transform(banner.begin(), banner.end(), banner.begin(), [](unsigned char c) { return (char)std::tolower(c); });
    if (banner.find("coordinate") == std::string::npos) throw std::runtime_error("Only coordinate Matrix Market files are supported");
    const bool symmetric = banner.find("symmetric") != std::string::npos;
    const bool pattern = banner.find("pattern") != std::string::npos;
    const bool weighted = banner.find("real") != std::string::npos || banner.find("double") != std::string::npos;
    p = next_line(p);
    while (*p == '%') p = next_line(p);
    const GrB_Index nrows = parse_u64(p), ncols = parse_u64(p), nnz = parse_u64(p);
    if (nrows != ncols) throw std::runtime_error("Adjacency matrix must be square");
    if (nrows >= 0xFFFFFFFEull) throw std::runtime_error("More than 2^32 - 2 vertices are not supported");
    p = next_line(p);
    // The body is cut into byte ranges that start at line starts; every thread parses its range into
    // vectors of its own, which are then copied into place in file order.
    const char *body = p, *end = file.data.data() + file.data.size() - 1; // (the trailing NUL is not part of the text)
    const unsigned T = loader_threads((size_t)(end - body), 4u << 20);
    std::vector<const char *> cut(T + 1);
    cut[0] = body;
    cut[T] = end;
    for (unsigned t = 1; t < T; t++) {
        const char *q = body + (size_t)(end - body) * t / T;
        while (q < end && *q != '\n') q++;
        cut[t] = q < end ? q + 1 : end;
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    struct Part { std::vector<uint32_t> src, dst; std::vector<double> val; size_t lines = 0; };
    std::vector<Part> part(T);
    run_parallel(T, [&](unsigned t) {
        Part &o = part[t];
        const char *q = cut[t], *stop = cut[t + 1];
        const size_t guess = (size_t)(stop - q) / 12 + 16;
        o.src.reserve(symmetric ? 2 * guess : guess);
        o.dst.reserve(symmetric ? 2 * guess : guess);
        if (weighted) o.val.reserve(symmetric ? 2 * guess : guess);
        for (;;) {
            while (q < stop && (*q == '\n' || *q == '\r' || *q == ' ')) q++;
            if (q >= stop) break;
            const GrB_Index i = parse_u64(q), j = parse_u64(q);
            if (i < 1 || j < 1 || i > nrows || j > nrows) throw std::runtime_error("Matrix Market entry out of range");
            double x = 1.0;
            if (!pattern) {
                q = skip_ws(q);
                char *e = nullptr;
                x = std::strtod(q, &e);
                if (e == q) throw std::runtime_error("Matrix Market entry without a value");
                q = e;
            }
            q = next_line(q);
            o.lines++;
            o.src.push_back((uint32_t)(i - 1));
            o.dst.push_back((uint32_t)(j - 1));
            if (weighted) o.val.push_back(x);
            if (symmetric && i != j) {
                o.src.push_back((uint32_t)(j - 1));
                o.dst.push_back((uint32_t)(i - 1));
                if (weighted) o.val.push_back(x);
            }
        }
    });
    size_t lines = 0, total = 0;
    std::vector<size_t> off(T + 1, 0);
    for (unsigned t = 0; t < T; t++) { lines += part[t].lines; total += part[t].src.size(); off[t + 1] = total; }
    if (lines < nnz) throw std::runtime_error("Matrix Market file ends before nnz entries were read");
    if (lines > nnz) throw std::runtime_error("Matrix Market file holds more entries than its size line announces");
    std::vector<uint32_t> src(total), dst(total);
    std::vector<double> val(weighted ? total : 0);
    run_parallel(T, [&](unsigned t) {
        std::copy(part[t].src.begin(), part[t].src.end(), src.begin() + off[t]);
        std::copy(part[t].dst.begin(), part[t].dst.end(), dst.begin() + off[t]);
        if (weighted) std::copy(part[t].val.begin(), part[t].val.end(), val.begin() + off[t]);
        Part().src.swap(part[t].src); Part().dst.swap(part[t].dst); Part().val.swap(part[t].val);
    });
    return coo_to_csr(nrows, src, dst, val, weighted);
}

// ------------------------------------------------------------------------- .grb
// SuiteSparse dump layout (graphio.h:88-105, 200-216, 549-606): 512-byte ASCII header, then
// fmt:int32 kind:int32 hyper:f64 nrows,ncols:u64 nonempty:i64 nvec,nvals:u64 typecode:int32
// typesize:size_t, then Ap[nvec+1] (Ah[nvec] if hypersparse) Ai[nvals] as uint64, then Ax
// (one value if iso).  Only the sparse / hypersparse by-row forms are accepted.
namespace {
template <class T>
void rd(FILE *f, T *dst, size_t count, const std::string &path)
{
    if (count && fread(dst, sizeof(T), count, f) != count) throw std::runtime_error("Truncated binary matrix file: " + path);
}
template <class T>
void wr(FILE *f, const T *src, size_t count)
{
    if (count && fwrite(src, sizeof(T), count, f) != count) throw std::runtime_error("Write failed");
}
} // namespace

HostMatrix ReadGrbFile(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open binary matrix file: " + path);
    HostMatrix A;
    try {
        char header[GRB_HEADER_LEN];
        rd(f, header, GRB_HEADER_LEN, path);
        int32_t fmt = 0, kind = 0, typecode = 0;
        double hyper = 0;
        uint64_t nrows = 0, ncols = 0, nvec = 0, nvals = 0, typesize = 0;
        int64_t nonempty = 0;
        rd(f, &fmt, 1, path); rd(f, &kind, 1, path); rd(f, &hyper, 1, path);
        rd(f, &nrows, 1, path); rd(f, &ncols, 1, path); rd(f, &nonempty, 1, path);
        rd(f, &nvec, 1, path); rd(f, &nvals, 1, path); rd(f, &typecode, 1, path); rd(f, &typesize, 1, path);
        bool iso = false;
        if (kind > 100) { iso = true; kind -= 100; }
        const bool is_hyper = kind == 1, is_sparse = kind == 0 || kind == 2;
        if (fmt != 0 || !(is_hyper || is_sparse)) throw std::runtime_error("Only by-row sparse .grb matrices are supported: " + path);
        if (nrows != ncols) throw std::runtime_error("Adjacency matrix must be square");
        if (nrows >= 0xFFFFFFFEull) throw std::runtime_error("More than 2^32 - 2 vertices are not supported");
        // the sizes come from the file: bound them by what the file can hold before allocating
        {
            const long here = ftell(f);
            fseek(f, 0, SEEK_END);
            const long fsz = ftell(f);
            fseek(f, here, SEEK_SET);
            const uint64_t left = fsz > here ? (uint64_t)(fsz - here) : 0;
            if (nvec > nrows || nvec + 1 > left / 8 || nvals > left / 8 || typesize > 1024)
                throw std::runtime_error("Corrupt .grb header (sizes exceed the file): " + path);
        }
        std::vector<GrB_Index> Ap(nvec + 1), Ah, Ai(nvals);
        rd(f, Ap.data(), nvec + 1, path);
        if (is_hyper) { Ah.resize(nvec); rd(f, Ah.data(), nvec, path); }
        rd(f, Ai.data(), nvals, path);
        if (Ap[0] != 0 || Ap[nvec] != nvals) throw std::runtime_error("Corrupt .grb offsets (Ap[0] != 0 or Ap[nvec] != nvals): " + path);
        for (uint64_t k = 0; k < nvec; k++)
            if (Ap[k] > Ap[k + 1]) throw std::runtime_error("Corrupt .grb offsets (not monotone): " + path);
        for (uint64_t k = 0; k < nvec && is_hyper; k++)
            if (Ah[k] >= nrows || (k && Ah[k] <= Ah[k - 1])) throw std::runtime_error("Corrupt .grb row list: " + path);
        // values: FP64 weights are kept; any other non-iso type (bool / integer dumps of the reference's binwrite)
        // is skipped and the matrix treated as structural, like an iso one
        const bool weighted = !iso && typecode == 10 && typesize == 8;
        std::vector<double> Ax;
        if (weighted) { Ax.resize(nvals); rd(f, Ax.data(), nvals, path); }
        // same policy as the .mtx path (coo_to_csr): self-loops dropped, rows sorted, duplicates merged
        // (smallest weight kept), so both loaders give the same graph for the same data.  A dump that is already
        // clean -- what the converter writes -- is adopted as it is after one parallel check.
        if (!is_hyper && nvec != nrows) throw std::runtime_error("Sparse .grb with nvec != nrows: " + path);
        const unsigned T = loader_threads((size_t)nvec, 1u << 14);
        std::atomic<int> dirty{0}, out_of_range{0};
        run_parallel(T, [&](unsigned t) {
            for (uint64_t k = nvec * t / T; k < nvec * (t + 1) / T; k++) {
                const uint64_t row = is_hyper ? Ah[k] : k;
                for (uint64_t e = Ap[k]; e < Ap[k + 1]; e++) {
                    if (Ai[e] >= nrows) { out_of_range = 1; return; }
                    if (Ai[e] == row || (e > Ap[k] && Ai[e] <= Ai[e - 1])) dirty = 1;
                }
            }
        });
        if (out_of_range) throw std::runtime_error("Column index out of range in " + path);
        if (!dirty && !is_hyper) {
            A.nrows = nrows;
            A.nvals = nvals;
            A.iso = !weighted;
            A.Aj.resize(nvals);
            run_parallel(T, [&](unsigned t) {
                for (uint64_t e = nvals * t / T; e < nvals * (t + 1) / T; e++) A.Aj[e] = (uint32_t)Ai[e];
            });
            A.Ap = std::move(Ap);
            A.Ax = std::move(Ax);
        } else {
            std::vector<uint32_t> src(nvals), dst(nvals);
            run_parallel(T, [&](unsigned t) {
                for (uint64_t k = nvec * t / T; k < nvec * (t + 1) / T; k++) {
                    const uint32_t row = (uint32_t)(is_hyper ? Ah[k] : k);
                    for (uint64_t e = Ap[k]; e < Ap[k + 1]; e++) { src[e] = row; dst[e] = (uint32_t)Ai[e]; }
                }
            });
            { std::vector<GrB_Index>().swap(Ai); std::vector<GrB_Index>().swap(Ap); }
            A = coo_to_csr(nrows, src, dst, Ax, weighted);
        }
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
    return A;
}

void WriteGrbFile(const std::string &path, const HostMatrix &A)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot create binary matrix file: " + path);
    try {
        char header[GRB_HEADER_LEN];
        const uint64_t typesize = A.iso ? 1 : 8;
        int len = snprintf(header, GRB_HEADER_LEN,
                           "SuiteSparse:GraphBLAS matrix\nv%-25s\nnrows:  %-18" PRIu64 "\nncols:  %-18" PRIu64
                           "\nnvec:   %-18" PRIu64 "\nnvals:  %-18" PRIu64 "\nformat: %-8s\nsize:   %-18" PRIu64
                           "\ntype:   %-72s\niso:    %1d\n%-210s\n\n",
                           "7.4.4 (gxb200 converter)", A.nrows, A.nrows, A.nrows, A.nvals, "CSR ", typesize,
                           A.iso ? "bool" : "double", A.iso ? 1 : 0, "\n");
        if (len < 0) len = 0;
        for (size_t k = (size_t)len; k < GRB_HEADER_LEN; k++) header[k] = ' ';
        header[GRB_HEADER_LEN - 1] = '\0';
        wr(f, header, GRB_HEADER_LEN);
        const int32_t fmt = 0, kind = 2 + (A.iso ? 100 : 0), typecode = A.iso ? 0 : 10;
        const double hyper = 0.0625;
        const int64_t nonempty = -1;
        const uint64_t n = A.nrows;
        wr(f, &fmt, 1); wr(f, &kind, 1); wr(f, &hyper, 1); wr(f, &n, 1); wr(f, &n, 1); wr(f, &nonempty, 1);
        wr(f, &n, 1); wr(f, &A.nvals, 1); wr(f, &typecode, 1); wr(f, &typesize, 1);
        wr(f, A.Ap.data(), A.Ap.size());
        std::vector<GrB_Index> Ai(A.Aj.begin(), A.Aj.end());
        wr(f, Ai.data(), Ai.size());
        if (A.iso) { const uint8_t one = 1; wr(f, &one, 1); }
        else wr(f, A.Ax.data(), A.Ax.size());
    } catch (...) {
        fclose(f);
        throw;
    }
    fclose(f);
}

// ------------------------------------------------------------------------- mappings
std::vector<GrB_Index> ReadVtxFile(const std::string &path)
{
    FileBytes file(path);
    std::vector<GrB_Index> mapping;
    const char *p = file.data.data();
    for (;;) {
        while (*p == '\n' || *p == '\r' || *p == ' ') p++;
        if (!*p) break;
        mapping.push_back(parse_u64(p));
        p = next_line(p);
    }
    return mapping;
}

std::vector<GrB_Index> ReadVtbFile(const std::string &path)
{
    FILE *f = fopen(path.c_str(), "rb");
    if (!f) throw std::runtime_error("Cannot open binary mapping file: " + path);
    fseek(f, 0, SEEK_END);
    const long sz = ftell(f);
    fseek(f, 0, SEEK_SET);
    std::vector<GrB_Index> mapping((size_t)sz / sizeof(GrB_Index));
    const size_t got = mapping.empty() ? 0 : fread(mapping.data(), sizeof(GrB_Index), mapping.size(), f);
    fclose(f);
    if (got != mapping.size()) throw std::runtime_error("Short read: " + path);
    return mapping;
}

void WriteVtbFile(const std::string &path, const std::vector<GrB_Index> &mapping)
{
    FILE *f = fopen(path.c_str(), "wb");
    if (!f) throw std::runtime_error("Cannot create binary mapping file: " + path);
    const size_t put = mapping.empty() ? 0 : fwrite(mapping.data(), sizeof(GrB_Index), mapping.size(), f);
    fclose(f);
    if (put != mapping.size()) throw std::runtime_error("Write failed: " + path);
}

HostMatrix ReadMatrixMarket(const BenchmarkParameters &parameters)
{
    ComputationTimer total_timer{"Loading the matrix"};
    if (parameters.binary) return ReadGrbFile(parameters.input_dir + "/graph.grb");
    return ReadMtxFile(parameters.input_dir + "/graph.mtx");
}

std::vector<GrB_Index> ReadMapping(const BenchmarkParameters &parameters)
{
    ComputationTimer timer{"Loading the mapping"};
    if (parameters.binary) return ReadVtbFile(parameters.input_dir + "/graph.vtb");
    return ReadVtxFile(parameters.input_dir + "/graph.vtx");
}

// ------------------------------------------------------------------------- result files
ResultWriter::ResultWriter(const std::string &path) : buf_(1 << 22)
{
    f_ = fopen(path.c_str(), "w");
}

ResultWriter::~ResultWriter()
{
    if (f_) { flush(); fclose(f_); }
}

void ResultWriter::flush()
{
    if (used_) fwrite(buf_.data(), 1, used_, f_);
    used_ = 0;
}

void ResultWriter::line_int(GrB_Index id, int64_t v)
{
    if (used_ + 64 > buf_.size()) flush();
    used_ += (size_t)snprintf(buf_.data() + used_, 64, "%" PRIu64 " %" PRId64 "\n", id, v);
}

void ResultWriter::line_uint(GrB_Index id, uint64_t v)
{
    if (used_ + 64 > buf_.size()) flush();
    used_ += (size_t)snprintf(buf_.data() + used_, 64, "%" PRIu64 " %" PRIu64 "\n", id, v);
}

void ResultWriter::line_sci(GrB_Index id, double v)
{
    if (used_ + 64 > buf_.size()) flush();
    used_ += (size_t)snprintf(buf_.data() + used_, 64, "%" PRIu64 " %.16e\n", id, v);
}

// decimal digits of v at dst (no terminator); returns the length
static inline size_t put_u64(char *dst, uint64_t v)
{
    char tmp[24];
    size_t k = 0;
    do { tmp[k++] = (char)('0' + v % 10); v /= 10; } while (v);
    for (size_t i = 0; i < k; i++) dst[i] = tmp[k - 1 - i];
    return k;
}

template <class F>
void ResultWriter::lines_parallel(GrB_Index n, F &&one)
{
    flush();
    const GrB_Index SUPER = GrB_Index(1) << 23; // lines per round: bounds the buffers to ~0.5 GB whatever n is
    const unsigned T = loader_threads((size_t)std::min<GrB_Index>(n, SUPER), 1u << 14);
    std::vector<std::vector<char>> out(T);
    for (GrB_Index base = 0; base < n; base += SUPER) {
        const GrB_Index cnt = std::min<GrB_Index>(SUPER, n - base);
        run_parallel(T, [&](unsigned t) {
            const GrB_Index a = base + cnt * t / T, b = base + cnt * (t + 1) / T;
            std::vector<char> &o = out[t];
            o.resize((size_t)(b - a) * 64); // a line is at most 20 + 1 + 24 + 1 bytes
            size_t used = 0;
            for (GrB_Index i = a; i < b; i++) used += one(o.data() + used, i);
            o.resize(used);
        });
        for (unsigned t = 0; t < T; t++)
            if (!out[t].empty()) fwrite(out[t].data(), 1, out[t].size(), f_);
    }
}

void ResultWriter::lines_int(const GrB_Index *ids, const int64_t *vals, GrB_Index n)
{
    lines_parallel(n, [&](char *dst, GrB_Index i) {
        size_t k = put_u64(dst, ids[i]);
        dst[k++] = ' ';
        int64_t v = vals[i];
        if (v < 0) { dst[k++] = '-'; k += put_u64(dst + k, (uint64_t)0 - (uint64_t)v); }
        else k += put_u64(dst + k, (uint64_t)v);
        dst[k++] = '\n';
        return k;
    });
}

void ResultWriter::lines_uint(const GrB_Index *ids, const uint64_t *vals, GrB_Index n, const GrB_Index *map)
{
    lines_parallel(n, [&](char *dst, GrB_Index i) {
        size_t k = put_u64(dst, ids[i]);
        dst[k++] = ' ';
        k += put_u64(dst + k, map ? map[vals[i]] : vals[i]);
        dst[k++] = '\n';
        return k;
    });
}

void ResultWriter::lines_ids(const GrB_Index *ids, GrB_Index n)
{
    lines_parallel(n, [&](char *dst, GrB_Index i) {
        size_t k = put_u64(dst, ids[i]);
        dst[k++] = '\n';
        return k;
    });
}

void ResultWriter::lines_sci(const GrB_Index *ids, const double *vals, GrB_Index n)
{
    lines_parallel(n, [&](char *dst, GrB_Index i) {
        size_t k = put_u64(dst, ids[i]);
        dst[k++] = ' ';
        const double v = vals[i];
        if (v == HUGE_VAL) { memcpy(dst + k, "infinity", 8); k += 8; }
        else k += (size_t)snprintf(dst + k, 32, "%.16e", v);
        dst[k++] = '\n';
        return k;
    });
}

void ResultWriter::line_text(GrB_Index id, const char *s)
{
    if (used_ + 64 > buf_.size()) flush();
    used_ += (size_t)snprintf(buf_.data() + used_, 64, "%" PRIu64 " %s\n", id, s);
}

// ------------------------------------------------------------------------- relabel (.v/.e -> graph.vtx/graph.mtx)
// The load stage the reference runs through DuckDB (bin/py/relabel.py:8-79, bin/sh/load-graph.sh:49-60): vertex ids
// of X.v become dense ids in .v row order, every edge of X.e is rewritten with 1-based dense endpoints, and the two
// text files the converter / the loaders read are written.  Everything -- parsing both files in byte ranges, the
// id lookup, formatting -- runs on all host threads; edge weights are carried as the text they came as, so they
// round-trip byte for byte.
namespace {

struct Span { const char *p; uint32_t len; };

// lines of [begin, end) -> callback(line_begin, line_end) for every non-empty line; ranges start at line starts
template <class F>
void for_each_line(const char *q, const char *stop, F &&f)
{
    while (q < stop) {
        const char *e = q;
        while (e < stop && *e != '\n') e++;
        const char *le = e;
        while (le > q && (le[-1] == '\r' || le[-1] == ' ' || le[-1] == '\t')) le--;
        if (le > q) f(q, le);
        q = e < stop ? e + 1 : stop;
    }
}

std::vector<const char *> line_cuts(const char *body, const char *end, unsigned T)
{
    std::vector<const char *> cut(T + 1);
    cut[0] = body;
    cut[T] = end;
    for (unsigned t = 1; t < T; t++) {
        const char *q = body + (size_t)(end - body) * t / T;
        while (q < end && *q != '\n') q++;
        cut[t] = q < end ? q + 1 : end;
        if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
    }
    return cut;
}

} // namespace

void RelabelGraph(const std::string &vertex_path, const std::string &edge_path, const std::string &out_dir, bool weighted,
                  bool directed, GrB_Index *n_out, GrB_Index *nnz_out)
{
    // ---- X.v: one original id per line
    std::vector<GrB_Index> ids;
    {
        FileBytes vf(vertex_path);
        const char *body = vf.data.data(), *end = body + vf.data.size() - 1;
        const unsigned T = loader_threads((size_t)(end - body), 1u << 20);
        const std::vector<const char *> cut = line_cuts(body, end, T);
        std::vector<std::vector<GrB_Index>> part(T);
        run_parallel(T, [&](unsigned t) {
            part[t].reserve((size_t)(cut[t + 1] - cut[t]) / 8 + 16);
            for_each_line(cut[t], cut[t + 1], [&](const char *a, const char *) { const char *q = a; part[t].push_back(parse_u64(q)); });
        });
        size_t total = 0;
        std::vector<size_t> off(T + 1, 0);
        for (unsigned t = 0; t < T; t++) { total += part[t].size(); off[t + 1] = total; }
        ids.resize(total);
        run_parallel(T, [&](unsigned t) { std::copy(part[t].begin(), part[t].end(), ids.begin() + off[t]); });
    }
    const GrB_Index n = ids.size();
    if (n >= 0xFFFFFFFEull) throw std::runtime_error("More than 2^32 - 2 vertices are not supported");
    // lookup id -> row: the ids themselves when the file is ascending (Graphalytics' .v files are), else a sorted copy
    bool ascending = true;
    for (GrB_Index i = 1; i < n && ascending; i++) ascending = ids[i] > ids[i - 1];
    std::vector<GrB_Index> sorted_ids;
    std::vector<uint32_t> sorted_row;
    if (!ascending) {
        sorted_row.resize(n);
        std::iota(sorted_row.begin(), sorted_row.end(), 0u);
        std::sort(sorted_row.begin(), sorted_row.end(), [&](uint32_t a, uint32_t b) { return ids[a] < ids[b]; });
        sorted_ids.resize(n);
        for (GrB_Index i = 0; i < n; i++) sorted_ids[i] = ids[sorted_row[i]];
    }
    const std::vector<GrB_Index> &keys = ascending ? ids : sorted_ids;
    auto dense = [&](GrB_Index id) -> uint32_t {
        const auto it = std::lower_bound(keys.begin(), keys.end(), id);
        if (it == keys.end() || *it != id) throw std::runtime_error("edge endpoint " + std::to_string(id) + " is not in the vertex file");
        const size_t pos = (size_t)(it - keys.begin());
        return ascending ? (uint32_t)pos : sorted_row[pos];
    };
    // ---- X.e: `src dst[ weight]`, parsed and relabelled in one pass
    FileBytes ef(edge_path);
    const char *ebody = ef.data.data(), *eend = ebody + ef.data.size() - 1;
    const unsigned T = loader_threads((size_t)(eend - ebody), 4u << 20);
    const std::vector<const char *> cut = line_cuts(ebody, eend, T);
    struct Part { std::vector<uint32_t> src, dst; std::vector<Span> w; };
    std::vector<Part> part(T);
    run_parallel(T, [&](unsigned t) {
        Part &o = part[t];
        const size_t guess = (size_t)(cut[t + 1] - cut[t]) / 12 + 16;
        o.src.reserve(guess);
        o.dst.reserve(guess);
        if (weighted) o.w.reserve(guess);
        for_each_line(cut[t], cut[t + 1], [&](const char *a, const char *le) {
            const char *q = a;
            const GrB_Index s = parse_u64(q), d = parse_u64(q);
            o.src.push_back(dense(s));
            o.dst.push_back(dense(d));
            if (weighted) {
                q = skip_ws(q);
                if (q >= le) throw std::runtime_error("edge without a weight in a weighted edge file");
                o.w.push_back(Span{q, (uint32_t)(le - q)});
            }
        });
    });
    GrB_Index nnz = 0;
    for (unsigned t = 0; t < T; t++) nnz += part[t].src.size();
    // ---- graph.vtx: the ids in .v row order
    {
        ResultWriter vtx(out_dir + "/graph.vtx");
        if (!vtx.ok()) throw std::runtime_error("Cannot create " + out_dir + "/graph.vtx");
        vtx.lines_ids(ids.data(), n);
    }
    // ---- graph.mtx (relabel.py:64-79): banner, GraphBLAS type line, `n n nnz`, then 1-based `src dst val` in .e order
    FILE *f = fopen((out_dir + "/graph.mtx").c_str(), "w");
    if (!f) throw std::runtime_error("Cannot create " + out_dir + "/graph.mtx");
    fprintf(f, "%%%%MatrixMarket matrix coordinate %s %s\n%%%%GraphBLAS %s\n%" PRIu64 " %" PRIu64 " %" PRIu64 "\n",
            weighted ? "real" : "integer", directed ? "general" : "symmetric", weighted ? "GrB_FP64" : "GrB_BOOL", n, n, nnz);
    std::vector<std::vector<char>> out(T);
    std::exception_ptr err;
    try {
        run_parallel(T, [&](unsigned t) {
            const Part &o = part[t];
            std::vector<char> &b = out[t];
            size_t need = o.src.size() * 24;
            if (weighted) for (const Span &x : o.w) need += x.len;
            b.resize(need);
            size_t k = 0;
            for (size_t i = 0; i < o.src.size(); i++) {
                k += put_u64(b.data() + k, (uint64_t)o.src[i] + 1);
                b[k++] = ' ';
                k += put_u64(b.data() + k, (uint64_t)o.dst[i] + 1);
                b[k++] = ' ';
                if (weighted) { memcpy(b.data() + k, o.w[i].p, o.w[i].len); k += o.w[i].len; }
                else b[k++] = '1';
                b[k++] = '\n';
            }
            b.resize(k);
        });
        for (unsigned t = 0; t < T; t++)
            if (!out[t].empty() && fwrite(out[t].data(), 1, out[t].size(), f) != out[t].size()) throw std::runtime_error("Write failed: graph.mtx");
    } catch (...) {
        err = std::current_exception();
    }
    fclose(f);
    if (err) std::rethrow_exception(err);
    if (n_out) *n_out = n;
    if (nnz_out) *nnz_out = nnz;
}
