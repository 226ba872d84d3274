// graphio.h -- graph and mapping loaders of the per-algorithm binaries, the
// GraphBLAS-free counterpart of ReadMatrixMarket / ReadMapping
// (src/graphio.cpp:4-60) and binread / binwrite (include/graphio.h:49-685).
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "utils.h"

typedef uint64_t GrB_Index; // the reference's index type, kept at the boundary

// What ReadMatrixMarket returns instead of a GrB_Matrix: CSR by row over dense
// ids, the arrays GxB_Matrix_export_CSR would hand out (cdlp_cuda.cu:181).
struct HostMatrix {
    GrB_Index nrows = 0;
    GrB_Index nvals = 0;
    bool iso = true;                 // structural (GrB_BOOL iso) vs FP64 values
    std::vector<GrB_Index> Ap;       // nrows + 1
    std::vector<uint32_t> Aj;        // nvals, sorted inside each row
    std::vector<double> Ax;          // nvals when !iso
};

// --binary true -> <input-dir>/graph.grb, else <input-dir>/graph.mtx (graphio.cpp:10-24)
HostMatrix ReadMatrixMarket(const BenchmarkParameters &parameters);
// --binary true -> graph.vtb (raw uint64[]), else graph.vtx (one id per line) (graphio.cpp:34-60)
std::vector<GrB_Index> ReadMapping(const BenchmarkParameters &parameters);

HostMatrix ReadMtxFile(const std::string &path);
HostMatrix ReadGrbFile(const std::string &path);
void WriteGrbFile(const std::string &path, const HostMatrix &A);
std::vector<GrB_Index> ReadVtxFile(const std::string &path);
std::vector<GrB_Index> ReadVtbFile(const std::string &path);
void WriteVtbFile(const std::string &path, const std::vector<GrB_Index> &mapping);

// Buffered "<original id> <value>\n" writer shared by the six Serialize*Result functions.
class ResultWriter {
    FILE *f_ = nullptr;
    std::vector<char> buf_;
    size_t used_ = 0;
    void flush();

  public:
    explicit ResultWriter(const std::string &path);
    ResultWriter(const ResultWriter &) = delete;
    ResultWriter &operator=(const ResultWriter &) = delete;
    ResultWriter(ResultWriter &&o) noexcept : f_(o.f_), buf_(std::move(o.buf_)), used_(o.used_) { o.f_ = nullptr; o.used_ = 0; }
    ~ResultWriter();
    bool ok() const { return f_ != nullptr; }
    void line_int(GrB_Index id, int64_t v);
    void line_uint(GrB_Index id, uint64_t v);
    void line_sci(GrB_Index id, double v);      // precision(16) << scientific
    void line_text(GrB_Index id, const char *s);
};
