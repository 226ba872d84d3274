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

// X.v / X.e -> out_dir/graph.vtx + out_dir/graph.mtx (bin/py/relabel.py:8-79), on all host threads
void RelabelGraph(const std::string &vertex_path, const std::string &edge_path, const std::string &out_dir, bool weighted,
                  bool directed, GrB_Index *n_out, GrB_Index *nnz_out);

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
    // Whole-vector forms: the n lines are formatted on all host threads (a block of consecutive lines per thread into
    // a buffer of its own) and written out in order -- 33 M lines of "%.16e" text are seconds of snprintf on one core.
    // ids[i] is the vertex id of line i.  lines_uint: the value is map[vals[i]] when map is given (cdlp.cpp:48).
    // lines_sci: +inf is written as the literal `infinity` (sssp.cpp:41-46).
    void lines_int(const GrB_Index *ids, const int64_t *vals, GrB_Index n);
    void lines_uint(const GrB_Index *ids, const uint64_t *vals, GrB_Index n, const GrB_Index *map = nullptr);
    void lines_sci(const GrB_Index *ids, const double *vals, GrB_Index n);
    void lines_ids(const GrB_Index *ids, GrB_Index n); // one id per line (graph.vtx)

  private:
    template <class F>
    void lines_parallel(GrB_Index n, F &&one);
};
