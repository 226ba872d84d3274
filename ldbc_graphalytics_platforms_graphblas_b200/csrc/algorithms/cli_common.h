// cli_common.h -- the few steps all six binaries share (device binding, upload, output file).
#pragma once

#include <cstdlib>
#include <iostream>
#include <string>
#include <vector>

#include "computation_timer.hpp"
#include "graphio.h"
#include "utils.h"

// LAGraph_Init + GxB_Global_Option_set(NTHREADS) in the reference (bfs.cpp:88-89); here the
// process binds to one B200.  GX_DEVICE selects it (default 0).  No device => the job fails.
inline void InitDevice()
{
    const char *dev = std::getenv("GX_DEVICE");
    OK(gx_init(dev ? std::atoi(dev) : 0));
}

// Device memory for the job, sized from the graph (resource allocation, like the pool a GraphBLAS runtime sets up at
// LAGraph_Init): 4x the adjacency + 64 bytes per vertex, served back to the algorithms by the stream-ordered pool.
inline void ReserveForGraph(const HostMatrix &A)
{
    const uint64_t adj = 8 * (A.nrows + 1) + (A.iso ? 4 : 12) * A.nvals;
    OK(gx_reserve(4 * adj + 64 * A.nrows));
}

inline gx_graph *UploadGraph(const HostMatrix &A, bool directed, unsigned cache)
{
    ComputationTimer timer{"Uploading the matrix"};
    gx_graph *G = nullptr;
    OK(gx_graph_create_csr32_cached(&G, A.nrows, A.nvals, A.Ap.data(), A.Aj.data(), A.iso ? nullptr : A.Ax.data(), directed ? 1 : 0,
                                    cache));
    return G;
}

// ReadMatrixMarket + the upload in one step for the wrappers that need nothing else from the host matrix.
// graph.mtx is tokenised on the device (gx_graph_load_mtx: the host reads the header only) unless GX_LOADER=host;
// graph.grb (--binary true) is read on the host and uploaded.  The device pool is sized for the job either way.
struct DeviceGraph {
    gx_graph *G = nullptr;
    GrB_Index nrows = 0, nvals = 0;
    bool weighted = false;
};

inline DeviceGraph LoadGraph(const BenchmarkParameters &parameters, unsigned cache)
{
    DeviceGraph D;
    const char *le = std::getenv("GX_LOADER");
    if (!parameters.binary && !(le && std::string(le) == "host")) {
        ComputationTimer timer{"Loading the matrix"};
        OK(gx_graph_load_mtx(&D.G, (parameters.input_dir + "/graph.mtx").c_str(), parameters.directed ? 1 : 0, cache));
        uint64_t n = 0, nnz = 0;
        int weighted = 0;
        OK(gx_graph_info(D.G, &n, &nnz, nullptr, &weighted));
        D.nrows = n; D.nvals = nnz; D.weighted = weighted != 0;
        OK(gx_reserve(4 * (8 * (n + 1) + (weighted ? 12 : 4) * nnz) + 64 * n));
        return D;
    }
    HostMatrix A = ReadMatrixMarket(parameters);
    ReserveForGraph(A);
    D.G = UploadGraph(A, parameters.directed, cache);
    D.nrows = A.nrows; D.nvals = A.nvals; D.weighted = !A.iso;
    return D;
}

// Result vector in pinned host memory, allocated before the timed window: the download inside the window then runs
// at PCIe speed instead of through a pageable staging copy (19 MB of levels: 0.4 ms instead of ~5 ms).
template <class T>
class PinnedVector {
    T *p_ = nullptr;
    size_t n_ = 0;

  public:
    explicit PinnedVector(size_t n) : n_(n) { void *q = nullptr; OK(gx_host_alloc(&q, n * sizeof(T))); p_ = (T *)q; }
    PinnedVector(const PinnedVector &) = delete;
    PinnedVector &operator=(const PinnedVector &) = delete;
    PinnedVector(PinnedVector &&o) noexcept : p_(o.p_), n_(o.n_) { o.p_ = nullptr; o.n_ = 0; }
    ~PinnedVector() { if (p_) gx_host_free(p_); }
    T *data() { return p_; }
    const T *data() const { return p_; }
    size_t size() const { return n_; }
    T &operator[](size_t i) { return p_[i]; }
    const T &operator[](size_t i) const { return p_[i]; }
};

inline ResultWriter OpenOutput(const BenchmarkParameters &parameters)
{
    ResultWriter file(parameters.output_file);
    if (!file.ok()) {
        std::cerr << "Output file " << parameters.output_file << " does not exists" << std::endl;
        exit(-1);
    }
    return file;
}
