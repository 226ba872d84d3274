// cli_common.h -- the few steps all six binaries share (device binding, upload, output file).
#pragma once

#include <cstdlib>
#include <iostream>
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

inline gx_graph *UploadGraph(const HostMatrix &A, bool directed, unsigned cache)
{
    ComputationTimer timer{"Uploading the matrix"};
    gx_graph *G = nullptr;
    OK(gx_graph_create_csr32_cached(&G, A.nrows, A.nvals, A.Ap.data(), A.Aj.data(), A.iso ? nullptr : A.Ax.data(), directed ? 1 : 0,
                                    cache));
    return G;
}

inline ResultWriter OpenOutput(const BenchmarkParameters &parameters)
{
    ResultWriter file(parameters.output_file);
    if (!file.ok()) {
        std::cerr << "Output file " << parameters.output_file << " does not exists" << std::endl;
        exit(-1);
    }
    return file;
}
