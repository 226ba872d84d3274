// cdlp -- per-algorithm binary `bin/exe/cdlp` (execute-job.sh:105-115), the drop-in for
// src/algorithms/cdlp.cpp:83-108.  The reference's MY_CDLP_GPU prints the Processing lines
// around cdlp_gpu (cdlp_cuda.cu:241-243), i.e. its timed window holds cudaMalloc, the H2D
// copy of the CSR, the kernels and the D2H copy; this one does the same.
#include <iostream>

#include "cli_common.h"

void SerializeCDLPResult(const PinnedVector<uint64_t> &label, const std::vector<GrB_Index> &mapping,
                         const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    // labels are dense ids; the file carries original ids (cdlp.cpp:48)
    file.lines_uint(mapping.data(), label.data(), mapping.size(), mapping.data());
}

void MY_CDLP_GPU(const HostMatrix &A, bool symmetric, int itermax, PinnedVector<uint64_t> &label)
{
    ComputationTimer timer{"CDLP"};
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    gx_graph *G = UploadGraph(A, !symmetric, 0);
    OK(gx_cdlp(G, itermax, label.data()));
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;
    OK(gx_graph_free(G));
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    HostMatrix A = ReadMatrixMarket(parameters);
    std::vector<GrB_Index> mapping = ReadMapping(parameters);
    ReserveForGraph(A);
    // the upload is part of this algorithm's window (as in the reference): page-lock the loaded arrays beforehand
    OK(gx_host_register(A.Ap.data(), A.Ap.size() * sizeof(GrB_Index)));
    OK(gx_host_register(A.Aj.data(), A.Aj.size() * sizeof(uint32_t)));
    PinnedVector<uint64_t> result(A.nrows);
    MY_CDLP_GPU(A, !parameters.directed, parameters.max_iteration, result);
    OK(gx_host_unregister(A.Aj.data()));
    OK(gx_host_unregister(A.Ap.data()));
    SerializeCDLPResult(result, mapping, parameters);
    return 0;
}
