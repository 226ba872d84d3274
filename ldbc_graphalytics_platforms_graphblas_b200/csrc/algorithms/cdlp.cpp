// cdlp -- per-algorithm binary `bin/exe/cdlp` (execute-job.sh:105-115), the drop-in for
// src/algorithms/cdlp.cpp:83-108.  The reference's MY_CDLP_GPU prints the Processing lines
// around cdlp_gpu (cdlp_cuda.cu:241-243), i.e. its timed window holds cudaMalloc, the H2D
// copy of the CSR, the kernels and the D2H copy; this one does the same.
#include <iostream>

#include "cli_common.h"

void SerializeCDLPResult(const std::vector<uint64_t> &label, const std::vector<GrB_Index> &mapping,
                         const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    // labels are dense ids; the file carries original ids (cdlp.cpp:48)
    for (GrB_Index v = 0; v < mapping.size(); v++) file.line_uint(mapping[v], mapping[label[v]]);
}

std::vector<uint64_t> MY_CDLP_GPU(const HostMatrix &A, bool symmetric, int itermax)
{
    ComputationTimer timer{"CDLP"};
    std::vector<uint64_t> label(A.nrows);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    gx_graph *G = UploadGraph(A, !symmetric, 0);
    OK(gx_cdlp(G, itermax, label.data()));
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;
    OK(gx_graph_free(G));
    return label;
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    HostMatrix A = ReadMatrixMarket(parameters);
    std::vector<GrB_Index> mapping = ReadMapping(parameters);
    std::vector<uint64_t> result = MY_CDLP_GPU(A, !parameters.directed, parameters.max_iteration);
    SerializeCDLPResult(result, mapping, parameters);
    return 0;
}
