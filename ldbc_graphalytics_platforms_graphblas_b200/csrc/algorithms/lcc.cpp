// lcc -- per-algorithm binary `bin/exe/lcc` (execute-job.sh:117-126), the drop-in for
// src/algorithms/lcc.cpp:73-90 with LAGraph_lcc replaced by gx_lcc.  The reference times all
// of LAGraph_lcc, including building A v A' (lcc.cpp:82-84); so does this one.
#include <iostream>

#include "cli_common.h"

void SerializeLCCResult(const PinnedVector<double> &lcc, const std::vector<GrB_Index> &mapping,
                        const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    // vertices LAGraph leaves without an entry are written as 0.0 (lcc.cpp:49-54); gx_lcc returns 0.0 for them
    file.lines_sci(mapping.data(), lcc.data(), mapping.size());
}

void LA_LCC(gx_graph *G, PinnedVector<double> &lcc)
{
    ComputationTimer timer{"LCC"};
    OK(gx_lcc(G, lcc.data()));
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    std::vector<GrB_Index> mapping = ReadMapping(parameters);

    DeviceGraph D = LoadGraph(parameters, 0);
    gx_graph *G = D.G;
    PinnedVector<double> result(D.nrows);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    LA_LCC(G, result);
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;

    SerializeLCCResult(result, mapping, parameters);
    OK(gx_graph_free(G));
    return 0;
}
