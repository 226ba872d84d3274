// sssp -- per-algorithm binary `bin/exe/sssp` (execute-job.sh:128-138), the drop-in for
// src/algorithms/sssp.cpp:83-111 with LAGr_SingleSourceShortestPath replaced by gx_sssp.
#include <algorithm>
#include <cmath>
#include <iostream>

#include "cli_common.h"

void SerializeSSSPResult(const PinnedVector<double> &dist, const std::vector<GrB_Index> &mapping,
                         const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    file.lines_sci(mapping.data(), dist.data(), mapping.size()); // +inf is written as `infinity` (sssp.cpp:41-46)
}

void LA_SSSP(gx_graph *G, GrB_Index sourceVertex, PinnedVector<double> &dist)
{
    ComputationTimer timer{"SSSP"};
    OK(gx_sssp(G, sourceVertex, dist.data()));
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    std::vector<GrB_Index> mapping = ReadMapping(parameters);

    auto it = std::find(mapping.begin(), mapping.end(), (GrB_Index)parameters.source_vertex);
    if (it == mapping.end()) {
        std::cout << "Source vertex not found in mapping" << std::endl;
        return -1;
    }
    const GrB_Index sourceVertex = (GrB_Index)std::distance(mapping.begin(), it);

    DeviceGraph D = LoadGraph(parameters, 0);
    gx_graph *G = D.G;
    if (!D.weighted) throw std::runtime_error("SSSP needs a weighted graph (graph.mtx of type real)");
    PinnedVector<double> result(D.nrows);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    LA_SSSP(G, sourceVertex, result);
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;

    SerializeSSSPResult(result, mapping, parameters);
    OK(gx_graph_free(G));
    return 0;
}
