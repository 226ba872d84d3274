// sssp -- per-algorithm binary `bin/exe/sssp` (execute-job.sh:128-138), the drop-in for
// src/algorithms/sssp.cpp:83-111 with LAGr_SingleSourceShortestPath replaced by gx_sssp.
#include <algorithm>
#include <cmath>
#include <iostream>

#include "cli_common.h"

void SerializeSSSPResult(const std::vector<double> &dist, const std::vector<GrB_Index> &mapping,
                         const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    for (GrB_Index v = 0; v < mapping.size(); v++) {
        if (std::isinf(dist[v])) file.line_text(mapping[v], "infinity"); // sssp.cpp:41-46
        else file.line_sci(mapping[v], dist[v]);
    }
}

std::vector<double> LA_SSSP(gx_graph *G, GrB_Index sourceVertex, GrB_Index n)
{
    ComputationTimer timer{"SSSP"};
    std::vector<double> dist(n);
    OK(gx_sssp(G, sourceVertex, dist.data()));
    return dist;
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    HostMatrix A = ReadMatrixMarket(parameters);
    std::vector<GrB_Index> mapping = ReadMapping(parameters);

    auto it = std::find(mapping.begin(), mapping.end(), (GrB_Index)parameters.source_vertex);
    if (it == mapping.end()) {
        std::cout << "Source vertex not found in mapping" << std::endl;
        return -1;
    }
    const GrB_Index sourceVertex = (GrB_Index)std::distance(mapping.begin(), it);
    if (A.iso) throw std::runtime_error("SSSP needs a weighted graph (graph.mtx of type real)");

    gx_graph *G = UploadGraph(A, parameters.directed, 0);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    std::vector<double> result = LA_SSSP(G, sourceVertex, A.nrows);
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;

    SerializeSSSPResult(result, mapping, parameters);
    OK(gx_graph_free(G));
    return 0;
}
