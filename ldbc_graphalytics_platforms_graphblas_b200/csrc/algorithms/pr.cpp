// pr -- per-algorithm binary `bin/exe/pr` (execute-job.sh:92-103), the drop-in for
// src/algorithms/pr.cpp:68-85 with LAGr_PageRankGX replaced by gx_pagerank.
#include <iostream>

#include "cli_common.h"

void SerializePageRankResult(const PinnedVector<double> &rank, const std::vector<GrB_Index> &mapping,
                             const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    file.lines_sci(mapping.data(), rank.data(), mapping.size());
}

void LA_PR(gx_graph *G, double damping_factor, int iteration_num, PinnedVector<double> &rank)
{
    ComputationTimer timer{"PageRank"};
    OK(gx_pagerank(G, damping_factor, iteration_num, rank.data()));
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    std::vector<GrB_Index> mapping = ReadMapping(parameters);

    // the reference transposes inside its timed window (LAGraph_Cached_AT + Cached_OutDegree, pr.cpp:58-61);
    // so does gx_pagerank here: the in-edge adjacency and the tile plan are built on first use, between the
    // two Processing lines
    DeviceGraph D = LoadGraph(parameters, 0);
    gx_graph *G = D.G;
    PinnedVector<double> result(D.nrows);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    LA_PR(G, parameters.damping_factor, parameters.max_iteration, result);
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;

    SerializePageRankResult(result, mapping, parameters);
    OK(gx_graph_free(G));
    return 0;
}
