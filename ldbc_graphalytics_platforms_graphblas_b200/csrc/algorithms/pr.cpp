// pr -- per-algorithm binary `bin/exe/pr` (execute-job.sh:92-103), the drop-in for
// src/algorithms/pr.cpp:68-85 with LAGr_PageRankGX replaced by gx_pagerank.
#include <iostream>

#include "cli_common.h"

void SerializePageRankResult(const std::vector<double> &rank, const std::vector<GrB_Index> &mapping,
                             const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    for (GrB_Index v = 0; v < mapping.size(); v++) file.line_sci(mapping[v], rank[v]);
}

std::vector<double> LA_PR(gx_graph *G, GrB_Index n, double damping_factor, int iteration_num)
{
    ComputationTimer timer{"PageRank"};
    std::vector<double> rank(n);
    OK(gx_pagerank(G, damping_factor, iteration_num, rank.data()));
    return rank;
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    HostMatrix A = ReadMatrixMarket(parameters);
    std::vector<GrB_Index> mapping = ReadMapping(parameters);

    // the reference transposes inside its timed window (LAGraph_Cached_AT + Cached_OutDegree, pr.cpp:58-61);
    // so does gx_pagerank here: the in-edge adjacency and the tile plan are built on first use, between the
    // two Processing lines
    gx_graph *G = UploadGraph(A, parameters.directed, 0);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    std::vector<double> result = LA_PR(G, A.nrows, parameters.damping_factor, parameters.max_iteration);
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;

    SerializePageRankResult(result, mapping, parameters);
    OK(gx_graph_free(G));
    return 0;
}
