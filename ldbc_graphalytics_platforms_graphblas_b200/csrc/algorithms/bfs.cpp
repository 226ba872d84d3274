// bfs -- per-algorithm binary `bin/exe/bfs` (execute-job.sh:70-79).  Same flags, log lines,
// exit codes and output format as the reference wrapper (src/algorithms/bfs.cpp:86-113);
// the LAGraph call is replaced by gx_bfs on the B200.
#include <algorithm>
#include <iostream>

#include "cli_common.h"

void SerializeBFSResult(const PinnedVector<int64_t> &level, const std::vector<GrB_Index> &mapping,
                        const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    // unreachable vertices carry GX_UNREACHED_LEVEL == 9223372036854775807 (bfs.cpp:59-63)
    file.lines_int(mapping.data(), level.data(), mapping.size());
}

void LA_BFS(gx_graph *G, GrB_Index sourceVertex, PinnedVector<int64_t> &level)
{
    ComputationTimer timer{"BFS"};
    OK(gx_bfs(G, sourceVertex, level.data()));
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

    // as in the reference (bfs.cpp:79-80: neither AT nor the out-degree is cached) nothing derived is built
    // before the timed window; without a cached A' gx_bfs runs push-only on directed graphs, LAGraph's own rule
    DeviceGraph D = LoadGraph(parameters, 0);
    gx_graph *G = D.G;
    PinnedVector<int64_t> result(D.nrows);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    LA_BFS(G, sourceVertex, result);
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;

    SerializeBFSResult(result, mapping, parameters);
    OK(gx_graph_free(G));
    return 0;
}
