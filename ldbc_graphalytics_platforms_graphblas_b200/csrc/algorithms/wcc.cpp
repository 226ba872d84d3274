// wcc -- per-algorithm binary `bin/exe/wcc` (execute-job.sh:81-90), the drop-in for
// src/algorithms/wcc.cpp:68-85 with A v A' + LAGr_ConnectedComponents replaced by gx_wcc.
#include <iostream>

#include "cli_common.h"

void SerializeWCCResult(const PinnedVector<uint64_t> &comp, const std::vector<GrB_Index> &mapping,
                        const BenchmarkParameters &parameters)
{
    ResultWriter file = OpenOutput(parameters);
    // like the reference, the component id is the dense representative, not mapped back
    // (wcc.cpp:31-34); the validator only needs equivalent partitions
    file.lines_uint(mapping.data(), comp.data(), mapping.size());
}

void WeaklyConnectedComponents(gx_graph *G, PinnedVector<uint64_t> &comp)
{
    ComputationTimer total_timer{"WeaklyConnectedComponents"};
    OK(gx_wcc(G, comp.data()));
}

int main(int argc, char **argv)
{
    BenchmarkParameters parameters = ParseBenchmarkParameters(argc, argv);
    InitDevice();
    std::vector<GrB_Index> mapping = ReadMapping(parameters);

    // the reference symmetrises (A v A', wcc.cpp:53-55) inside its timed window; gx_wcc builds what it needs
    // of the in-edges on first use, between the two Processing lines
    DeviceGraph D = LoadGraph(parameters, 0);
    gx_graph *G = D.G;
    PinnedVector<uint64_t> result(D.nrows);
    std::cout << "Processing starts at: " << GetCurrentMilliseconds() << std::endl;
    WeaklyConnectedComponents(G, result);
    std::cout << "Processing ends at: " << GetCurrentMilliseconds() << std::endl;

    SerializeWCCResult(result, mapping, parameters);
    OK(gx_graph_free(G));
    return 0;
}
