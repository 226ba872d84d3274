// converter -- `bin/exe/converter --data-dir D`: graph.vtx/graph.mtx -> graph.vtb/graph.grb,
// the GraphBLAS-free counterpart of src/tools/converter.cpp:16-60 (load-graph.sh:62-67 runs it
// and execute-job.sh:72 always passes --binary true, so the binaries read its output).
#include <iostream>

#include "graphio.h"

int main(int argc, char **argv)
{
    ConverterParameters parameters = ParseConverterParameters(argc, argv);
    std::vector<GrB_Index> mapping = ReadVtxFile(parameters.data_dir + "/graph.vtx");
    HostMatrix A = ReadMtxFile(parameters.data_dir + "/graph.mtx");
    if (mapping.size() != A.nrows) throw std::runtime_error("graph.vtx and graph.mtx disagree on the vertex count");

    std::cout << "Serializing binary mapping file (vtb)" << std::endl;
    WriteVtbFile(parameters.data_dir + "/graph.vtb", mapping);
    std::cout << "Serializing binary matrix file (grb)" << std::endl;
    WriteGrbFile(parameters.data_dir + "/graph.grb", A);
    return 0;
}
