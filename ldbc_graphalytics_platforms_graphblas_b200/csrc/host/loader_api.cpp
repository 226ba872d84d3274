// loader_api.cpp -- gx_graph_load: file -> device graph in one call (the loader row of the
// hot-path scope: .mtx/.vtx or .grb/.vtb -> device CSR).
#include <cstdlib>
#include <cstring>

#include "graphio.h"

extern "C" int gx_graph_load(gx_graph **g, const char *dir, int binary, int directed, uint64_t **mapping, uint64_t *n_out)
{
    if (!g || !dir) return GX_ERR_INVALID;
    try {
        BenchmarkParameters p;
        p.binary = binary != 0;
        p.input_dir = dir;
        p.directed = directed != 0;
        HostMatrix A = p.binary ? ReadGrbFile(p.input_dir + "/graph.grb") : ReadMtxFile(p.input_dir + "/graph.mtx");
        std::vector<GrB_Index> map = p.binary ? ReadVtbFile(p.input_dir + "/graph.vtb") : ReadVtxFile(p.input_dir + "/graph.vtx");
        if (map.size() != A.nrows) return GX_ERR_IO;
        int rc = gx_graph_create_csr32(g, A.nrows, A.nvals, A.Ap.data(), A.Aj.data(), A.iso ? nullptr : A.Ax.data(), directed);
        if (rc != GX_OK) return rc;
        if (mapping) {
            *mapping = (uint64_t *)malloc((map.size() ? map.size() : 1) * sizeof(uint64_t));
            if (!*mapping) return GX_ERR_OOM;
            memcpy(*mapping, map.data(), map.size() * sizeof(uint64_t));
        }
        if (n_out) *n_out = A.nrows;
        return GX_OK;
    } catch (const std::exception &) {
        return GX_ERR_IO;
    }
}
