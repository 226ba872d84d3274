// loader_api.cpp -- gx_graph_load: file -> device graph in one call (the loader row of the
// hot-path scope: .mtx/.vtx or .grb/.vtb -> device CSR).
#include <cstdio>
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
        std::vector<GrB_Index> map = p.binary ? ReadVtbFile(p.input_dir + "/graph.vtb") : ReadVtxFile(p.input_dir + "/graph.vtx");
        const char *le = std::getenv("GX_LOADER"); // "host": the host-threaded text parser instead of the device tokenizer
        uint64_t n = 0;
        if (!p.binary && !(le && std::strcmp(le, "host") == 0)) {
            const int rc = gx_graph_load_mtx(g, (p.input_dir + "/graph.mtx").c_str(), directed, 0);
            if (rc != GX_OK) return rc;
            gx_graph_info(*g, &n, nullptr, nullptr, nullptr);
            if (map.size() != n) { gx_graph_free(*g); *g = nullptr; return GX_ERR_IO; }
        } else {
            HostMatrix A = p.binary ? ReadGrbFile(p.input_dir + "/graph.grb") : ReadMtxFile(p.input_dir + "/graph.mtx");
            if (map.size() != A.nrows) return GX_ERR_IO;
            const int rc = gx_graph_create_csr32(g, A.nrows, A.nvals, A.Ap.data(), A.Aj.data(), A.iso ? nullptr : A.Ax.data(), directed);
            if (rc != GX_OK) return rc;
            n = A.nrows;
        }
        if (mapping) {
            *mapping = (uint64_t *)malloc((map.size() ? map.size() : 1) * sizeof(uint64_t));
            if (!*mapping) return GX_ERR_OOM;
            memcpy(*mapping, map.data(), map.size() * sizeof(uint64_t));
        }
        if (n_out) *n_out = n;
        return GX_OK;
    } catch (const std::exception &) {
        return GX_ERR_IO;
    }
}

// gx_result_write: the six Serialize*Result functions of the reference wrappers (bfs.cpp:11-68, pr.cpp:17-45,
// wcc.cpp:11-37, cdlp.cpp:21-52, lcc.cpp:17-59, sssp.cpp:11-51) as one library call: "<id> <value>\n" per vertex,
// formatted on all host threads.  Host-only: needs no device.
extern "C" int gx_result_write(const char *path, int kind, const uint64_t *ids, const void *values, uint64_t n, const uint64_t *value_map)
{
    if (!path || (n && (!ids || !values))) return GX_ERR_INVALID;
    try {
        ResultWriter file(path);
        if (!file.ok()) return GX_ERR_IO;
        switch (kind) {
        case GX_RESULT_INT64: file.lines_int(ids, (const int64_t *)values, n); break;
        case GX_RESULT_UINT64: file.lines_uint(ids, (const uint64_t *)values, n, value_map); break;
        case GX_RESULT_FP64: file.lines_sci(ids, (const double *)values, n); break;
        default: return GX_ERR_INVALID;
        }
        return GX_OK;
    } catch (const std::exception &) {
        return GX_ERR_IO;
    }
}

// gx_relabel: the relabelling stage of load-graph.sh (bin/py/relabel.py:8-79) without DuckDB.  Host-only.
extern "C" int gx_relabel(const char *vertex_path, const char *edge_path, const char *out_dir, int weighted, int directed,
                          uint64_t *n_out, uint64_t *nnz_out)
{
    if (!vertex_path || !edge_path || !out_dir) return GX_ERR_INVALID;
    try {
        RelabelGraph(vertex_path, edge_path, out_dir, weighted != 0, directed != 0, n_out, nnz_out);
        return GX_OK;
    } catch (const std::exception &e) {
        fprintf(stderr, "gx_relabel: %s\n", e.what());
        return GX_ERR_IO;
    }
}
