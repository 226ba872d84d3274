#!/usr/bin/env python3
"""DuckDB-free drop-in for the reference's bin/py/relabel.py (same flags, same output files).

    relabel.py --graph-name G --input-vertex-path X.v --input-edge-path X.e --output-path DIR \
               --weighted true|false --directed true|false [--use-disk]

Writes DIR/graph.vtx (original id of dense vertex k on line k, .v row order) and DIR/graph.mtx
(`%%MatrixMarket matrix coordinate integer|real general|symmetric`, `%%GraphBLAS GrB_BOOL|GrB_FP64`,
`n n nnz`, then 1-based `src dst val`, entries in .e order) -- the format of relabel.py:52-79 that
bin/sh/load-graph.sh:50-60 expects before it runs bin/exe/converter.

The work is done by gx_relabel in libgxb200.so (csrc/host/graphio.cpp RelabelGraph): both files are parsed
in byte ranges, the endpoints looked up and the text formatted on all host threads -- an RMAT-22 edge
file (65 M lines) takes seconds, where a numpy.loadtxt + Python string join took minutes and would not
survive RMAT-24.  Host-only: no GPU is needed for this stage."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))


def truth(x):
    return str(x).lower() in ("true", "1", "yes")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--graph-name", type=str, required=True)
    ap.add_argument("--input-vertex-path", type=str, required=True)
    ap.add_argument("--input-edge-path", type=str, required=True)
    ap.add_argument("--output-path", type=str, required=True)
    ap.add_argument("--weighted", type=truth, required=True)
    ap.add_argument("--directed", type=truth, required=True)
    ap.add_argument("--use-disk", action="store_true", required=False)  # accepted, no effect
    a = ap.parse_args()

    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    print("Loading...")
    print("Relabelling...")
    print("Serializing textual mapping file (vtx)")
    print("Serializing textual matrix file (mtx)")
    try:
        capi.relabel(a.input_vertex_path, a.input_edge_path, a.output_path, a.weighted, a.directed)
    except capi.GxError as e:
        sys.exit(str(e))


if __name__ == "__main__":
    main()
