#!/usr/bin/env python3
"""DuckDB-free drop-in for the reference's bin/py/relabel.py (same flags, same output files).

    relabel.py --graph-name G --input-vertex-path X.v --input-edge-path X.e --output-path DIR \
               --weighted true|false --directed true|false [--use-disk]

Writes DIR/graph.vtx (original id of dense vertex k on line k, .v row order) and DIR/graph.mtx
(`%%MatrixMarket matrix coordinate integer|real general|symmetric`, `%%GraphBLAS GrB_BOOL|GrB_FP64`,
`n n nnz`, then 1-based `src dst val`, entries in .e order) -- the format of relabel.py:52-79 that
bin/sh/load-graph.sh:50-60 expects before it runs bin/exe/converter."""
import argparse
import os
import sys

import numpy as np


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

    print("Loading...")
    ids = np.loadtxt(a.input_vertex_path, dtype=np.uint64, ndmin=1)
    cols = 3 if a.weighted else 2
    if os.path.getsize(a.input_edge_path):
        # edge weights are kept as text so that they round-trip byte for byte
        raw = np.loadtxt(a.input_edge_path, dtype=str, ndmin=2, usecols=range(cols))
    else:
        raw = np.zeros((0, cols), dtype=str)
    print("Relabelling...")
    order = np.argsort(ids, kind="stable")
    sorted_ids = ids[order]

    def dense(col):
        v = col.astype(np.uint64)
        pos = np.searchsorted(sorted_ids, v)
        bad = (pos >= ids.size) | (sorted_ids[np.minimum(pos, ids.size - 1)] != v)
        if bad.any():
            sys.exit(f"edge endpoint {v[bad][0]} is not in the vertex file")
        return order[pos] + 1  # Matrix Market indexes from 1

    src = dense(raw[:, 0]) if raw.size else np.zeros(0, np.int64)
    dst = dense(raw[:, 1]) if raw.size else np.zeros(0, np.int64)
    os.makedirs(a.output_path, exist_ok=True)
    print("Serializing textual mapping file (vtx)")
    np.savetxt(os.path.join(a.output_path, "graph.vtx"), ids, fmt="%d")
    print("Serializing textual matrix file (mtx)")
    with open(os.path.join(a.output_path, "graph.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate {} {}\n".format("real" if a.weighted else "integer",
                                                                    "general" if a.directed else "symmetric"))
        f.write("%%GraphBLAS {}\n".format("GrB_FP64" if a.weighted else "GrB_BOOL"))
        f.write(f"{ids.size} {ids.size} {src.size}\n")
        val = raw[:, 2] if a.weighted else np.full(src.size, "1")
        f.write("".join(f"{s} {d} {v}\n" for s, d, v in zip(src, dst, val)))


if __name__ == "__main__":
    main()
