#!/usr/bin/env python
"""Development tool: a Graphalytics input directory (graph.mtx + graph.vtx) of a Graph500 RMAT graph, written with the
library's own multi-threaded text writers (gx_result_write for the .e file, gx_relabel for the .mtx / .vtx pair).
    python tools/make_mtx.py --scale 20 --out /tmp/rmat20 [--undirected] [--weighted]
Host-only (no device needed)."""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi, rmat  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=20)
    ap.add_argument("--out", required=True)
    ap.add_argument("--undirected", action="store_true")
    ap.add_argument("--weighted", action="store_true")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    t0 = time.perf_counter()
    s, d = rmat.rmat_edges(args.scale)
    keep = s != d
    s, d = s[keep], d[keep]
    if args.undirected:
        lo, hi = np.minimum(s, d), np.maximum(s, d)
        s, d = lo, hi
    key = np.unique((s << np.uint64(32)) | d)
    s, d = key >> np.uint64(32), key & np.uint64(0xFFFFFFFF)
    ids = np.unique(np.concatenate([s, d]))
    vpath, epath = os.path.join(args.out, "graph.v"), os.path.join(args.out, "graph.e")
    with open(vpath, "w") as f:
        f.write("\n".join(map(str, ids.tolist())) + "\n")
    if args.weighted:
        w = rmat.edge_weights(s, d, rmat.default_seed(args.scale))
        with open(epath, "w") as f:
            for a, b, x in zip(s.tolist(), d.tolist(), w.tolist()):
                f.write(f"{a} {b} {x!r}\n")
    else:
        capi.write_result(epath, s, d)
    t1 = time.perf_counter()
    n, nnz = capi.relabel(vpath, epath, args.out, weighted=args.weighted, directed=not args.undirected)
    print(f"RMAT-{args.scale}: n={n} nnz={nnz}; .v/.e in {t1 - t0:.1f}s, relabel in {time.perf_counter() - t1:.1f}s; "
          f"graph.mtx is {os.path.getsize(os.path.join(args.out, 'graph.mtx')) / 1e6:.0f} MB")


if __name__ == "__main__":
    main()
