#!/usr/bin/env python
"""Development tool: graph.mtx -> device graph, device tokenizer against the host-threaded parser.
    python tools/load_bench.py --scale 22 [--undirected] [--weighted] [--dir /tmp/gxload]
Builds the RMAT graph on the device, writes it as a Graphalytics input directory with the library's own writers
(.e / .v -> gx_relabel -> graph.mtx + graph.vtx), loads it back through gx_graph_load with both loaders and checks
that the three graphs are identical.  One JSON line."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--undirected", action="store_true")
    ap.add_argument("--weighted", action="store_true")
    ap.add_argument("--dir", default="/tmp/gxload")
    args = ap.parse_args()
    capi.init(0)
    os.makedirs(args.dir, exist_ok=True)
    g = capi.Graph.rmat(args.scale, not args.undirected, weighted=args.weighted, want_mapping=True)
    rp, ci, w = g.download()
    w_all = w
    ids = g.mapping
    n = g.n
    t0 = time.perf_counter()
    rows = np.repeat(np.arange(n, dtype=np.uint64), np.diff(rp.astype(np.int64)))
    cols = ci.astype(np.uint64)
    if args.undirected:
        keep = rows < cols
        rows, cols = rows[keep], cols[keep]
        w = None if w is None else w[keep]
    vpath, epath = os.path.join(args.dir, "graph.v"), os.path.join(args.dir, "graph.e")
    capi.write_result(vpath, ids, ids)                 # "<id> <id>": relabel reads the first column of a .v line
    if args.weighted:
        # "<src> <dst> <weight>": the .e writer has two columns, so the weights go through %.17g text
        with open(epath, "w") as f:
            src, dst = ids[rows.astype(np.int64)], ids[cols.astype(np.int64)]
            step = 1 << 20
            for a in range(0, rows.size, step):
                f.write("".join(f"{s} {d} {x!r}\n" for s, d, x in zip(src[a:a + step].tolist(), dst[a:a + step].tolist(), w[a:a + step].tolist())))
    else:
        capi.write_result(epath, ids[rows.astype(np.int64)], ids[cols.astype(np.int64)])
    t1 = time.perf_counter()
    capi.relabel(vpath, epath, args.dir, weighted=args.weighted, directed=not args.undirected)
    t2 = time.perf_counter()
    size = os.path.getsize(os.path.join(args.dir, "graph.mtx"))
    out = {"graph": f"RMAT-{args.scale} {'undirected' if args.undirected else 'directed'}{' weighted' if args.weighted else ''}",
           "n": n, "nnz": g.nnz, "mtx_bytes": size, "write_e_v_s": round(t1 - t0, 2), "relabel_s": round(t2 - t1, 2)}
    res = {}
    for loader in ("device", "host", "device"):          # the second device run reads a warm page cache, like the host run
        if loader == "host":
            os.environ["GX_LOADER"] = "host"
        else:
            os.environ.pop("GX_LOADER", None)
        t = time.perf_counter()
        h = capi.Graph.load(args.dir, False, not args.undirected)
        dt = time.perf_counter() - t
        res[loader] = h.download()
        out[f"{loader}_load_s"] = round(dt, 3)
        out[f"{loader}_gb_per_s"] = round(size / dt / 1e9, 3)
        if loader == "device":
            tm = capi.last_timing()
            out["device_text_upload_ms"] = round(tm["h2d_ms"], 1)
            out["device_tokenise_and_build_ms"] = round(tm["build_ms"], 1)
            out["values_settled_by_host_strtod"] = tm["iterations"]
        assert np.array_equal(h.mapping, ids)
        h.free()
    same = all((x is None and y is None) or np.array_equal(x, y) for x, y in zip(res["device"], res["host"]))
    orig = all((x is None and y is None) or np.array_equal(x, y) for x, y in zip(res["device"], (rp, ci, w_all)))
    out["device_equals_host"] = bool(same)
    out["equals_generated_graph"] = bool(orig)
    out["speedup"] = round(out["host_load_s"] / out["device_load_s"], 2)
    print(json.dumps(out))
    g.free()
    sys.exit(0 if same and orig else 1)


if __name__ == "__main__":
    main()
