#!/usr/bin/env python
"""Per-algorithm device timings on one GPU (development tool; bench.py is the contract).

    python tools/bench_algos.py --algos wcc,cdlp --scale 24 --undirected [--check]

Prints one JSON line per algorithm: kernel ms (CUDA events inside the library), EVPS,
algorithmic bytes and fraction of the measured HBM peak, per-kernel time split."""
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
    ap.add_argument("--algos", default="bfs,pr,wcc,cdlp,lcc,sssp")
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--undirected", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare with the CPU oracle (slow at large scale)")
    ap.add_argument("--hash", action="store_true", help="SHA-256 of the result (to compare a multi-GPU run against this checked one)")
    ap.add_argument("--cache-at", action="store_true", help="directed graphs: build the in-edge adjacency first (BFS may pull)")
    args = ap.parse_args()
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        peak = 6650.0
    capi.init(0)
    algos = args.algos.split(",")
    weighted = "sssp" in algos
    t0 = time.perf_counter()
    g = capi.Graph.rmat(args.scale, not args.undirected, weighted=weighted, want_mapping=False)
    if args.cache_at:
        g.cache(capi.GX_CACHE_AT)
    src = g.max_degree_vertex()
    print(f"# RMAT-{args.scale} {'undirected' if args.undirected else 'directed'}: n={g.n} nnz={g.nnz} "
          f"|E|={g.num_edges} built in {time.perf_counter() - t0:.2f}s", file=sys.stderr)
    ev = g.n + g.num_edges
    run = {"bfs": lambda o: g.bfs(src, out=o), "pr": lambda o: g.pagerank(0.85, 10, out=o), "wcc": lambda o: g.wcc(out=o),
           "cdlp": lambda o: g.cdlp(10, out=o), "lcc": lambda o: g.lcc(out=o), "sssp": lambda o: g.sssp(src, out=o)}
    host = None
    for alg in algos:
        run[alg](False)            # warm-up: builds the cached structures
        build_ms = capi.last_timing()["build_ms"]
        best = None
        for _ in range(args.reps):
            run[alg](False)
            t = capi.last_timing()
            if best is None or t["kernel_ms"] < best["kernel_ms"]:
                best = t
        capi.profile(True)
        run[alg](False)
        capi.profile(False)
        prof = capi.profile_report()
        line = {"alg": alg, "scale": args.scale, "directed": not args.undirected, "n": g.n, "nnz": g.nnz,
                "kernel_ms": round(best["kernel_ms"], 4), "first_call_build_ms": round(build_ms, 3),
                "evps": ev / (best["kernel_ms"] * 1e-3), "iterations": best["iterations"],
                "algorithmic_bytes": best["algorithmic_bytes"], "edges_inspected": best["edges_inspected"],
                "hbm_frac": best["algorithmic_bytes"] / (best["kernel_ms"] * 1e-3) / 1e9 / peak,
                "launches": best["kernel_launches"],
                "kernels": {k: [v[0], round(v[1], 4)] for k, v in list(prof.items())[:8]}}
        if args.hash:
            import hashlib
            line["sha256"] = hashlib.sha256(run[alg](None).tobytes()).hexdigest()
        if args.check:
            import oracle
            if host is None:
                host = g.download()
            rp, ci, w = host
            n, directed = g.n, not args.undirected
            out = run[alg](None)
            t1 = time.perf_counter()
            ref = {"bfs": lambda: oracle.bfs(n, rp, ci, src), "pr": lambda: oracle.pagerank(n, rp, ci, 0.85, 10),
                   "wcc": lambda: oracle.wcc(n, rp, ci, directed), "cdlp": lambda: oracle.cdlp(n, rp, ci, directed, 10),
                   "lcc": lambda: None, "sssp": lambda: oracle.sssp(n, rp, ci, w, src)}[alg]()
            line["oracle_s"] = round(time.perf_counter() - t1, 2)
            if ref is not None:
                if alg == "pr":
                    line["max_rel_err"] = float(np.max(np.abs(out - ref) / ref))
                    line["match"] = line["max_rel_err"] <= 1e-6
                else:
                    line["match"] = bool(np.array_equal(out, ref))
        print(json.dumps(line), flush=True)
    g.free()


if __name__ == "__main__":
    main()
