#!/usr/bin/env python
"""Contract processing time of the drop-in binaries: what the Graphalytics harness records.

    python tools/cold_tproc.py --scale 22 [--algos bfs,pr,wcc,cdlp,lcc,sssp] [--undirected]

Writes the RMAT graph as graph.grb/graph.vtb (and the weighted variant for SSSP), runs bin/exe/<alg> exactly as
bin/sh/execute-job.sh:70-139 does (one fresh process per job, `--binary true`), parses `Processing ends at` minus
`Processing starts at` (GraphblasCollector.java:54-95) and prints it next to the device time of the same algorithm on
the same graph with every cached structure warm (kernel_ms of a second call through the C ABI).  One JSON line per
algorithm: cold EVPS is what the harness would report, warm EVPS what a resident graph sustains."""
import argparse
import json
import os
import re
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi, graphio  # noqa: E402

EXE = os.path.join(ROOT, "bin", "exe")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--algos", default="bfs,pr,wcc,cdlp,lcc,sssp")
    ap.add_argument("--undirected", action="store_true")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--workdir", default=None)
    args = ap.parse_args()
    algos = args.algos.split(",")
    directed = not args.undirected
    capi.init(0)
    g = capi.Graph.rmat(args.scale, directed, weighted="sssp" in algos, want_mapping=True)
    n, nnz = g.n, g.nnz
    ev = n + g.num_edges
    src = g.max_degree_vertex()
    run = {"bfs": lambda: g.bfs(src, out=False), "pr": lambda: g.pagerank(0.85, 10, out=False), "wcc": lambda: g.wcc(out=False),
           "cdlp": lambda: g.cdlp(10, out=False), "lcc": lambda: g.lcc(out=False), "sssp": lambda: g.sssp(src, out=False)}
    if directed:
        g.cache(capi.GX_CACHE_AT)
    warm = {}
    for alg in algos:
        run[alg]()
        best = None
        for _ in range(args.reps):
            run[alg]()
            t = capi.last_timing()
            best = t["kernel_ms"] if best is None else min(best, t["kernel_ms"])
        warm[alg] = best
    rp, ci, w = g.download()
    mapping = g.mapping
    g.free()
    work = args.workdir or tempfile.mkdtemp(prefix="gx_cold_")
    d_plain, d_w = os.path.join(work, "g"), os.path.join(work, "g.e_weight")
    graphio.write_graph_dir(d_plain, graphio.HostGraph(n, rp, ci, None, directed, mapping), binary=True)
    if "sssp" in algos:
        graphio.write_graph_dir(d_w, graphio.HostGraph(n, rp, ci, w, directed, mapping), binary=True)
    src_orig = int(mapping[src])
    del rp, ci, w
    for alg in algos:
        cmd = [os.path.join(EXE, alg), "--binary", "true", "--jobid", "cold", "--input-dir", d_w if alg == "sssp" else d_plain,
               "--output-file", os.path.join(work, f"out-{alg}"), "--directed", "true" if directed else "false"]
        if alg in ("bfs", "sssp"):
            cmd += ["--source-vertex", str(src_orig)]
        if alg == "pr":
            cmd += ["--damping-factor", "0.85", "--max-iteration", "10"]
        if alg == "cdlp":
            cmd += ["--max-iteration", "10"]
        cmd += ["--log-path", work, "--threadnum", "16"]
        tp = []
        for _ in range(args.reps):
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=1800)
            if r.returncode != 0:
                print(json.dumps({"alg": alg, "error": (r.stdout + r.stderr)[-400:]}), flush=True)
                break
            a = int(re.findall(r"Processing starts at: (\d+)", r.stdout)[-1])
            b = int(re.findall(r"Processing ends at: (\d+)", r.stdout)[-1])
            tp.append(b - a)
        if not tp:
            continue
        cold = min(tp)
        print(json.dumps({"alg": alg, "graph": f"RMAT-{args.scale} {'directed' if directed else 'undirected'}", "n": n, "nnz": nnz,
                          "contract_tproc_ms": cold, "all_runs_ms": tp, "warm_device_ms": round(warm[alg], 3),
                          "cold_over_warm": round(cold / max(warm[alg], 1e-9), 2),
                          "contract_evps": ev / (max(cold, 0.5) * 1e-3), "warm_evps": ev / (warm[alg] * 1e-3),
                          "note": "contract = epoch-ms between the two Processing lines of a fresh bin/exe process (1 ms resolution); "
                                  "warm = CUDA-event kernel time of a repeated call on a resident graph"}), flush=True)


if __name__ == "__main__":
    main()
