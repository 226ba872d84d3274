#!/usr/bin/env python
"""Run the chosen algorithms ONCE on a fresh RMAT graph (every cached structure is built inside the call) and print the
library's own timing split -- the short command line the ncu launch lists of profiles/ are taken with.

    python tools/one_shot.py --algos lcc --scale 22 --undirected [--twice]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--algos", default="lcc")
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--undirected", action="store_true")
    ap.add_argument("--twice", action="store_true", help="second (warm) call of every algorithm as well")
    args = ap.parse_args()
    capi.init(0)
    algos = args.algos.split(",")
    g = capi.Graph.rmat(args.scale, not args.undirected, weighted="sssp" in algos, want_mapping=False)
    src = g.max_degree_vertex()
    run = {"bfs": lambda: g.bfs(src, out=False), "pr": lambda: g.pagerank(0.85, 10, out=False), "wcc": lambda: g.wcc(out=False),
           "cdlp": lambda: g.cdlp(10, out=False), "lcc": lambda: g.lcc(out=False), "sssp": lambda: g.sssp(src, out=False)}
    for alg in algos:
        for k in range(2 if args.twice else 1):
            run[alg]()
            t = capi.last_timing()
            print(json.dumps({"alg": alg, "call": k + 1, "n": g.n, "nnz": g.nnz, "build_ms": round(t["build_ms"], 3),
                              "kernel_ms": round(t["kernel_ms"], 3), "launches": t["kernel_launches"]}), flush=True)
    g.free()


if __name__ == "__main__":
    main()
