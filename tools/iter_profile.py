#!/usr/bin/env python
"""Development tool: per-iteration kernel split of CDLP (runs with 1..K iterations, differences of the
per-kernel totals) and an SSSP bucket-width sweep, on one RMAT graph.

    python tools/iter_profile.py --scale 24 --undirected --cdlp 10 --sssp-deltas 0.02,0.05,0.1"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--undirected", action="store_true")
    ap.add_argument("--cdlp", type=int, default=0)
    ap.add_argument("--sssp-deltas", default="")
    args = ap.parse_args()
    capi.init(0)
    deltas = [float(x) for x in args.sssp_deltas.split(",") if x]
    g = capi.Graph.rmat(args.scale, not args.undirected, weighted=bool(deltas), want_mapping=False)
    if args.cdlp:
        g.cdlp(1, out=False)
        prev, prev_ms, prev_insp = {}, 0.0, 0
        for k in range(1, args.cdlp + 1):
            g.cdlp(k, out=False)
            t = capi.last_timing()
            capi.profile(True)
            g.cdlp(k, out=False)
            capi.profile(False)
            prof = {name: v[1] for name, v in capi.profile_report().items()}
            diff = {name: round(ms - prev.get(name, 0.0), 3) for name, ms in prof.items() if ms - prev.get(name, 0.0) > 0.02}
            print(json.dumps({"alg": "cdlp", "iteration": k, "ms": round(t["kernel_ms"] - prev_ms, 3),
                              "entries": t["edges_inspected"] - prev_insp, "kernels": diff}), flush=True)
            prev, prev_ms, prev_insp = prof, t["kernel_ms"], t["edges_inspected"]
    if deltas:
        src = g.max_degree_vertex()
        for d in [None] + deltas:
            if d is None:
                os.environ.pop("GX_SSSP_DELTA", None)
            else:
                os.environ["GX_SSSP_DELTA"] = repr(d)
            g.sssp(src, out=False)
            best = None
            for _ in range(2):
                g.sssp(src, out=False)
                t = capi.last_timing()
                if best is None or t["kernel_ms"] < best["kernel_ms"]:
                    best = t
            print(json.dumps({"alg": "sssp", "delta": d, "ms": round(best["kernel_ms"], 3), "rounds": best["iterations"],
                              "relaxed_per_entry": round(best["edges_inspected"] / g.nnz, 3)}), flush=True)
    g.free()


if __name__ == "__main__":
    main()
