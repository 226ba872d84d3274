#!/usr/bin/env python
"""Development tool: A/B of PageRank kernel variants in one process (GX_PR_VAR / GX_PR_HOT), RMAT-22 directed.
    python tools/pr_ab.py --vars 0,1 --reps 5"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vars", default="0,1")
ap.add_argument("--hot", default="")
ap.add_argument("--pipe", default="0")
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
capi.init(0)
g = capi.Graph.rmat(args.scale, True, weighted=False, want_mapping=False)
g.cache(capi.GX_CACHE_AT)
hots = [h for h in args.hot.split(",") if h] or [None]
for rnd in range(2):                         # two interleaved rounds: drift shows up as disagreement between them
    for v, pp in [(v, pp) for v in args.vars.split(",") for pp in args.pipe.split(",")]:
        for h in hots:
            os.environ["GX_PR_VAR"] = v
            os.environ["GX_PR_PIPE"] = pp
            if h is not None:
                os.environ["GX_PR_HOT"] = h
            g.pagerank(0.85, 10, out=False)
            ms = []
            for _ in range(args.reps):
                g.pagerank(0.85, 10, out=False)
                ms.append(capi.last_timing()["kernel_ms"])
            capi.profile(True); g.pagerank(0.85, 10, out=False); capi.profile(False)
            prof = {k: round(x[1] / x[0] * 1e3, 1) for k, x in capi.profile_report().items() if k.startswith("k_pr_tile")}
            print(json.dumps({"var": v, "pipe": pp, "hot": h, "round": rnd, "pr_ms_min": round(min(ms), 4), "pr_ms_med": round(sorted(ms)[len(ms) // 2], 4), "us": prof}), flush=True)
g.free()
