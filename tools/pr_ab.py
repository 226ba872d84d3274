#!/usr/bin/env python
"""Development tool: PageRank kernel variants timed in one process on RMAT-22 directed.
    python tools/pr_ab.py --vars 0,2,4 --reps 5
GX_PR_VAR 0 is the product kernel; 2 and 4 are timing diagnostics that break the result (epilogue operands
not loaded / no gathers): what each dependent memory phase of a tile costs.  Two libraries can be compared
by running the tool twice in one gpurun call with GX_LIB pointing at the other build."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--vars", default="0")
ap.add_argument("--scale", type=int, default=22)
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
capi.init(0)
g = capi.Graph.rmat(args.scale, True, weighted=False, want_mapping=False)
g.cache(capi.GX_CACHE_AT)
for rnd in range(2):                         # two interleaved rounds: drift shows up as disagreement between them
    for v in args.vars.split(","):
        os.environ["GX_PR_VAR"] = v
        g.pagerank(0.85, 10, out=False)
        ms = []
        for _ in range(args.reps):
            g.pagerank(0.85, 10, out=False)
            ms.append(capi.last_timing()["kernel_ms"])
        capi.profile(True)
        g.pagerank(0.85, 10, out=False)
        capi.profile(False)
        prof = {k: round(x[1] / x[0] * 1e3, 1) for k, x in capi.profile_report().items() if k.startswith("k_pr_tile")}
        print(json.dumps({"var": v, "round": rnd, "pr_ms_min": round(min(ms), 4),
                          "pr_ms_med": round(sorted(ms)[len(ms) // 2], 4), "us": prof}), flush=True)
g.free()
