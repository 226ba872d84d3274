"""bench.py contract, CPU side: the reference arm (`--impl reference`) prints ONE JSON line with the keys the
driver reads, runs only the oracle on host cores, and the host-side graph construction it uses builds exactly
the graph the device generator builds (same vertices, same entries)."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from ldbc_graphalytics_platforms_graphblas_b200 import rmat


def test_reference_arm_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--scale", "14",
                        "--steps", "2", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["unit"] == "edges+vertices/s"
    for k in ("metric", "value", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "dtype", "data", "config"):
        assert k in d, k
    assert d["value"] > 0 and d["config"]["workload"].startswith("BFS + PageRank")
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and "sample" in cb
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_builds_the_benchmark_graph():
    sys.path.insert(0, ROOT)
    import bench
    g = bench.host_rmat(15)
    ref = rmat.rmat_graph(15, directed=True)
    assert g.n == ref.n and np.array_equal(g.rowptr, ref.rowptr) and np.array_equal(g.colidx, ref.colidx)
    assert np.array_equal(g.mapping, ref.mapping)
