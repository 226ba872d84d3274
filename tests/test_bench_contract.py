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


def test_nccl_log_routing(tmp_path):
    """bench.py keeps stdout to its one JSON line: NCCL's INFO log goes into a per-rank file whose lines are copied to
    stderr at the end.  The image's exported NCCL_DEBUG=VERSION counts as unset; an explicit choice of the launcher
    (level or log file) is left alone."""
    code = (
        "import os, sys, importlib.util\n"
        f"spec = importlib.util.spec_from_file_location('bench', {os.path.join(ROOT, 'bench.py')!r})\n"
        "b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)\n"
        "print(os.environ.get('NCCL_DEBUG'), os.environ.get('NCCL_DEBUG_FILE') == b.NCCL_LOG, b.NCCL_LOG is None)\n"
        "if b.NCCL_LOG:\n"
        "    open(b.NCCL_LOG, 'w').write('h:1:1 [0] NCCL INFO Bootstrap\\nh:1:1 [0] NCCL INFO comm 0x1 rank 0 nranks 2\\n')\n"
        "    b.echo_nccl_log()\n"
        "    print(os.path.exists(b.NCCL_LOG))\n")

    def run(env_extra):
        env = {k: v for k, v in os.environ.items() if not k.startswith("NCCL_DEBUG")}
        env.update(env_extra)
        env["TMPDIR"] = str(tmp_path)
        return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=env, timeout=120)

    for preset in ({}, {"NCCL_DEBUG": "VERSION"}):
        r = run(preset)
        assert r.returncode == 0, r.stderr[-1500:]
        assert r.stdout.split() == ["INFO", "True", "False", "False"], r.stdout
        err = r.stderr.splitlines()
        assert err[0].endswith("nranks 2") and "Bootstrap" in err[1]   # communicator line first, nothing on stdout
    r = run({"NCCL_DEBUG": "WARN", "NCCL_DEBUG_FILE": "/dev/null"})
    assert r.returncode == 0 and r.stdout.split() == ["WARN", "False", "True"], r.stdout + r.stderr[-500:]
