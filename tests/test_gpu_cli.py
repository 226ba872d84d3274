"""Drop-in boundary: the six per-algorithm binaries invoked exactly as
bin/sh/execute-job.sh:70-139 invokes them, on the reference's fixture graphs,
validated against the bundled golden outputs with the Graphalytics rules."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden_cases
from helpers import golden
from ldbc_graphalytics_platforms_graphblas_b200 import graphio, validator

pytestmark = pytest.mark.gpu
EXE = os.path.join(ROOT, "bin", "exe")


def prepare(tmp_path, name, weighted, binary):
    params = graphio.graph_params(os.path.join(GOLDEN, name + ".properties"))
    d = tmp_path / (name + (".e_weight" if weighted else ""))
    graphio.write_vtx_mtx(str(d), os.path.join(GOLDEN, name + ".v"), os.path.join(GOLDEN, name + ".e"),
                          params["directed"], weighted)
    if binary:
        subprocess.check_call([os.path.join(EXE, "converter"), "--data-dir", str(d)])
    return d, params


@pytest.mark.parametrize("binary", ["true", "false"])
@pytest.mark.parametrize("name,alg", golden_cases())
def test_binary_reproduces_golden(tmp_path, name, alg, binary):
    weighted = alg == "SSSP"
    d, params = prepare(tmp_path, name, weighted, binary == "true")
    out = tmp_path / "output"
    cmd = [os.path.join(EXE, alg.lower()), "--binary", binary, "--jobid", "j1", "--input-dir", str(d),
           "--output-file", str(out), "--directed", "true" if params["directed"] else "false"]
    if alg == "BFS":
        cmd += ["--source-vertex", str(params["bfs_source"])]
    if alg == "SSSP":
        cmd += ["--source-vertex", str(params["sssp_source"])]
    if alg == "PR":
        cmd += ["--damping-factor", str(np.float32(params["pr_damping"])), "--max-iteration", str(params["pr_iters"])]
    if alg == "CDLP":
        cmd += ["--max-iteration", str(params["cdlp_iters"])]
    cmd += ["--log-path", str(tmp_path), "--threadnum", "4"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    # GraphblasCollector.java:68-76: substring match, last token parsed as a long (epoch ms)
    start = re.findall(r"Processing starts at: (\d+)", r.stdout)
    end = re.findall(r"Processing ends at: (\d+)", r.stdout)
    assert len(start) == 1 and len(end) == 1 and int(end[0]) >= int(start[0]) > 10**12
    kind = "int" if alg in ("BFS", "WCC", "CDLP") else "float"
    ids, vals = graphio.read_result(str(out), kind)
    gids, gvals = golden(name, alg)
    assert np.array_equal(ids, gids)
    assert validator.validate(alg, vals, gvals)
    if alg in ("BFS", "CDLP"):
        assert np.array_equal(vals, gvals)
    if alg in ("PR", "LCC", "SSSP"):
        text = open(out).read().split("\n")[0].split()[1]
        assert text == "infinity" or re.fullmatch(r"\d\.\d{16}e[+-]\d{2}", text), "precision(16) << scientific"


def test_unknown_source_vertex_fails_like_the_reference(tmp_path):
    d, params = prepare(tmp_path, "example-directed", False, False)
    r = subprocess.run([os.path.join(EXE, "bfs"), "--input-dir", str(d), "--output-file", str(tmp_path / "o"),
                        "--directed", "true", "--source-vertex", "12345"], capture_output=True, text=True)
    assert r.returncode != 0 and "Source vertex not found in mapping" in r.stdout   # bfs.cpp:99-102
