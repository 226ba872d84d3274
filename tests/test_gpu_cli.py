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


# ------------------------------------------------------------------ gx_graph_load (loader row of the scope table)
@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("name", ["example-directed", "example-undirected", "test-wcc-directed", "test-sssp-undirected"])
def test_gx_graph_load_against_goldens(tmp_path, name, binary):
    """gx_graph_load (.mtx/.vtx and .grb/.vtb -> device CSR in one call) replaces ReadMatrixMarket + ReadMapping
    (graphio.cpp:4-60): the loaded graph must give the golden answers of every algorithm that has one."""
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    capi.init(0)
    for alg in ("BFS", "PR", "WCC", "CDLP", "LCC", "SSSP"):
        if not os.path.exists(os.path.join(GOLDEN, f"{name}-{alg}")):
            continue
        weighted = alg == "SSSP"
        d, params = prepare(tmp_path / alg, name, weighted, binary)
        g = capi.Graph.load(str(d), binary, params["directed"])
        try:
            ids, ref = golden(name, alg)
            assert np.array_equal(g.mapping, ids), "graph.vtx / graph.vtb order"
            dense = {int(v): i for i, v in enumerate(g.mapping)}
            if alg == "BFS":
                out = g.bfs(dense[params["bfs_source"]])
            elif alg == "PR":
                out = g.pagerank(params["pr_damping"], params["pr_iters"])
            elif alg == "WCC":
                out = g.mapping[g.wcc().astype(np.int64)]
            elif alg == "CDLP":
                out = g.mapping[g.cdlp(params["cdlp_iters"]).astype(np.int64)]
            elif alg == "LCC":
                out = g.lcc()
            else:
                assert g.weighted
                out = g.sssp(dense[params["sssp_source"]])
            assert validator.validate(alg, out, ref), alg
        finally:
            g.free()


def test_grb_and_mtx_loaders_agree_on_dirty_input(tmp_path):
    """A .mtx with self-loops and duplicate entries and the .grb the converter writes from it load as the same
    graph (both loaders drop self-loops, sort the rows and merge duplicates keeping the smallest weight)."""
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    capi.init(0)
    d = tmp_path / "dirty"
    d.mkdir()
    edges = [(1, 2, 0.5), (1, 2, 0.25), (2, 2, 9.0), (3, 1, 1.5), (2, 3, 0.75), (4, 4, 1.0), (3, 1, 2.5)]
    with open(d / "graph.mtx", "w") as f:
        f.write(f"%%MatrixMarket matrix coordinate real general\n%%GraphBLAS GrB_FP64\n4 4 {len(edges)}\n")
        for s, t, w in edges:
            f.write(f"{s} {t} {w}\n")
    with open(d / "graph.vtx", "w") as f:
        f.write("10\n20\n30\n40\n")
    subprocess.check_call([os.path.join(EXE, "converter"), "--data-dir", str(d), "--weighted", "true", "--directed", "true"])
    a = capi.Graph.load(str(d), False, True)
    b = capi.Graph.load(str(d), True, True)
    try:
        ra, rb = a.download(), b.download()
        for x, y in zip(ra, rb):
            assert np.array_equal(x, y)
        assert ra[0].tolist() == [0, 1, 2, 3, 3] and ra[1].tolist() == [1, 2, 0] and ra[2].tolist() == [0.25, 0.75, 1.5]
        assert np.array_equal(a.mapping, [10, 20, 30, 40]) and np.array_equal(b.mapping, a.mapping)
    finally:
        a.free(); b.free()
    # the same entries as a raw .grb dump (jumbled rows, duplicates and self-loops left in, as GxB could hand them out)
    hg = graphio.HostGraph(4, np.array([0, 2, 4, 6, 7], dtype=np.uint64), np.array([1, 1, 2, 1, 0, 0, 3], dtype=np.uint32),
                           np.array([0.5, 0.25, 0.75, 9.0, 1.5, 2.5, 1.0]), True, np.array([10, 20, 30, 40], dtype=np.uint64))
    d2 = tmp_path / "dirty_grb"
    graphio.write_graph_dir(str(d2), hg, binary=True)
    c = capi.Graph.load(str(d2), True, True)
    try:
        rc = c.download()
        assert rc[0].tolist() == [0, 1, 2, 3, 3] and rc[1].tolist() == [1, 2, 0] and rc[2].tolist() == [0.25, 0.75, 1.5]
    finally:
        c.free()


# ------------------------------------------------------------------ device tokenizer of graph.mtx (csrc/mtx_device.cu)
def _messy_mtx(path, n, m, symmetric, weighted, rng):
    """A graph.mtx in random line order with repeated entries, self-loops, blank lines, CRLF line ends, leading blanks,
    tabs, several number formats and no newline after the last line."""
    src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
    src[:3000], dst[:3000] = src[3000:6000], dst[3000:6000]      # repeated entries (other weights)
    src[6000:6200] = dst[6000:6200]                              # self-loops
    w = rng.random(m) * 100 + 1e-9
    fmts = ["{:.6f}", "{:.17g}", "{:.16e}", "{:g}", "{!r}", "+{:.3f}", "{:.25f}"]   # the last one has > 19 digits: host fallback
    lines = []
    for k, (a, b, x) in enumerate(zip(src.tolist(), dst.tolist(), w.tolist())):
        val = " 1" if not weighted else " " + fmts[k % len(fmts)].format(x)
        sep = "\t" if k % 97 == 0 else " "
        lead = "  " if k % 1013 == 0 else ""
        end = "\r\n" if k % 31 == 0 else "\n"
        lines.append(f"{lead}{a + 1}{sep}{b + 1}{val}{end}")
        if k % 5003 == 0:
            lines.append("\n")
    body = "".join(lines).rstrip("\n")
    kind = "symmetric" if symmetric else "general"
    field = "real" if weighted else "integer"
    with open(path, "w", newline="") as f:
        f.write(f"%%MatrixMarket matrix coordinate {field} {kind}\n%%GraphBLAS {'GrB_FP64' if weighted else 'GrB_BOOL'}\n"
                f"% a comment line\n{n} {n} {m}\n{body}")


@pytest.mark.parametrize("symmetric,weighted", [(False, False), (True, False), (False, True), (True, True)])
def test_device_and_host_mtx_loaders_build_the_same_graph(tmp_path, monkeypatch, symmetric, weighted):
    """gx_graph_load tokenises graph.mtx on the device by default; GX_LOADER=host selects the host-threaded parser
    (ReadMtxFile).  Both must produce bit-identical CSR arrays -- offsets, columns and FP64 weights (the device
    converts decimals with correct rounding and hands what it cannot decide to strtod)."""
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    capi.init(0)
    rng = np.random.default_rng(11 + 2 * symmetric + weighted)
    n, m = 40000, 300000
    _messy_mtx(tmp_path / "graph.mtx", n, m, symmetric, weighted, rng)
    (tmp_path / "graph.vtx").write_text("".join(f"{7 + 2 * i}\n" for i in range(n)))
    monkeypatch.setenv("GX_LOADER", "host")
    a = capi.Graph.load(str(tmp_path), False, not symmetric)
    monkeypatch.delenv("GX_LOADER")
    b = capi.Graph.load(str(tmp_path), False, not symmetric)
    c = capi.Graph.load_mtx(str(tmp_path / "graph.mtx"), not symmetric, capi.GX_CACHE_AT)
    try:
        ra, rb, rc = a.download(), b.download(), c.download()
        assert a.weighted == weighted and b.weighted == weighted
        for x, y, z in zip(ra, rb, rc):
            if x is None:
                assert y is None and z is None
            else:
                assert np.array_equal(x, y) and np.array_equal(x, z)
        assert np.array_equal(a.mapping, b.mapping)
        assert np.array_equal(a.wcc(), c.wcc())
    finally:
        a.free(); b.free(); c.free()


def test_device_mtx_loader_rejects_what_the_host_parser_rejects(tmp_path):
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    capi.init(0)
    head = "%%MatrixMarket matrix coordinate real general\n%%GraphBLAS GrB_FP64\n"
    cases = {"too few entries": head + "3 3 3\n1 2 0.5\n2 3 0.5\n",
             "too many entries": head + "3 3 1\n1 2 0.5\n2 3 0.5\n",
             "out of range": head + "3 3 2\n1 2 0.5\n2 4 0.5\n",
             "zero index": head + "3 3 2\n0 2 0.5\n2 3 0.5\n",
             "no value": head + "3 3 2\n1 2\n2 3 0.5\n",
             "garbage": head + "3 3 2\n1 x 0.5\n2 3 0.5\n",
             "not square": head + "3 4 1\n1 2 0.5\n",
             "not matrix market": "1 2 0.5\n"}
    for what, text in cases.items():
        p = tmp_path / "bad.mtx"
        p.write_text(text)
        with pytest.raises(capi.GxError):
            capi.Graph.load_mtx(str(p), True).free()
    # and the smallest good ones: an empty body, a single entry without a trailing newline
    p = tmp_path / "ok.mtx"
    p.write_text(head + "3 3 0\n")
    g = capi.Graph.load_mtx(str(p), True)
    assert g.n == 3 and g.nnz == 0
    g.free()
    p.write_text(head + "3 3 1\n3 1 1e-3")
    g = capi.Graph.load_mtx(str(p), True)
    rp, ci, w = g.download()
    assert rp.tolist() == [0, 0, 0, 1] and ci.tolist() == [0] and w.tolist() == [0.001]
    g.free()
