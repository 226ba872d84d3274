"""CPU-side checks (no GPU): the C-ABI library loads and exports every symbol the
header declares, the host file formats round-trip, the converter binary writes the
SuiteSparse dump layout, the binaries fail loudly without a device, and the
world_size-2 sharding logic agrees across ranks."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT
from helpers import load_fixture
from ldbc_graphalytics_platforms_graphblas_b200 import graphio, rmat, validator

EXE = os.path.join(ROOT, "bin", "exe")


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__
    __graft_entry__.build()


def test_library_exports_every_declared_symbol():
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    header = open(capi.HEADER_PATH).read()
    declared = set(re.findall(r"^(?:int|void|const char \*)\s*\*?(gx_\w+)\(", header, flags=re.M))
    assert len(declared) >= 28
    L = capi.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in gxb200.h but not exported"
    out = subprocess.run(["nm", "-D", "--defined-only", capi.LIB_PATH], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (gx_\w+)", out))
    assert declared <= exported


def test_no_gpu_means_loud_failure_not_fallback():
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    if capi.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(capi.GxError):
        capi.init(0)
    with pytest.raises(capi.GxError):
        capi.Graph.from_csr(2, np.array([0, 1, 2], dtype=np.uint64), np.array([1, 0], dtype=np.uint32))
    r = subprocess.run([os.path.join(EXE, "pr"), "--input-dir", "/nonexistent", "--output-file", "/tmp/x"],
                       capture_output=True, text=True)
    assert r.returncode != 0


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "ldbc_graphalytics_platforms_graphblas_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")) or f == "Makefile":
                text = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b|liboracle", text, flags=re.M), f


@pytest.mark.parametrize("name", ["example-directed", "example-undirected", "test-wcc-directed"])
@pytest.mark.parametrize("weighted", [False, True])
def test_relabel_and_readers_roundtrip(tmp_path, name, weighted):
    g, params = load_fixture(name)
    if weighted and not params["weighted"]:
        pytest.skip("unweighted fixture")
    graphio.write_vtx_mtx(str(tmp_path), os.path.join(GOLDEN, name + ".v"), os.path.join(GOLDEN, name + ".e"),
                          params["directed"], weighted)
    head = open(tmp_path / "graph.mtx").read().split("\n")
    assert head[0] == "%%MatrixMarket matrix coordinate {} {}".format(
        "real" if weighted else "integer", "general" if params["directed"] else "symmetric")
    assert head[1] == "%%GraphBLAS " + ("GrB_FP64" if weighted else "GrB_BOOL")
    h = graphio.read_vtx_mtx(str(tmp_path))
    assert np.array_equal(h.rowptr, g.rowptr) and np.array_equal(h.colidx, g.colidx)
    assert np.array_equal(h.mapping, g.mapping)
    if weighted:
        assert np.array_equal(h.weights, g.weights)
    # C++ converter (GraphBLAS-free tools/converter.cpp) -> SuiteSparse dump layout, read by the numpy reader
    subprocess.check_call([os.path.join(EXE, "converter"), "--data-dir", str(tmp_path)])
    b = graphio.read_grb(str(tmp_path / "graph.grb"), params["directed"])
    assert np.array_equal(b.rowptr, g.rowptr) and np.array_equal(b.colidx, g.colidx)
    assert (b.weights is None) == (not weighted)
    if weighted:
        assert np.array_equal(b.weights, g.weights)
    assert np.array_equal(np.fromfile(tmp_path / "graph.vtb", dtype="<u8"), g.mapping)
    assert os.path.getsize(tmp_path / "graph.grb") == 512 + 68 + 8 * (g.n + 1) + 8 * g.nnz + (8 * g.nnz if weighted else 1)


@pytest.mark.parametrize("name", ["example-directed", "example-undirected", "test-sssp-directed"])
def test_relabel_drop_in_matches_reference_format(tmp_path, name):
    """bin/py/relabel.py (DuckDB-free, same flags as the reference's) writes the files load-graph.sh expects."""
    g, params = load_fixture(name)
    import sys
    subprocess.check_call([sys.executable, os.path.join(ROOT, "bin", "py", "relabel.py"), "--graph-name", name,
                           "--input-vertex-path", os.path.join(GOLDEN, name + ".v"),
                           "--input-edge-path", os.path.join(GOLDEN, name + ".e"), "--output-path", str(tmp_path),
                           "--weighted", str(params["weighted"]).lower(), "--directed", str(params["directed"]).lower()],
                          stdout=subprocess.DEVNULL)
    h = graphio.read_vtx_mtx(str(tmp_path))
    assert np.array_equal(h.rowptr, g.rowptr) and np.array_equal(h.colidx, g.colidx) and np.array_equal(h.mapping, g.mapping)
    if params["weighted"]:
        assert np.array_equal(h.weights, g.weights)


@pytest.mark.parametrize("symmetric", [False, True])
def test_parallel_mtx_loader_matches_numpy(tmp_path, symmetric, monkeypatch):
    """The C++ loader parses graph.mtx in byte ranges on all host threads and builds the CSR in parallel
    (host/graphio.cpp).  A 16 MB weighted file in random line order with repeated entries, self-loops and
    blank lines must give the same graph.grb with 1 and with 8 threads, equal to the numpy construction
    (duplicates keep the smallest weight, self-loops are dropped, symmetric files are mirrored)."""
    rng = np.random.default_rng(5 + symmetric)
    n, m = 50000, 700000
    src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
    src[:5000], dst[:5000] = src[5000:10000], dst[5000:10000]          # repeated entries with other weights
    w = np.round(rng.random(m) * 100, 6) + 0.000001
    (tmp_path / "graph.vtx").write_text("".join(f"{100 + 3 * i}\n" for i in range(n)))
    lines = [f"{a + 1} {b + 1} {x:.6f}\n" for a, b, x in zip(src, dst, w)]
    lines[1000] += "\n"                                                 # a blank line inside
    kind = "symmetric" if symmetric else "general"
    (tmp_path / "graph.mtx").write_text(f"%%MatrixMarket matrix coordinate real {kind}\n%%GraphBLAS GrB_FP64\n{n} {n} {m}\n" + "".join(lines))
    assert (tmp_path / "graph.mtx").stat().st_size > 12 << 20           # several 4 MB ranges
    out = {}
    for threads in ("1", "8"):
        monkeypatch.setenv("GX_LOADER_THREADS", threads)
        subprocess.check_call([os.path.join(EXE, "converter"), "--data-dir", str(tmp_path)], stdout=subprocess.DEVNULL)
        out[threads] = (tmp_path / "graph.grb").read_bytes()
    assert out["1"] == out["8"]
    g = graphio.read_grb(str(tmp_path / "graph.grb"), directed=not symmetric)
    ref = graphio.csr_from_edges(n, src, dst, np.array([float(f"{x:.6f}") for x in w]), not symmetric)
    assert np.array_equal(g.rowptr, ref.rowptr) and np.array_equal(g.colidx, ref.colidx) and np.array_equal(g.weights, ref.weights)


def test_numpy_grb_writer_matches_cpp_reader_layout(tmp_path):
    g = rmat.rmat_graph(8, directed=True, weighted=True)
    graphio.write_graph_dir(str(tmp_path), g, binary=True)
    b = graphio.read_grb(str(tmp_path / "graph.grb"), True)
    assert np.array_equal(b.colidx, g.colidx) and np.array_equal(b.weights, g.weights)
    graphio.write_graph_dir(str(tmp_path), g, binary=False)
    h = graphio.read_vtx_mtx(str(tmp_path))
    assert np.array_equal(h.colidx, g.colidx) and np.array_equal(h.weights, g.weights)


def test_validator_rules():
    assert validator.validate("bfs", [0, 1, 2], [0, 1, 2]) and not validator.validate("bfs", [0, 1, 3], [0, 1, 2])
    assert validator.validate("wcc", [5, 5, 9], [1, 1, 2]) and not validator.validate("wcc", [5, 5, 5], [1, 1, 2])
    assert not validator.validate("cdlp", [1, 2, 3], [1, 1, 2])
    assert validator.validate("pr", [1.0, 2.00001], [1.0, 2.0]) and not validator.validate("pr", [1.0, 2.001], [1.0, 2.0])
    assert validator.validate("sssp", [np.inf, 1.0], [np.inf, 1.0]) and not validator.validate("sssp", [5.0, 1.0], [np.inf, 1.0])
    assert np.array_equal(validator.canonical_min_labels([7, 7, 3, 3, 7]), [0, 0, 2, 2, 0])


def test_result_writer_matches_the_reference_format(tmp_path):
    """gx_result_write (the six Serialize*Result functions, formatted on all host threads): byte-identical to
    `<id> <value>` lines with int64 / uint64 in decimal and precision(16) << scientific doubles, `infinity` for +inf
    (bfs.cpp:59-63, pr.cpp:27-28, sssp.cpp:41-46, cdlp.cpp:48)."""
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    rng = np.random.default_rng(9)
    n = 200_003                                      # several thread blocks, not a multiple of anything
    ids = np.sort(rng.choice(10**12, n, replace=False)).astype(np.uint64)
    lv = rng.integers(0, 50, n).astype(np.int64)
    lv[::7] = np.iinfo(np.int64).max                 # unreached
    lv[3] = -5                                       # (never produced, but the formatter must not mangle it)
    p = tmp_path / "bfs"
    capi.write_result(str(p), ids, lv)
    assert p.read_text() == "".join(f"{i} {v}\n" for i, v in zip(ids.tolist(), lv.tolist()))
    lab = rng.integers(0, n, n).astype(np.uint64)
    capi.write_result(str(p), ids, lab, value_map=ids)
    assert p.read_text() == "".join(f"{i} {ids[v]}\n" for i, v in zip(ids.tolist(), lab.tolist()))
    capi.write_result(str(p), ids, lab)
    assert p.read_text() == "".join(f"{i} {v}\n" for i, v in zip(ids.tolist(), lab.tolist()))
    x = rng.random(n) * 10.0 ** rng.integers(-300, 300, n)
    x[::5] = 0.0
    x[1::11] = np.inf
    x[2] = 1.0
    capi.write_result(str(p), ids, x)
    want = "".join(f"{i} {'infinity' if np.isinf(v) else '%.16e' % v}\n" for i, v in zip(ids.tolist(), x.tolist()))
    assert p.read_text() == want
    capi.write_result(str(p), ids[:0], x[:0])
    assert p.read_text() == ""


def test_relabel_at_scale_and_edge_cases(tmp_path):
    """gx_relabel on a generated edge list (RMAT-18: 4 M lines), byte for byte against the lines the reference's
    relabel.py:64-79 would write; a vertex file that is NOT ascending; an endpoint missing from the vertex file."""
    import time
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    src, dst = rmat.rmat_edges(18)
    ids = np.unique(np.concatenate([src, dst]))
    w = rmat.edge_weights(src, dst, 7)
    vp, ep = tmp_path / "g.v", tmp_path / "g.e"
    np.savetxt(vp, ids, fmt="%d")
    wtxt = np.char.mod("%.17g", w)
    with open(ep, "w") as f:
        f.write("\n".join(" ".join(t) for t in zip(src.astype(str), dst.astype(str), wtxt)) + "\n")
    out = tmp_path / "out"
    t0 = time.perf_counter()
    n, nnz = capi.relabel(str(vp), str(ep), str(out), True, True)
    dt = time.perf_counter() - t0
    assert (n, nnz) == (ids.size, src.size)
    assert dt < 20, f"relabelling 4 M edges took {dt:.1f}s"
    lines = open(out / "graph.mtx").read().split("\n")
    assert lines[:3] == ["%%MatrixMarket matrix coordinate real general", "%%GraphBLAS GrB_FP64", f"{n} {n} {nnz}"]
    s = np.searchsorted(ids, src) + 1
    d = np.searchsorted(ids, dst) + 1
    for k in (0, 1, 12345, nnz // 2, nnz - 1):
        assert lines[3 + k] == f"{s[k]} {d[k]} {wtxt[k]}"
    assert len(lines) == 3 + nnz + 1 and lines[-1] == ""
    assert np.array_equal(np.loadtxt(out / "graph.vtx", dtype=np.uint64), ids)
    # a shuffled vertex file: dense ids follow the file's row order, not the id order
    perm = np.random.default_rng(1).permutation(ids.size)
    np.savetxt(vp, ids[perm], fmt="%d")
    with open(ep, "w") as f:
        f.write("\n".join(f"{a} {b}" for a, b in zip(src[:1000], dst[:1000])) + "\n")
    n2, nnz2 = capi.relabel(str(vp), str(ep), str(out), False, False)
    assert (n2, nnz2) == (ids.size, 1000)
    lines = open(out / "graph.mtx").read().split("\n")
    assert lines[0] == "%%MatrixMarket matrix coordinate integer symmetric" and lines[1] == "%%GraphBLAS GrB_BOOL"
    row_of = np.empty(ids.size, dtype=np.int64)
    row_of[perm] = np.arange(ids.size)
    for k in (0, 17, 999):
        assert lines[3 + k] == f"{row_of[np.searchsorted(ids, src[k])] + 1} {row_of[np.searchsorted(ids, dst[k])] + 1} 1"
    assert np.array_equal(np.loadtxt(out / "graph.vtx", dtype=np.uint64), ids[perm])
    with open(ep, "w") as f:
        f.write(f"{ids[0]} {int(ids.max()) + 12345}\n")
    with pytest.raises(capi.GxError):
        capi.relabel(str(vp), str(ep), str(out), False, True)


def test_device_tokenizer_decimal_conversion_matches_strtod(tmp_path):
    """csrc/decimal_to_double.cuh (Clinger fast path + Eisel-Lemire, what k_mtx_parse runs per weight) compiled for the
    host: every double it returns must be the one strtod returns, and the shapes weight files actually carry (%.17g,
    %.16e, %g of values in (0, 1] and small decimals) must never need the host fallback."""
    csrc = os.path.join(ROOT, "ldbc_graphalytics_platforms_graphblas_b200", "csrc")
    exe = tmp_path / "decimal_check"
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-I", csrc, os.path.join(ROOT, "tests", "decimal_check.cpp"), "-o", str(exe)])
    r = subprocess.run([str(exe), "800000"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout
    lines = r.stdout.strip().split("\n")
    assert lines[0].endswith("bad 0")
    undecided = {int(l.split()[1]): int(l.split()[3]) for l in lines[1:]}
    assert undecided[1] == 0 and undecided[2] == 0 and undecided[5] == 0 and undecided[7] == 0
