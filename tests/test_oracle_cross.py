"""Cross-checks of the oracle beyond the goldens: independent scipy/numpy
restatements on random and RMAT graphs (sizes the CPU finishes in seconds)."""
import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.csgraph as csg

import oracle
from ldbc_graphalytics_platforms_graphblas_b200 import rmat, validator
from ldbc_graphalytics_platforms_graphblas_b200.graphio import csr_from_edges


def random_graph(n, m, directed, weighted, seed):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    w = rng.random(m) + 1e-3 if weighted else None
    return csr_from_edges(n, src, dst, w, directed)


def to_scipy(g, weights=False):
    data = g.weights if (weights and g.weights is not None) else np.ones(g.nnz)
    return sp.csr_matrix((data, g.colidx.astype(np.int64), g.rowptr.astype(np.int64)), shape=(g.n, g.n))


CASES = [(50, 120, True, 1), (50, 120, False, 2), (400, 3000, True, 3), (400, 3000, False, 4), (1000, 900, True, 5)]


@pytest.mark.parametrize("n,m,directed,seed", CASES)
def test_bfs_vs_scipy(n, m, directed, seed):
    g = random_graph(n, m, directed, False, seed)
    src = rmat.max_out_degree_vertex(g)
    d = csg.shortest_path(to_scipy(g), method="D", unweighted=True, indices=src)
    ref = np.full(g.n, np.iinfo(np.int64).max, dtype=np.int64)
    ref[np.isfinite(d)] = d[np.isfinite(d)].astype(np.int64)
    assert np.array_equal(oracle.bfs(g.n, g.rowptr, g.colidx, src), ref)


@pytest.mark.parametrize("n,m,directed,seed", CASES)
def test_sssp_vs_scipy(n, m, directed, seed):
    g = random_graph(n, m, directed, True, seed)
    src = rmat.max_out_degree_vertex(g)
    ref = csg.dijkstra(to_scipy(g, True), indices=src)
    out = oracle.sssp(g.n, g.rowptr, g.colidx, g.weights, src)
    assert np.array_equal(np.isinf(out), np.isinf(ref))
    fin = np.isfinite(ref)
    assert np.allclose(out[fin], ref[fin], rtol=1e-12, atol=0)


@pytest.mark.parametrize("n,m,directed,seed", CASES)
def test_wcc_vs_scipy(n, m, directed, seed):
    g = random_graph(n, m, directed, False, seed)
    _, lab = csg.connected_components(to_scipy(g), directed=True, connection="weak")
    out = oracle.wcc(g.n, g.rowptr, g.colidx, g.directed)
    assert np.array_equal(out, validator.canonical_min_labels(lab))


@pytest.mark.parametrize("n,m,directed,seed", CASES)
def test_pagerank_vs_dense(n, m, directed, seed):
    g = random_graph(n, m, directed, False, seed)
    A = to_scipy(g).toarray()
    outdeg = A.sum(1)
    d = float(np.float32(0.85))
    r = np.full(n, 1.0 / n)
    for _ in range(7):
        sink = r[outdeg == 0].sum()
        w = np.where(outdeg > 0, r * d / np.maximum(outdeg, 1), 0.0)
        r = (1 - d) / n + d * sink / n + A.T @ w
    out = oracle.pagerank(g.n, g.rowptr, g.colidx, 0.85, 7)
    assert np.allclose(out, r, rtol=1e-12, atol=0)
    assert abs(out.sum() - 1.0) < 1e-12


@pytest.mark.parametrize("n,m,directed,seed", CASES[:4])
def test_lcc_vs_dense(n, m, directed, seed):
    g = random_graph(n, m, directed, False, seed)
    A = to_scipy(g).toarray() > 0
    U = A | A.T
    ref = np.zeros(n)
    for v in range(n):
        nb = np.nonzero(U[v])[0]
        if nb.size >= 2:
            ref[v] = A[np.ix_(nb, nb)].sum() / (nb.size * (nb.size - 1))
    out = oracle.lcc(g.n, g.rowptr, g.colidx, g.directed)
    assert np.allclose(out, ref, rtol=1e-14, atol=0)
    sub = np.array([0, 3, n - 1], dtype=np.uint64)
    part = oracle.lcc(g.n, g.rowptr, g.colidx, g.directed, subset=sub)
    assert np.array_equal(part[sub.astype(int)], out[sub.astype(int)])
    assert np.isnan(np.delete(part, sub.astype(int))).all()


@pytest.mark.parametrize("n,m,directed,seed", CASES[:4])
def test_cdlp_vs_python(n, m, directed, seed):
    g = random_graph(n, m, directed, False, seed)
    A = to_scipy(g).tocsr()
    AT = A.T.tocsr()
    lab = np.arange(n)
    for _ in range(5):
        new = lab.copy()
        for v in range(n):
            nb = list(A.indices[A.indptr[v]:A.indptr[v + 1]])
            if directed:
                nb += list(AT.indices[AT.indptr[v]:AT.indptr[v + 1]])
            if nb:
                vals, cnt = np.unique(lab[nb], return_counts=True)
                new[v] = vals[np.argmax(cnt)]  # first max == smallest label
        lab = new
    out = oracle.cdlp(g.n, g.rowptr, g.colidx, g.directed, 5)
    assert np.array_equal(out, lab.astype(np.uint64))


def test_transpose_roundtrip():
    g = random_graph(300, 2000, True, True, 7)
    trp, tci, tw = oracle.transpose(g.n, g.rowptr, g.colidx, g.weights)
    rp2, ci2, w2 = oracle.transpose(g.n, trp, tci, tw)
    assert np.array_equal(rp2, g.rowptr) and np.array_equal(ci2, g.colidx) and np.array_equal(w2, g.weights)


@pytest.mark.parametrize("scale", [6, 11])
def test_rmat_generators_agree(scale):
    seed = rmat.default_seed(scale)
    s0, d0 = rmat.rmat_edges(scale, seed)
    s1, d1 = oracle.rmat_edges(scale, seed, 0, 16 << scale)
    assert np.array_equal(s0, s1) and np.array_equal(d0, d1)
    a, b = oracle.rmat_edges(scale, seed, 100, 50)  # any slice is independently reproducible
    assert np.array_equal(a, s0[100:150]) and np.array_equal(b, d0[100:150])
    assert np.array_equal(rmat.edge_weights(s0, d0, seed), oracle.edge_weights(s0, d0, seed))
    w = rmat.edge_weights(s0, d0, seed)
    assert (w > 0).all() and (w <= 1).all()
    assert np.array_equal(w, rmat.edge_weights(d0, s0, seed))


def test_scramble_is_a_bijection():
    for scale in (1, 5, 10, 13):
        x = rmat.scramble(np.arange(1 << scale), scale, 42)
        assert np.array_equal(np.sort(x), np.arange(1 << scale, dtype=np.uint64))


def test_rmat_graph_is_clean():
    g = rmat.rmat_graph(10, directed=False, weighted=True)
    rows = np.repeat(np.arange(g.n), np.diff(g.rowptr.astype(np.int64)))
    assert (rows != g.colidx).all()
    key = rows.astype(np.int64) * g.n + g.colidx
    assert (np.diff(key) > 0).all(), "sorted, no duplicates"
    assert (np.diff(g.rowptr.astype(np.int64)) > 0).all(), "no isolated vertices"
    T = sp.csr_matrix((g.weights, g.colidx.astype(np.int64), g.rowptr.astype(np.int64)), shape=(g.n, g.n))
    assert (abs(T - T.T)).nnz == 0, "symmetric incl. weights"
