"""Parity tests proper: the CUDA path, called through the C ABI, against the
oracle on the same inputs.  Bars (BASELINE.json north_star): BFS levels, WCC
and CDLP labels bit-exact; PageRank <= 1e-6 relative; LCC <= 1e-9 relative
(integer ratio, one rounding); SSSP bit-exact (unique fix-point, see
algo_sssp.cu) -- all far inside the Graphalytics validator's 1e-4."""
import os

import numpy as np
import pytest

import oracle
from conftest import golden_cases
from helpers import golden, load_fixture, rel_err
from ldbc_graphalytics_platforms_graphblas_b200 import rmat, validator
from ldbc_graphalytics_platforms_graphblas_b200.graphio import HostGraph, csr_from_edges

pytestmark = pytest.mark.gpu

PR_TOL = 1e-6
LCC_TOL = 1e-9


@pytest.fixture(scope="module")
def capi():
    from ldbc_graphalytics_platforms_graphblas_b200 import capi as c
    c.init(0)
    return c


def check_all(capi, hg, iters_pr=10, iters_cdlp=10, src=None, what="bfs pr wcc cdlp lcc sssp"):
    g = capi.Graph.from_host(hg)
    n, rp, ci = hg.n, hg.rowptr, hg.colidx
    if src is None:
        src = rmat.max_out_degree_vertex(hg)
    T = oracle.transpose(n, rp, ci) if hg.directed else None
    try:
        # a directed graph without the cached transposed adjacency takes the push-only BFS and the out-entries-only
        # WCC (the state the drop-in binaries run in); after gx_graph_cache(GX_CACHE_AT) both take the paths that
        # use the in-edges -- all four must give the oracle's answer
        if "bfs" in what:
            ref_bfs = oracle.bfs(n, rp, ci, src)
            assert np.array_equal(g.bfs(src), ref_bfs), "bfs"
        if "wcc" in what and hg.directed:
            ref_wcc = oracle.wcc(n, rp, ci, hg.directed, transposed=T)
            assert np.array_equal(g.wcc(), ref_wcc), "wcc (out-entries only)"
        if hg.directed:
            g.cache(capi.GX_CACHE_AT)
            if "bfs" in what:
                assert np.array_equal(g.bfs(src), ref_bfs), "bfs (push/pull)"
        if "pr" in what:
            out = g.pagerank(0.85, iters_pr)
            ref = oracle.pagerank(n, rp, ci, 0.85, iters_pr, transposed=T)
            assert rel_err(out, ref) <= PR_TOL, "pr"
        if "wcc" in what:
            assert np.array_equal(g.wcc(), oracle.wcc(n, rp, ci, hg.directed, transposed=T)), "wcc"
        if "cdlp" in what:
            assert np.array_equal(g.cdlp(iters_cdlp), oracle.cdlp(n, rp, ci, hg.directed, iters_cdlp, transposed=T)), "cdlp"
        if "lcc" in what:
            assert rel_err(g.lcc(), oracle.lcc(n, rp, ci, hg.directed, transposed=T)) <= LCC_TOL, "lcc"
        if "sssp" in what and hg.weights is not None:
            assert np.array_equal(g.sssp(src), oracle.sssp(n, rp, ci, hg.weights, src)), "sssp"
    finally:
        g.free()


# ------------------------------------------------------------------ the reference's own fixtures
@pytest.mark.parametrize("name,alg", golden_cases())
def test_golden_fixture(capi, name, alg):
    hg, params = load_fixture(name)
    ids, ref = golden(name, alg)
    g = capi.Graph.from_host(hg)
    try:
        if alg == "BFS":
            out = g.bfs(hg.dense_id(params["bfs_source"]))
        elif alg == "PR":
            out = g.pagerank(params["pr_damping"], params["pr_iters"])
        elif alg == "WCC":
            out = hg.mapping[g.wcc().astype(np.int64)]
        elif alg == "CDLP":
            out = hg.mapping[g.cdlp(params["cdlp_iters"]).astype(np.int64)]
        elif alg == "LCC":
            out = g.lcc()
        else:
            out = g.sssp(hg.dense_id(params["sssp_source"]))
    finally:
        g.free()
    assert validator.validate(alg, out, ref)
    if alg in ("BFS", "WCC", "CDLP"):
        assert np.array_equal(np.asarray(out, dtype=np.int64), ref)
    elif alg == "PR":
        assert rel_err(out, ref) <= 5e-6
    else:
        assert rel_err(out, ref) <= 1e-11


# ------------------------------------------------------------------ random graphs
def random_graph(n, m, directed, seed, weighted=True):
    rng = np.random.default_rng(seed)
    src = rng.integers(0, n, m)
    dst = rng.integers(0, n, m)
    w = rng.random(m) + 1e-3 if weighted else None
    return csr_from_edges(n, src, dst, w, directed)


@pytest.mark.parametrize("n,m,directed,seed", [
    (1, 0, True, 1), (2, 1, False, 2), (33, 40, True, 3), (64, 300, False, 4), (1000, 5000, True, 5),
    (1000, 5000, False, 6), (5000, 200000, True, 7), (5000, 200000, False, 8), (100000, 150000, True, 9),
])
def test_random_graphs(capi, n, m, directed, seed):
    check_all(capi, random_graph(n, m, directed, seed))


def test_empty_rows_and_isolated_vertices(capi):
    # vertices 40..99 are isolated; n is not a multiple of 32
    hg = random_graph(40, 120, True, 11)
    rp = np.concatenate([hg.rowptr, np.full(60, hg.rowptr[-1], dtype=np.uint64)])
    check_all(capi, HostGraph(100, rp, hg.colidx, hg.weights, True), src=3)


@pytest.mark.parametrize("directed", [True, False])
def test_hub_rows(capi, directed):
    """Star + noise: one vertex with 20000 neighbours exercises the chunked long-row paths
    (PR/WCC/BFS/SSSP CHUNK pieces, CDLP spill tables), mid rows the warp-per-row paths."""
    n = 30000
    rng = np.random.default_rng(5)
    hub = np.zeros(20000, dtype=np.int64)
    leaves = rng.choice(np.arange(1, n), 20000, replace=False)
    mid_src = np.repeat(np.arange(1, 41), 600)           # 40 rows with ~600 entries
    mid_dst = rng.integers(1, n, mid_src.size)
    noise_s = rng.integers(0, n, 60000)
    noise_d = rng.integers(0, n, 60000)
    src = np.concatenate([hub, leaves[:5000], mid_src, noise_s])
    dst = np.concatenate([leaves, hub[:5000], mid_dst, noise_d])
    w = rng.random(src.size) + 1e-3
    check_all(capi, csr_from_edges(n, src, dst, w, directed), src=0)


@pytest.mark.parametrize("directed", [True, False])
@pytest.mark.parametrize("small_frac", [0.02, 0.2, 0.6])
def test_wcc_giant_plus_small_components(capi, directed, small_frac):
    """WCC's sampled start (algo_wcc.cu): a giant component plus many small ones whose smallest ids
    lie INSIDE other components' id ranges.  small_frac 0.02 / 0.2 exercise the rows-outside-S path
    (incl. edges whose smaller endpoint is outside the giant tree), 0.6 the plain-FastSV fallback.
    In the directed case every edge points from the larger to the smaller id or at random, so
    half of the small components are only reachable through in-edges."""
    n = 60000
    rng = np.random.default_rng(int(small_frac * 100) + (7 if directed else 0))
    perm = rng.permutation(n)
    n_small = int(n * small_frac)
    giant, small = perm[n_small:], perm[:n_small]
    # giant: a random tree plus noise, so it is connected but the sample does not see every edge
    gs = giant[1:]
    gd = giant[(rng.random(giant.size - 1) * np.arange(1, giant.size)).astype(np.int64)]
    ns, nd = rng.choice(giant, 3 * giant.size), rng.choice(giant, 3 * giant.size)
    # small components: chains of 1..6 vertices
    cs, cd = [], []
    i = 0
    while i < n_small:
        k = int(rng.integers(1, 7))
        grp = small[i:i + k]
        cs.append(grp[:-1]); cd.append(grp[1:])
        i += k
    # a late bridge: one small chain is attached to the giant by a single edge stored far down a long row
    src = np.concatenate([gs, ns] + cs)
    dst = np.concatenate([gd, nd] + cd)
    if directed:
        flip = rng.random(src.size) < 0.5
        src, dst = np.where(flip, dst, src), np.where(flip, src, dst)
    hg = csr_from_edges(n, src, dst, None, directed)
    check_all(capi, hg, what="wcc")


@pytest.mark.parametrize("directed", [True, False])
def test_cdlp_active_rows_match_full_recompute(capi, directed, monkeypatch):
    """CDLP recomputes only rows with a changed neighbour once few rows change (algo_cdlp.cu);
    the labels must equal both the oracle's and a run with GX_CDLP_ACTIVE=0, for every iteration
    count (the switch happens around iteration 4-5 on RMAT)."""
    hg = rmat.rmat_graph(14, directed=directed)
    n, rp, ci = hg.n, hg.rowptr, hg.colidx
    g = capi.Graph.from_host(hg)
    try:
        for iters in (1, 3, 5, 6, 8, 12, 30):
            ref = oracle.cdlp(n, rp, ci, directed, iters)
            monkeypatch.setenv("GX_CDLP_ACTIVE", "1")
            a = g.cdlp(iters)
            monkeypatch.setenv("GX_CDLP_ACTIVE", "0")
            b = g.cdlp(iters)
            assert np.array_equal(a, ref), f"active rows, {iters} iterations"
            assert np.array_equal(b, ref), f"full recompute, {iters} iterations"
    finally:
        g.free()


@pytest.mark.parametrize("directed", [True, False])
def test_cdlp_hub_tables_across_many_iterations(capi, directed, monkeypatch):
    """Hub rows (> 4096 entries) count their labels in global tables whose slots carry a 4-bit epoch tag instead of being
    cleared (algo_cdlp.cu): 60 iterations on one plan wrap the tag several times, with dense iterations (every piece
    inserts), sparse ones (only the pieces of marked hubs, drawn from the compacted list) and full recomputes mixed.
    Three hubs of different sizes, one of them reached through the in-adjacency only."""
    n = 40000
    rng = np.random.default_rng(77)
    e_s, e_d = [], []
    for hub, deg in ((0, 21000), (1, 9000), (2, 4300)):
        other = rng.choice(np.arange(3, n), deg, replace=False)
        if hub == 1:
            e_s.append(other); e_d.append(np.full(deg, hub))      # in-entries only (directed case)
        else:
            e_s.append(np.full(deg, hub)); e_d.append(other)
    e_s.append(rng.integers(0, n, 120000)); e_d.append(rng.integers(0, n, 120000))
    hg = csr_from_edges(n, np.concatenate(e_s), np.concatenate(e_d), None, directed)
    g = capi.Graph.from_host(hg)
    try:
        for iters, active in ((10, "1"), (3, "1"), (17, "0"), (9, "1"), (2, "0"), (19, "1")):
            monkeypatch.setenv("GX_CDLP_ACTIVE", active)
            ref = oracle.cdlp(hg.n, hg.rowptr, hg.colidx, directed, iters)
            assert np.array_equal(g.cdlp(iters), ref), (iters, active)
    finally:
        g.free()


def test_cdlp_first_iteration_closed_form_and_repeated_entries(capi, monkeypatch):
    """Undirected graphs take iteration 1 in closed form (label = smallest neighbour) unless a row
    repeats an entry; a multigraph (dedupe=False) must fall back to counting.  Both against the
    oracle and against GX_CDLP_FIRST=0."""
    rng = np.random.default_rng(21)
    n, m = 3000, 12000
    src, dst = rng.integers(0, n, m), rng.integers(0, n, m)
    src = np.concatenate([src, src[:3000]])   # 3000 edges listed twice
    dst = np.concatenate([dst, dst[:3000]])
    for dedupe in (True, False):
        hg = csr_from_edges(n, src, dst, None, False, dedupe=dedupe)
        g = capi.Graph.from_host(hg)
        try:
            for iters in (1, 2, 10):
                ref = oracle.cdlp(hg.n, hg.rowptr, hg.colidx, False, iters)
                monkeypatch.setenv("GX_CDLP_FIRST", "1")
                assert np.array_equal(g.cdlp(iters), ref), (dedupe, iters)
                monkeypatch.setenv("GX_CDLP_FIRST", "0")
                assert np.array_equal(g.cdlp(iters), ref), (dedupe, iters, "general kernels")
        finally:
            g.free()


@pytest.mark.parametrize("delta", ["", "1e-3", "0.05", "0.5", "100", "0"])
@pytest.mark.parametrize("light_heavy", ["1", "0"])
def test_sssp_bucket_width_never_changes_a_bit(capi, monkeypatch, delta, light_heavy):
    """delta-stepping with light/heavy entries (algo_sssp.cu): whatever the bucket width -- a thousand
    buckets, one bucket, the default, plain sweeps (0) -- and with or without the light/heavy split
    the distances are Dijkstra's bit for bit.  Hub rows exercise the chunked pieces of both parts."""
    if delta:
        monkeypatch.setenv("GX_SSSP_DELTA", delta)
    else:
        monkeypatch.delenv("GX_SSSP_DELTA", raising=False)
    monkeypatch.setenv("GX_SSSP_LH", light_heavy)
    rng = np.random.default_rng(31)
    n = 20000
    hub = np.zeros(12000, dtype=np.int64)
    leaves = rng.choice(np.arange(1, n), 12000, replace=False)
    ns, nd = rng.integers(0, n, 80000), rng.integers(0, n, 80000)
    src = np.concatenate([hub, ns])
    dst = np.concatenate([leaves, nd])
    w = rng.random(src.size) + 1e-4
    w[:50] = 0.0                                     # zero weights and exact ties
    w[50:100] = 0.25
    for directed in (True, False):
        check_all(capi, csr_from_edges(n, src, dst, w, directed), src=0, what="sssp")
    check_all(capi, rmat.rmat_graph(13, directed=False, weighted=True), what="sssp")


def test_pipelined_upload_with_transposition(capi, monkeypatch):
    """gx_graph_create_csr32_cached(GX_CACHE_AT): the in-edge adjacency is assembled from chunk-local sorted
    runs while the column ids upload (graph.cu).  Its results must equal the plain path's and the oracle's:
    BFS (pull levels read A'), PageRank (sums over A' in row order), CDLP (counts over A and A'); also with
    weights, with jumbled rows, and with an invalid column id in the last chunk."""
    rng = np.random.default_rng(77)
    n, m = 60000, 3_000_000                                  # > 2^21 entries: the pipelined path applies
    hg = csr_from_edges(n, rng.integers(0, n, m), (rng.integers(0, n, m) ** 2 // n), rng.random(m) + 1e-3, True)
    rp, ci, w = hg.rowptr, hg.colidx, hg.weights
    T = oracle.transpose(n, rp, ci)
    src = rmat.max_out_degree_vertex(hg)
    ref_bfs = oracle.bfs(n, rp, ci, src)
    ref_pr = oracle.pagerank(n, rp, ci, 0.85, 5, transposed=T)
    ref_cdlp = oracle.cdlp(n, rp, ci, True, 3, transposed=T)
    ref_sssp = oracle.sssp(n, rp, ci, w, src)

    def check(g):
        try:
            assert np.array_equal(g.bfs(src), ref_bfs)
            assert rel_err(g.pagerank(0.85, 5), ref_pr) <= PR_TOL
            assert np.array_equal(g.cdlp(3), ref_cdlp)
            assert np.array_equal(g.sssp(src), ref_sssp)
        finally:
            g.free()

    check(capi.Graph.from_csr(n, rp, ci, w, True, cache=capi.GX_CACHE_AT))
    monkeypatch.setenv("GX_UPLOAD_PIPELINE", "0")
    check(capi.Graph.from_csr(n, rp, ci, w, True, cache=capi.GX_CACHE_AT))
    monkeypatch.delenv("GX_UPLOAD_PIPELINE")
    # PageRank is bit-identical on both paths: the in-rows list their sources in the same (ascending) order
    a = capi.Graph.from_csr(n, rp, ci, None, True, cache=capi.GX_CACHE_AT)
    b = capi.Graph.from_csr(n, rp, ci, None, True)
    try:
        assert np.array_equal(a.pagerank(0.85, 10), b.pagerank(0.85, 10))
    finally:
        a.free(); b.free()
    # jumbled rows: every row reversed
    ci2, w2 = ci.copy(), w.copy()
    for v in rng.choice(n, 2000, replace=False):
        lo, hi = int(rp[v]), int(rp[v + 1])
        ci2[lo:hi] = ci2[lo:hi][::-1]
        w2[lo:hi] = w2[lo:hi][::-1]
    check(capi.Graph.from_csr(n, rp, ci2, w2, True, cache=capi.GX_CACHE_AT))
    # invalid input is rejected, not crashed on
    bad = ci.copy()
    bad[-5] = n + 7
    with pytest.raises(capi.GxError):
        capi.Graph.from_csr(n, rp, bad, None, True, cache=capi.GX_CACHE_AT)
    # column ids beyond 2^bits_for(n): the chunk sort looks at the low bits only, so the runs are not sorted in the
    # full key -- the merge pass must not run on them (it used to write out of bounds and kill the context)
    for pos, val in ((7, 0xFFFFFFF0), (ci.size // 2, (1 << 20) + 3), (ci.size - 1, 1 << 31)):
        bad = ci.copy()
        bad[pos] = val
        with pytest.raises(capi.GxError):
            capi.Graph.from_csr(n, rp, bad, None, True, cache=capi.GX_CACHE_AT)
    rp_bad = rp.copy()
    rp_bad[n // 2] = rp_bad[n // 2 + 1] + 3
    with pytest.raises(capi.GxError):
        capi.Graph.from_csr(n, rp_bad, ci, None, True, cache=capi.GX_CACHE_AT)
    # ... and the context is still alive afterwards
    check(capi.Graph.from_csr(n, rp, ci, w, True, cache=capi.GX_CACHE_AT))


def test_lcc_every_apex_size_class(capi):
    """A clique: oriented out-degrees run from 0 to n-1 (long and short lists meet in every
    intersection), and the answer is known in closed form: LCC is exactly 1 everywhere."""
    n = 3000
    i, j = np.triu_indices(n, 1)
    hg = csr_from_edges(n, i, j, None, directed=False)
    g = capi.Graph.from_host(hg)
    try:
        out = g.lcc()
    finally:
        g.free()
    assert np.array_equal(out, np.ones(n))


def test_unreachable_hub_in_pull(capi):
    """A vertex with a long in-list that BFS never reaches is scanned by the warp-per-row pull."""
    n = 5000
    rng = np.random.default_rng(9)
    comp = rng.integers(0, 2500, (40000, 2))                       # component A: ids < 2500
    into_hub = np.stack([np.arange(2600, 4600), np.full(2000, 2550)], 1)  # 2000 unreachable sources -> 2550
    e = np.concatenate([comp, into_hub])
    hg = csr_from_edges(n, e[:, 0], e[:, 1], None, True)
    check_all(capi, hg, src=int(np.argmax(np.diff(hg.rowptr.astype(np.int64))[:2500])), what="bfs wcc")


@pytest.mark.parametrize("scale,directed", [(10, True), (10, False), (14, True), (14, False), (16, True), (16, False)])
def test_rmat_small(capi, scale, directed):
    check_all(capi, rmat.rmat_graph(scale, directed, weighted=True))


def test_u64_and_u32_entry_points_agree(capi):
    hg = rmat.rmat_graph(12, True, weighted=True)
    g64 = capi.Graph.from_csr(hg.n, hg.rowptr, hg.colidx.astype(np.uint64), hg.weights, True)
    g32 = capi.Graph.from_csr(hg.n, hg.rowptr, hg.colidx, hg.weights, True)
    try:
        for a, b in zip(g64.download(), g32.download()):
            assert np.array_equal(a, b)
        assert np.array_equal(g64.cdlp(5), g32.cdlp(5))
    finally:
        g64.free()
        g32.free()


def test_unsorted_rows_are_sorted_on_upload(capi):
    hg = rmat.rmat_graph(11, True, weighted=True)
    rng = np.random.default_rng(3)
    ci, w = hg.colidx.copy(), hg.weights.copy()
    rp = hg.rowptr.astype(np.int64)
    for v in range(hg.n):
        p = rng.permutation(rp[v + 1] - rp[v]) + rp[v]
        ci[rp[v]:rp[v + 1]], w[rp[v]:rp[v + 1]] = ci[p], w[p]
    g = capi.Graph.from_csr(hg.n, hg.rowptr, ci, w, True)
    try:
        rp2, ci2, w2 = g.download()
        assert np.array_equal(ci2, hg.colidx) and np.array_equal(w2, hg.weights)
    finally:
        g.free()


def test_invalid_input_is_rejected(capi):
    rp = np.array([0, 1, 2], dtype=np.uint64)
    with pytest.raises(capi.GxError):
        capi.Graph.from_csr(2, rp, np.array([1, 7], dtype=np.uint32))      # column id >= n
    with pytest.raises(capi.GxError):
        capi.Graph.from_csr(2, np.array([0, 2, 1], dtype=np.uint64), np.array([1, 0], dtype=np.uint32))
    g = capi.Graph.from_csr(2, rp, np.array([1, 0], dtype=np.uint32))
    try:
        with pytest.raises(capi.GxError):
            g.bfs(5)                                                        # source out of range
        with pytest.raises(capi.GxError):
            g.sssp(0)                                                       # unweighted graph
    finally:
        g.free()


# ------------------------------------------------------------------ device generator
@pytest.mark.parametrize("scale,directed,weighted", [(10, True, False), (13, False, True), (16, True, True)])
def test_device_rmat_matches_host_generator(capi, scale, directed, weighted):
    hg = rmat.rmat_graph(scale, directed, weighted)
    g = capi.Graph.rmat(scale, directed, weighted)
    try:
        assert (g.n, g.nnz) == (hg.n, hg.nnz)
        rp, ci, w = g.download()
        assert np.array_equal(rp, hg.rowptr) and np.array_equal(ci, hg.colidx)
        assert np.array_equal(g.mapping, hg.mapping)
        if weighted:
            assert np.array_equal(w, hg.weights)
        assert g.max_degree_vertex() == rmat.max_out_degree_vertex(hg)
    finally:
        g.free()


# ------------------------------------------------------------------ benchmark-size graphs
@pytest.fixture(scope="module")
def rmat20(capi):
    out = {}
    for directed in (True, False):
        g = capi.Graph.rmat(20, directed, weighted=True)
        rp, ci, w = g.download()
        out[directed] = (g, HostGraph(g.n, rp, ci, w, directed, g.mapping))
    yield out
    for g, _ in out.values():
        g.free()


@pytest.mark.parametrize("directed", [True, False])
def test_rmat20_against_oracle(capi, rmat20, directed):
    g, hg = rmat20[directed]
    n, rp, ci = hg.n, hg.rowptr, hg.colidx
    src = g.max_degree_vertex()
    T = oracle.transpose(n, rp, ci) if directed else None
    assert np.array_equal(g.bfs(src), oracle.bfs(n, rp, ci, src))
    assert rel_err(g.pagerank(0.85, 10), oracle.pagerank(n, rp, ci, 0.85, 10, transposed=T)) <= PR_TOL
    assert np.array_equal(g.wcc(), oracle.wcc(n, rp, ci, directed, transposed=T))
    assert np.array_equal(g.cdlp(10), oracle.cdlp(n, rp, ci, directed, 10, transposed=T))
    assert np.array_equal(g.sssp(src), oracle.sssp(n, rp, ci, hg.weights, src))
    # LCC: the oracle's cost is quadratic in hub degrees, so spot-check a sample incl. the top hubs
    deg = np.diff(rp.astype(np.int64))
    rng = np.random.default_rng(1)
    sample = np.unique(np.concatenate([np.argsort(deg)[-4:], rng.integers(0, n, 3000)])).astype(np.uint64)
    ref = oracle.lcc(n, rp, ci, directed, transposed=T, subset=sample)
    out = g.lcc()
    assert rel_err(out[sample.astype(np.int64)], ref[sample.astype(np.int64)]) <= LCC_TOL


def test_rmat22_size_independent_properties(capi):
    """BASELINE config [1]: RMAT-22 directed, BFS + PR.  Full-size checks that do not need the
    oracle: PR mass conservation and positivity, BFS level consistency along every edge."""
    g = capi.Graph.rmat(22, True)
    try:
        rp, ci, _ = g.download()
        n = g.n
        src = g.max_degree_vertex()
        pr = g.pagerank(0.85, 10)
        assert abs(pr.sum() - 1.0) < 1e-9 and (pr > 0).all()
        lvl = g.bfs(src)
        assert lvl[src] == 0
        rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp.astype(np.int64)))
        lu, lv = lvl[rows], lvl[ci.astype(np.int64)]
        reach = lu != capi.UNREACHED
        assert (lv[reach] != capi.UNREACHED).all(), "a reached vertex has only reached out-neighbours"
        assert (lv[reach] <= lu[reach] + 1).all(), "levels grow by at most one along an edge"
        # every reached non-source vertex has an in-neighbour exactly one level up
        best = np.full(n, np.iinfo(np.int64).max, dtype=np.int64)
        np.minimum.at(best, ci.astype(np.int64)[reach], lu[reach] + 1)
        others = (lvl != capi.UNREACHED) & (np.arange(n) != src)
        assert np.array_equal(best[others], lvl[others])
        # and the oracle agrees at this size too (a few seconds of CPU)
        assert np.array_equal(lvl, oracle.bfs(n, rp, ci, src))
        assert rel_err(pr, oracle.pagerank(n, rp, ci, 0.85, 10)) <= PR_TOL
    finally:
        g.free()


def test_rmat24_undirected_against_oracle(capi):
    """BASELINE config [2] at full size: WCC (sampled FastSV) + CDLP (active rows, closed-form first iteration)
    on the undirected RMAT-24 graph (8.9 M vertices, 521 M stored entries), plus BFS and SSSP (delta-stepping
    with light/heavy entries) on the same graph -- all bit-exact against the oracle (~25 s of CPU)."""
    g = capi.Graph.rmat(24, False, weighted=True, want_mapping=False)
    try:
        rp, ci, w = g.download()
        n = g.n
        src = g.max_degree_vertex()
        comp = g.wcc()
        assert np.array_equal(comp, oracle.wcc(n, rp, ci, False)), "wcc"
        # size-independent properties of the labelling: a label is the smallest id of its component
        c64 = comp.astype(np.int64)
        assert (c64 <= np.arange(n)).all() and np.array_equal(c64[c64], c64)
        assert np.array_equal(g.cdlp(10), oracle.cdlp(n, rp, ci, False, 10)), "cdlp"
        assert np.array_equal(g.bfs(src), oracle.bfs(n, rp, ci, src)), "bfs"
        dist = g.sssp(src)
        assert np.array_equal(dist, oracle.sssp(n, rp, ci, w, src)), "sssp"
        # fix-point property on a slice of the entries: d(v) <= fl(d(u) + w(u,v))
        lo, hi = int(rp[n // 2]), int(rp[n // 2 + 200000])
        rows = np.repeat(np.arange(n // 2, n // 2 + 200000), np.diff(rp[n // 2:n // 2 + 200001].astype(np.int64)))
        assert (dist[ci[lo:hi].astype(np.int64)] <= dist[rows] + w[lo:hi]).all()
    finally:
        g.free()


def test_rmat22_undirected_lcc(capi):
    """BASELINE config [3]: LCC on the undirected RMAT-22 graph (membership tables, owner-ordered entries).
    A random sample of vertices AND the 32 largest hubs are compared with the oracle (<= 1e-9 relative) -- the hubs are
    where the membership tables and the owner-ordered runs do their work (the oracle walks a hub's neighbours on all
    host threads: seconds per hub); every value must lie in [0, 1] and vertices of degree < 2 must be 0."""
    g = capi.Graph.rmat(22, False, want_mapping=False)
    try:
        rp, ci, _ = g.download()
        n = g.n
        out = g.lcc()
        deg = np.diff(rp.astype(np.int64))
        assert ((out >= 0.0) & (out <= 1.0)).all() and (out[deg < 2] == 0.0).all()
        rng = np.random.default_rng(3)
        hubs = np.argsort(deg)[-32:]
        sample = np.unique(np.concatenate([rng.integers(0, n, 4000), hubs])).astype(np.uint64)
        oracle.set_threads(os.cpu_count() or 1)
        ref = oracle.lcc(n, rp, ci, False, subset=sample)
        idx = sample.astype(np.int64)
        assert rel_err(out[idx], ref[idx]) <= LCC_TOL
        assert rel_err(out[hubs], ref[hubs]) <= LCC_TOL and (out[hubs] > 0).all(), "hubs"
        assert (out[idx] > 0).sum() > 100, "the sample must exercise non-trivial values"
    finally:
        g.free()
