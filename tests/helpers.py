"""Shared test helpers: load fixture graphs, run the oracle, compare."""
import os

import numpy as np

from ldbc_graphalytics_platforms_graphblas_b200 import graphio

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "graphs")


def load_fixture(name):
    params = graphio.graph_params(os.path.join(GOLDEN, name + ".properties"))
    g = graphio.read_ve(os.path.join(GOLDEN, name + ".v"), os.path.join(GOLDEN, name + ".e"),
                        params["directed"], params["weighted"])
    return g, params


def golden(name, alg):
    kind = "int" if alg in ("BFS", "WCC", "CDLP") else "float"
    return graphio.read_result(os.path.join(GOLDEN, f"{name}-{alg}"), kind)


def oracle_run(oracle, g, params, alg):
    """Run one algorithm of the oracle on a HostGraph; returns values in the
    representation the reference serialises (cdlp/wcc mapped to original ids)."""
    alg = alg.upper()
    n, rp, ci = g.n, g.rowptr, g.colidx
    if alg == "BFS":
        return oracle.bfs(n, rp, ci, g.dense_id(params["bfs_source"]))
    if alg == "PR":
        return oracle.pagerank(n, rp, ci, params["pr_damping"], params["pr_iters"])
    if alg == "WCC":
        return g.mapping[oracle.wcc(n, rp, ci, g.directed).astype(np.int64)]
    if alg == "CDLP":
        return g.mapping[oracle.cdlp(n, rp, ci, g.directed, params["cdlp_iters"]).astype(np.int64)]
    if alg == "LCC":
        return oracle.lcc(n, rp, ci, g.directed)
    if alg == "SSSP":
        return oracle.sssp(n, rp, ci, g.weights, g.dense_id(params["sssp_source"]))
    raise ValueError(alg)


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    fin = np.isfinite(a) & np.isfinite(b)
    if not np.array_equal(np.isfinite(a), np.isfinite(b)):
        return float("inf")
    den = np.maximum(np.abs(b[fin]), 1e-300)
    return float(np.max(np.abs(a[fin] - b[fin]) / den)) if fin.any() else 0.0
