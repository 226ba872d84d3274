"""ctypes front-end of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(ldbc_graphalytics_platforms_graphblas_b200) never imports this module.

Parity status: PINNED -- oracle.c reproduces all 24 golden output files the
reference ships under example-data-sets/graphs/ (tests/test_oracle_golden.py).
The reference's own arithmetic (SuiteSparse:GraphBLAS 7.4.4 + LAGraph dev) is
not vendored in /root/reference and not installed, so there is no oracle/_ref.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = ctypes.CDLL(so)
        vp = ctypes.c_void_p
        L.oracle_num_threads.restype = ctypes.c_int
        L.oracle_set_threads.argtypes = [ctypes.c_int]
        L.oracle_transpose.argtypes = [ctypes.c_uint64, _u64p, _u32p, vp, _u64p, _u32p, vp]
        L.oracle_bfs.argtypes = [ctypes.c_uint64, _u64p, _u32p, ctypes.c_uint64, _i64p]
        L.oracle_pagerank.argtypes = [ctypes.c_uint64, _u64p, _u32p, vp, vp, ctypes.c_double, ctypes.c_int, _f64p]
        L.oracle_wcc.argtypes = [ctypes.c_uint64, _u64p, _u32p, vp, vp, ctypes.c_int, _u64p]
        L.oracle_cdlp.argtypes = [ctypes.c_uint64, _u64p, _u32p, vp, vp, ctypes.c_int, ctypes.c_int, _u64p]
        L.oracle_lcc.argtypes = [ctypes.c_uint64, _u64p, _u32p, vp, vp, ctypes.c_int, vp, ctypes.c_uint64, _f64p]
        L.oracle_sssp.argtypes = [ctypes.c_uint64, _u64p, _u32p, _f64p, ctypes.c_uint64, _f64p]
        L.oracle_scramble.argtypes = [ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64]
        L.oracle_scramble.restype = ctypes.c_uint64
        L.oracle_rmat_edges.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, _u64p, _u64p]
        L.oracle_edge_weights.argtypes = [ctypes.c_uint64, _u64p, _u64p, ctypes.c_uint64, _f64p]
        L.oracle_rmat_csr.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, _u64p, _u64p, _u32p, _u64p,
                                      ctypes.POINTER(ctypes.c_uint64)]
        L.oracle_rmat_csr.restype = ctypes.c_int64
        for f in ("oracle_transpose", "oracle_bfs", "oracle_pagerank", "oracle_wcc", "oracle_cdlp",
                  "oracle_lcc", "oracle_sssp"):
            getattr(L, f).restype = ctypes.c_int
        _LIB = L
    return _LIB


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _chk(rc, what):
    if rc != 0:
        raise RuntimeError(f"oracle {what} failed with {rc}")


def _csr(rowptr, colidx):
    rp = np.ascontiguousarray(rowptr, dtype=np.uint64)
    ci = np.ascontiguousarray(colidx, dtype=np.uint32)
    if ci.size == 0:
        ci = np.zeros(1, dtype=np.uint32)
    return rp, ci


def num_threads():
    return lib().oracle_num_threads()


def set_threads(t):
    lib().oracle_set_threads(int(t))


def transpose(n, rowptr, colidx, weights=None):
    rp, ci = _csr(rowptr, colidx)
    m = int(rp[n])
    trp = np.zeros(n + 1, dtype=np.uint64)
    tci = np.zeros(max(m, 1), dtype=np.uint32)
    w = tw = None
    if weights is not None:
        w = np.ascontiguousarray(weights, dtype=np.float64)
        tw = np.zeros(max(m, 1), dtype=np.float64)
    _chk(lib().oracle_transpose(n, rp, ci, _ptr(w), trp, tci, _ptr(tw)), "transpose")
    return (trp, tci[:m]) if weights is None else (trp, tci[:m], tw[:m])


def bfs(n, rowptr, colidx, src):
    rp, ci = _csr(rowptr, colidx)
    level = np.empty(n, dtype=np.int64)
    _chk(lib().oracle_bfs(n, rp, ci, int(src), level), "bfs")
    return level


def pagerank(n, rowptr, colidx, damping, iters, transposed=None, float_damping=True):
    """float_damping: round the damping factor through FP32 first, as the
    reference does by passing it to LAGr_PageRankGX(..., float damping, ...)."""
    if float_damping:
        damping = float(np.float32(damping))
    rp, ci = _csr(rowptr, colidx)
    r = np.empty(max(n, 1), dtype=np.float64)
    trp = tci = None
    if transposed is not None:
        trp, tci = _csr(*transposed)
    _chk(lib().oracle_pagerank(n, rp, ci, _ptr(trp), _ptr(tci), float(damping), int(iters), r), "pagerank")
    return r[:n]


def _maybe_t(n, rp, ci, directed, transposed):
    if not directed:
        return None, None
    if transposed is None:
        transposed = transpose(n, rp, ci)
    return _csr(*transposed)


def wcc(n, rowptr, colidx, directed, transposed=None):
    rp, ci = _csr(rowptr, colidx)
    trp, tci = _maybe_t(n, rp, ci, directed, transposed)
    comp = np.empty(max(n, 1), dtype=np.uint64)
    _chk(lib().oracle_wcc(n, rp, ci, _ptr(trp), _ptr(tci), int(bool(directed)), comp), "wcc")
    return comp[:n]


def cdlp(n, rowptr, colidx, directed, itermax, transposed=None):
    rp, ci = _csr(rowptr, colidx)
    trp, tci = _maybe_t(n, rp, ci, directed, transposed)
    lab = np.empty(max(n, 1), dtype=np.uint64)
    _chk(lib().oracle_cdlp(n, rp, ci, _ptr(trp), _ptr(tci), int(bool(directed)), int(itermax), lab), "cdlp")
    return lab[:n]


def lcc(n, rowptr, colidx, directed, transposed=None, subset=None):
    rp, ci = _csr(rowptr, colidx)
    trp, tci = _maybe_t(n, rp, ci, directed, transposed)
    out = np.empty(max(n, 1), dtype=np.float64)
    sub = None
    ns = 0
    if subset is not None:
        sub = np.ascontiguousarray(subset, dtype=np.uint64)
        ns = sub.size
    _chk(lib().oracle_lcc(n, rp, ci, _ptr(trp), _ptr(tci), int(bool(directed)), _ptr(sub), ns, out), "lcc")
    return out[:n]


def sssp(n, rowptr, colidx, weights, src):
    rp, ci = _csr(rowptr, colidx)
    w = np.ascontiguousarray(weights, dtype=np.float64)
    if w.size == 0:
        w = np.zeros(1, dtype=np.float64)
    dist = np.empty(n, dtype=np.float64)
    _chk(lib().oracle_sssp(n, rp, ci, w, int(src), dist), "sssp")
    return dist


def rmat_edges(scale, seed, first, count):
    src = np.empty(count, dtype=np.uint64)
    dst = np.empty(count, dtype=np.uint64)
    lib().oracle_rmat_edges(int(scale), int(seed), int(first), int(count), src, dst)
    return src, dst


def edge_weights(a, b, seed):
    a = np.ascontiguousarray(a, dtype=np.uint64)
    b = np.ascontiguousarray(b, dtype=np.uint64)
    w = np.empty(a.size, dtype=np.float64)
    lib().oracle_edge_weights(a.size, a, b, int(seed), w)
    return w


def rmat_csr(scale, seed, edgefactor=16):
    """Directed benchmark graph built on all host threads: (n, rowptr, colidx, ids) -- the graph
    gx_rmat_create builds on the device (used by the CPU reference arm of bench.py)."""
    nedges = int(edgefactor) << int(scale)
    N = 1 << int(scale)
    ids = np.empty(N, dtype=np.uint64)
    rowptr = np.empty(N + 1, dtype=np.uint64)
    colidx = np.empty(max(nedges, 1), dtype=np.uint32)
    scratch = np.empty(2 * max(nedges, 1), dtype=np.uint64)
    m = ctypes.c_uint64()
    n = lib().oracle_rmat_csr(int(scale), int(seed), nedges, ids, rowptr, colidx, scratch, ctypes.byref(m))
    if n < 0:
        raise RuntimeError(f"oracle rmat_csr failed with {n}")
    del scratch
    return int(n), rowptr[: n + 1].copy(), colidx[: m.value].copy(), ids[:n].copy()
