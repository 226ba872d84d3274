"""ctypes binding of libgxb200.so (include/gxb200.h) -- the same C ABI the C++
wrappers under csrc/algorithms/ call.  Used by tests/ and bench.py.

There is no CPU fallback: importing works without a GPU (so that the exported
symbols can be checked), but every compute call needs `init()` to have bound a
B200 and raises GxError otherwise.  A missing library is a hard error.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GX_LIB", os.path.join(_HERE, "lib", "libgxb200.so"))  # GX_LIB: A/B runs of two builds
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "gxb200.h")

GX_OK = 0
GX_CACHE_AT = 1
GX_CACHE_LCC = 2
UNREACHED = np.iinfo(np.int64).max


class GxError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gxb200 error [{code}] {msg}")
        self.code = code


class Timing(ctypes.Structure):
    _fields_ = [("h2d_ms", ctypes.c_double), ("build_ms", ctypes.c_double), ("kernel_ms", ctypes.c_double),
                ("comm_ms", ctypes.c_double), ("d2h_ms", ctypes.c_double), ("algorithmic_bytes", ctypes.c_uint64),
                ("edges_inspected", ctypes.c_uint64), ("kernel_launches", ctypes.c_uint32),
                ("iterations", ctypes.c_uint32)]

    def asdict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(make -C ldbc_graphalytics_platforms_graphblas_b200/csrc). There is no fallback path.")
        L = ctypes.CDLL(LIB_PATH)
        vp, u64, i32, dbl = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_int, ctypes.c_double
        pp = ctypes.POINTER(vp)
        sig = {
            "gx_init": [i32], "gx_finalize": [], "gx_device_count": [], "gx_reserve": [u64],
            "gx_comm_unique_id": [vp], "gx_comm_init": [i32, i32, vp], "gx_comm_destroy": [],
            "gx_graph_create_csr": [pp, u64, u64, vp, vp, vp, i32],
            "gx_graph_create_csr32": [pp, u64, u64, vp, vp, vp, i32],
            "gx_graph_create_csr32_cached": [pp, u64, u64, vp, vp, vp, i32, ctypes.c_uint],
            "gx_graph_load": [pp, ctypes.c_char_p, i32, i32, pp, ctypes.POINTER(u64)],
            "gx_graph_load_mtx": [pp, ctypes.c_char_p, i32, ctypes.c_uint],
            "gx_graph_free": [vp],
            "gx_graph_info": [vp, ctypes.POINTER(u64), ctypes.POINTER(u64), ctypes.POINTER(i32), ctypes.POINTER(i32)],
            "gx_graph_cache": [vp, ctypes.c_uint], "gx_graph_download": [vp, vp, vp, vp],
            "gx_bfs": [vp, u64, vp], "gx_pagerank": [vp, dbl, i32, vp], "gx_wcc": [vp, vp],
            "gx_cdlp": [vp, i32, vp], "gx_lcc": [vp, vp], "gx_sssp": [vp, u64, vp],
            "gx_last_timing": [ctypes.POINTER(Timing)], "gx_timer_start": [], "gx_timer_stop": [ctypes.POINTER(dbl)],
            "gx_sync": [], "gx_flush_l2": [], "gx_profile": [i32], "gx_host_alloc": [pp, u64], "gx_host_free": [vp], "gx_host_register": [vp, u64], "gx_host_unregister": [vp],
            "gx_result_write": [ctypes.c_char_p, i32, vp, vp, u64, vp],
            "gx_relabel": [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, i32, i32, ctypes.POINTER(u64), ctypes.POINTER(u64)],
            "gx_rmat_create": [pp, i32, i32, u64, i32, i32, pp], "gx_graph_max_degree_vertex": [vp, ctypes.POINTER(u64)],
        }
        for name, args in sig.items():
            f = getattr(L, name)
            f.argtypes = args
            f.restype = i32
        L.gx_last_error.restype = ctypes.c_char_p
        L.gx_profile_report.restype = ctypes.c_char_p
        L.gx_free_host.argtypes = [vp]
        L.gx_free_host.restype = None
        _lib = L
    return _lib


def _chk(rc):
    if rc != GX_OK:
        raise GxError(rc, lib().gx_last_error().decode(errors="replace"))


def _p(a):
    return None if a is None else ctypes.c_void_p(a.ctypes.data)


def device_count():
    return lib().gx_device_count()


def init(device=0):
    _chk(lib().gx_init(int(device)))


def finalize():
    _chk(lib().gx_finalize())


def last_timing():
    t = Timing()
    _chk(lib().gx_last_timing(ctypes.byref(t)))
    return t.asdict()


def timer_start():
    _chk(lib().gx_timer_start())


def timer_stop():
    ms = ctypes.c_double()
    _chk(lib().gx_timer_stop(ctypes.byref(ms)))
    return ms.value


def sync():
    _chk(lib().gx_sync())


def flush_l2():
    _chk(lib().gx_flush_l2())


def profile(enable):
    _chk(lib().gx_profile(int(bool(enable))))


def profile_report():
    """{kernel name: (launches, total ms)} of the current profiling session."""
    out = {}
    for line in lib().gx_profile_report().decode().splitlines():
        name, cnt, ms = line.split("\t")
        out[name.strip("()")] = (int(cnt), float(ms))
    return out


def comm_unique_id():
    buf = ctypes.create_string_buffer(128)
    _chk(lib().gx_comm_unique_id(buf))
    return buf.raw


def comm_init(rank, nranks, unique_id):
    buf = ctypes.create_string_buffer(bytes(unique_id), 128)
    _chk(lib().gx_comm_init(int(rank), int(nranks), buf))


def comm_destroy():
    _chk(lib().gx_comm_destroy())


def write_result(path, ids, values, value_map=None):
    """Serialize*Result of the reference wrappers: `<id> <value>` lines (int64 / uint64 / %.16e with `infinity`)."""
    ids = np.ascontiguousarray(ids, dtype=np.uint64)
    values = np.ascontiguousarray(values)
    kind = {np.dtype(np.int64): 0, np.dtype(np.uint64): 1, np.dtype(np.float64): 2}[values.dtype]
    vm = None if value_map is None else np.ascontiguousarray(value_map, dtype=np.uint64)
    _chk(lib().gx_result_write(os.fsencode(path), kind, _p(ids), _p(values), ids.size, _p(vm)))


def relabel(vertex_path, edge_path, out_dir, weighted, directed):
    """bin/py/relabel.py of the reference without DuckDB: writes out_dir/graph.vtx and graph.mtx; returns (n, nnz)."""
    n, nnz = ctypes.c_uint64(), ctypes.c_uint64()
    os.makedirs(out_dir, exist_ok=True)
    _chk(lib().gx_relabel(os.fsencode(vertex_path), os.fsencode(edge_path), os.fsencode(out_dir), int(bool(weighted)),
                          int(bool(directed)), ctypes.byref(n), ctypes.byref(nnz)))
    return n.value, nnz.value


class PinnedArray:
    """numpy view over cudaMallocHost memory (full-speed H2D/D2H for the e2e leg)."""

    def __init__(self, shape, dtype):
        self.dtype = np.dtype(dtype)
        self.nbytes = int(np.prod(shape)) * self.dtype.itemsize
        self._ptr = ctypes.c_void_p()
        _chk(lib().gx_host_alloc(ctypes.byref(self._ptr), max(self.nbytes, 1)))
        buf = (ctypes.c_char * max(self.nbytes, 1)).from_address(self._ptr.value)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=int(np.prod(shape))).reshape(shape)

    def free(self):
        if self._ptr is not None and self._ptr.value:
            self.array = None
            lib().gx_host_free(self._ptr)
            self._ptr = None


class Graph:
    """Device-resident graph; the object the six algorithm calls take
    (the role LAGraph_Graph plays in the reference wrappers)."""

    def __init__(self, handle, mapping=None):
        self._h = handle
        n, m, d, w = ctypes.c_uint64(), ctypes.c_uint64(), ctypes.c_int(), ctypes.c_int()
        _chk(lib().gx_graph_info(self._h, ctypes.byref(n), ctypes.byref(m), ctypes.byref(d), ctypes.byref(w)))
        self.n, self.nnz, self.directed, self.weighted = n.value, m.value, bool(d.value), bool(w.value)
        self.mapping = mapping

    # -- construction ------------------------------------------------------------------------
    @classmethod
    def from_csr(cls, n, rowptr, colidx, weights=None, directed=True, mapping=None, cache=0):
        rp = np.ascontiguousarray(rowptr, dtype=np.uint64)
        ci = np.ascontiguousarray(colidx)
        w = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        nnz = int(rp[n]) if n else 0
        h = ctypes.c_void_p()
        if ci.dtype == np.uint32 and cache:
            # upload and cached structures in one call (the transposition rides along with the upload)
            _chk(lib().gx_graph_create_csr32_cached(ctypes.byref(h), n, nnz, _p(rp), _p(ci), _p(w), int(directed), int(cache)))
        elif ci.dtype == np.uint32:
            _chk(lib().gx_graph_create_csr32(ctypes.byref(h), n, nnz, _p(rp), _p(ci), _p(w), int(directed)))
        else:
            ci = np.ascontiguousarray(ci, dtype=np.uint64)
            _chk(lib().gx_graph_create_csr(ctypes.byref(h), n, nnz, _p(rp), _p(ci), _p(w), int(directed)))
            if cache:
                _chk(lib().gx_graph_cache(h, int(cache)))
        return cls(h, mapping)

    @classmethod
    def from_host(cls, g, use_weights=True):
        return cls.from_csr(g.n, g.rowptr, g.colidx, g.weights if use_weights else None, g.directed, g.mapping)

    @classmethod
    def load(cls, directory, binary, directed):
        h, mp, n = ctypes.c_void_p(), ctypes.c_void_p(), ctypes.c_uint64()
        _chk(lib().gx_graph_load(ctypes.byref(h), os.fsencode(directory), int(binary), int(directed),
                                 ctypes.byref(mp), ctypes.byref(n)))
        mapping = np.ctypeslib.as_array(ctypes.cast(mp, ctypes.POINTER(ctypes.c_uint64)), shape=(n.value,)).copy()
        lib().gx_free_host(mp)
        return cls(h, mapping)

    @classmethod
    def load_mtx(cls, path, directed, cache=0):
        """graph.mtx -> device CSR, the text tokenised on the device (no mapping: dense ids)."""
        h = ctypes.c_void_p()
        _chk(lib().gx_graph_load_mtx(ctypes.byref(h), os.fsencode(path), int(directed), int(cache)))
        return cls(h)

    @classmethod
    def rmat(cls, scale, directed, weighted=False, seed=None, edgefactor=16, want_mapping=True):
        seed = 0x5EED0000 + scale if seed is None else seed
        h, mp = ctypes.c_void_p(), ctypes.c_void_p()
        _chk(lib().gx_rmat_create(ctypes.byref(h), scale, edgefactor, seed, int(directed), int(weighted),
                                  ctypes.byref(mp) if want_mapping else None))
        g = cls(h)
        if want_mapping:
            g.mapping = np.ctypeslib.as_array(ctypes.cast(mp, ctypes.POINTER(ctypes.c_uint64)), shape=(g.n,)).copy()
            lib().gx_free_host(mp)
        return g

    def free(self):
        if self._h is not None:
            _chk(lib().gx_graph_free(self._h))
            self._h = None

    def cache(self, what):
        _chk(lib().gx_graph_cache(self._h, what))

    def download(self):
        rp = np.empty(self.n + 1, dtype=np.uint64)
        ci = np.empty(max(self.nnz, 1), dtype=np.uint32)
        w = np.empty(max(self.nnz, 1), dtype=np.float64) if self.weighted else None
        _chk(lib().gx_graph_download(self._h, _p(rp), _p(ci), _p(w)))
        return rp, ci[: self.nnz], (None if w is None else w[: self.nnz])

    def max_degree_vertex(self):
        v = ctypes.c_uint64()
        _chk(lib().gx_graph_max_degree_vertex(self._h, ctypes.byref(v)))
        return v.value

    @property
    def num_edges(self):
        return self.nnz if self.directed else self.nnz // 2

    # -- the six kernels (out=False: leave the result on the device) -------------------------
    def _run(self, fn, dtype, out, *args):
        if out is False:
            _chk(fn(self._h, *args, None))
            return None
        res = np.empty(max(self.n, 1), dtype=dtype) if out is None else out
        _chk(fn(self._h, *args, _p(res)))
        return res[: self.n]

    def bfs(self, src, out=None):
        return self._run(lib().gx_bfs, np.int64, out, int(src))

    def pagerank(self, damping, iters, out=None):
        return self._run(lib().gx_pagerank, np.float64, out, float(damping), int(iters))

    def wcc(self, out=None):
        return self._run(lib().gx_wcc, np.uint64, out)

    def cdlp(self, itermax, out=None):
        return self._run(lib().gx_cdlp, np.uint64, out, int(itermax))

    def lcc(self, out=None):
        return self._run(lib().gx_lcc, np.float64, out)

    def sssp(self, src, out=None):
        return self._run(lib().gx_sssp, np.float64, out, int(src))
