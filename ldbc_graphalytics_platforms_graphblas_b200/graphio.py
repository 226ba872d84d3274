"""Host-side file formats on either side of the hot path (numpy, no GPU).

Mirrors, DuckDB- and GraphBLAS-free, the formats the unchanged reference stages
produce and consume:

* ``X.v`` / ``X.e`` / ``X.properties``  -- raw Graphalytics input
* ``graph.vtx`` / ``graph.mtx``         -- bin/py/relabel.py:52-79
* ``graph.vtb`` / ``graph.grb``         -- tools/converter.cpp:38-56, graphio.h:310-685
* result files ``<orig id> <value>``    -- Serialize*Result in the six wrappers

The C++ loader (csrc/host/graphio.cpp) is the shipped reader; this module is
the test/bench-side writer and an independent reader used to cross-check it.
"""
import os
import struct

import numpy as np

GRB_HEADER_LEN = 512  # LAGRAPH_BIN_HEADER, graphio.h:549-567


class HostGraph:
    """CSR by row over dense ids 0..n-1 (row i = out-neighbours of i).

    ``rowptr`` uint64[n+1], ``colidx`` uint32[m] sorted within rows,
    ``weights`` float64[m] or None, ``mapping`` uint64[n] = original ids
    (graph.vtx order).  Undirected graphs are stored symmetric.
    """

    def __init__(self, n, rowptr, colidx, weights, directed, mapping=None):
        self.n = int(n)
        self.rowptr = np.ascontiguousarray(rowptr, dtype=np.uint64)
        self.colidx = np.ascontiguousarray(colidx, dtype=np.uint32)
        self.weights = None if weights is None else np.ascontiguousarray(weights, dtype=np.float64)
        self.directed = bool(directed)
        self.mapping = (np.arange(self.n, dtype=np.uint64) if mapping is None
                        else np.ascontiguousarray(mapping, dtype=np.uint64))

    @property
    def nnz(self):
        return int(self.rowptr[self.n])

    @property
    def num_edges(self):
        """|E| the way Graphalytics counts it: an undirected edge once."""
        return self.nnz if self.directed else self.nnz // 2

    def dense_id(self, original):
        idx = np.nonzero(self.mapping == np.uint64(original))[0]
        if idx.size == 0:
            raise KeyError(f"vertex {original} not in mapping")
        return int(idx[0])


def csr_from_edges(n, src, dst, w=None, directed=True, mapping=None, dedupe=True):
    """Build a HostGraph from dense 0-based edge arrays.

    Undirected input lists each edge once (any orientation); it is mirrored.
    Self-loops are dropped; duplicates keep one entry (min weight), which is
    the defensive behaviour documented for the native loader.
    """
    src = np.asarray(src, dtype=np.int64)
    dst = np.asarray(dst, dtype=np.int64)
    if w is not None:
        w = np.asarray(w, dtype=np.float64)
    if not directed:
        src, dst = np.concatenate([src, dst]), np.concatenate([dst, src])
        if w is not None:
            w = np.concatenate([w, w])
    keep = src != dst
    src, dst = src[keep], dst[keep]
    if w is not None:
        w = w[keep]
    key = src * np.int64(max(n, 1)) + dst
    if w is not None:
        order = np.lexsort((w, key))
    else:
        order = np.argsort(key, kind="stable")
    key = key[order]
    if dedupe and key.size:
        first = np.ones(key.size, dtype=bool)
        first[1:] = key[1:] != key[:-1]
        order = order[first]
        key = key[first]
    src, dst = src[order], dst[order]
    if w is not None:
        w = w[order]
    counts = np.bincount(src, minlength=n).astype(np.uint64) if n else np.zeros(0, np.uint64)
    rowptr = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(counts, out=rowptr[1:])
    return HostGraph(n, rowptr, dst.astype(np.uint32), w, directed, mapping)


# --------------------------------------------------------------------------- raw input

def read_properties(path):
    props = {}
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#") or "=" not in line:
                continue
            k, v = line.split("=", 1)
            props[k.strip()] = v.strip()
    return props


def graph_params(props_path):
    """Per-graph parameters the Java driver turns into CLI flags
    (BreadthFirstSearchJob.java:28-37, PageRankJob.java:30-40, ...)."""
    props = read_properties(props_path)
    name = os.path.basename(props_path)[: -len(".properties")]
    pre = f"graph.{name}."
    g = lambda k, d=None: props.get(pre + k, d)
    out = {
        "name": name,
        "directed": g("directed", "false").lower() == "true",
        "weighted": g("edge-properties.names") is not None,
        "algorithms": [a.strip() for a in g("algorithms", "").split(",") if a.strip()],
    }
    for k, conv, dst in (("bfs.source-vertex", int, "bfs_source"), ("sssp.source-vertex", int, "sssp_source"),
                         ("cdlp.max-iterations", int, "cdlp_iters"), ("pr.damping-factor", float, "pr_damping"),
                         ("pr.num-iterations", int, "pr_iters")):
        if g(k) is not None:
            out[dst] = conv(g(k))
    return out


def read_ve(vpath, epath, directed, weighted):
    """X.v / X.e -> HostGraph with dense ids in .v row order (relabel.py:36-47)."""
    ids = np.loadtxt(vpath, dtype=np.uint64, ndmin=1)
    raw = np.loadtxt(epath, dtype=np.float64, ndmin=2) if os.path.getsize(epath) else np.zeros((0, 3))
    order = np.argsort(ids, kind="stable")
    sorted_ids = ids[order]
    def dense(col):
        pos = np.searchsorted(sorted_ids, col.astype(np.uint64))
        if np.any(pos >= ids.size) or np.any(sorted_ids[np.minimum(pos, ids.size - 1)] != col.astype(np.uint64)):
            raise ValueError("edge endpoint missing from vertex file")
        return order[pos]
    src = dense(raw[:, 0]) if raw.size else np.zeros(0, np.int64)
    dst = dense(raw[:, 1]) if raw.size else np.zeros(0, np.int64)
    w = raw[:, 2] if (weighted and raw.shape[1] > 2) else None
    return csr_from_edges(ids.size, src, dst, w, directed, mapping=ids)


# --------------------------------------------------------------------------- vtx / mtx

def write_vtx_mtx(outdir, vpath, epath, directed, weighted):
    """DuckDB-free restatement of bin/py/relabel.py:8-79 (same bytes for the
    files, entries in .e order)."""
    os.makedirs(outdir, exist_ok=True)
    ids = np.loadtxt(vpath, dtype=np.uint64, ndmin=1)
    rank = {int(v): i + 1 for i, v in enumerate(ids)}
    with open(os.path.join(outdir, "graph.vtx"), "w") as f:
        for v in ids:
            f.write(f"{int(v)}\n")
    lines = []
    with open(epath) as f:
        for line in f:
            p = line.split()
            if len(p) < 2:
                continue
            s, d = rank[int(p[0])], rank[int(p[1])]
            lines.append(f"{s} {d} {p[2]}" if weighted else f"{s} {d} 1")
    with open(os.path.join(outdir, "graph.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate {} {}\n".format(
            "real" if weighted else "integer", "general" if directed else "symmetric"))
        f.write("%%GraphBLAS {}\n".format("GrB_FP64" if weighted else "GrB_BOOL"))
        f.write(f"{ids.size} {ids.size} {len(lines)}\n")
        for l in lines:
            f.write(l + "\n")


def write_graph_dir(outdir, g, binary=False):
    """Write a HostGraph as graph.vtx+graph.mtx (or graph.vtb+graph.grb)."""
    os.makedirs(outdir, exist_ok=True)
    if binary:
        g.mapping.astype("<u8").tofile(os.path.join(outdir, "graph.vtb"))
        write_grb(os.path.join(outdir, "graph.grb"), g)
        return
    np.savetxt(os.path.join(outdir, "graph.vtx"), g.mapping, fmt="%d")
    rows = np.repeat(np.arange(g.n, dtype=np.int64), np.diff(g.rowptr.astype(np.int64)))
    cols = g.colidx.astype(np.int64)
    w = g.weights
    if not g.directed:
        keep = rows < cols
        rows, cols = rows[keep], cols[keep]
        if w is not None:
            w = w[keep]
    with open(os.path.join(outdir, "graph.mtx"), "w") as f:
        f.write("%%MatrixMarket matrix coordinate {} {}\n".format(
            "real" if w is not None else "integer", "general" if g.directed else "symmetric"))
        f.write("%%GraphBLAS {}\n".format("GrB_FP64" if w is not None else "GrB_BOOL"))
        f.write(f"{g.n} {g.n} {rows.size}\n")
        if w is None:
            np.savetxt(f, np.stack([rows + 1, cols + 1, np.ones_like(rows)], 1), fmt="%d")
        else:
            for r, c, x in zip(rows + 1, cols + 1, w):
                f.write(f"{r} {c} {float(x)!r}\n")


def read_vtx_mtx(indir):
    """Independent reader of graph.vtx/graph.mtx (cross-checks the C++ loader)."""
    mapping = np.loadtxt(os.path.join(indir, "graph.vtx"), dtype=np.uint64, ndmin=1)
    with open(os.path.join(indir, "graph.mtx")) as f:
        banner = f.readline().split()
        symmetric = banner[4].lower() == "symmetric"
        weighted = banner[3].lower() == "real"
        line = f.readline()
        while line.startswith("%"):
            line = f.readline()
        n, _, nnz = (int(x) for x in line.split())
        raw = np.loadtxt(f, dtype=np.float64, ndmin=2) if nnz else np.zeros((0, 3))
    src = raw[:, 0].astype(np.int64) - 1
    dst = raw[:, 1].astype(np.int64) - 1
    w = raw[:, 2] if weighted else None
    return csr_from_edges(n, src, dst, w, directed=not symmetric, mapping=mapping)


# --------------------------------------------------------------------------- grb / vtb

def write_grb(path, g):
    """SuiteSparse binary dump layout (graphio.h:549-606): 512-byte ASCII
    header, packed scalars, Ap[nvec+1], Ai[nvals] as uint64, then Ax
    (one bool if iso, else nvals FP64)."""
    iso = g.weights is None
    typecode, typesize, tname = (0, 1, "bool") if iso else (10, 8, "double")
    hdr = ("SuiteSparse:GraphBLAS matrix\nv%-25s\nnrows:  %-18d\nncols:  %-18d\nnvec:   %-18d\n"
           "nvals:  %-18d\nformat: %-8s\nsize:   %-18d\ntype:   %-72s\niso:    %1d\n%-210s\n\n"
           % ("7.4.4 (gxb200 writer)", g.n, g.n, g.n, g.nnz, "CSR ", typesize, tname, int(iso), "\n"))
    hb = hdr.encode()[: GRB_HEADER_LEN - 1]
    hb = hb + b" " * (GRB_HEADER_LEN - 1 - len(hb)) + b"\0"
    kind = 2 + (100 if iso else 0)  # GxB_SPARSE (+100 iso), graphio.h:573-577
    with open(path, "wb") as f:
        f.write(hb)
        f.write(struct.pack("<iidQQqQQiQ", 0, kind, 0.0625, g.n, g.n, -1, g.n, g.nnz, typecode, typesize))
        g.rowptr.astype("<u8").tofile(f)
        g.colidx.astype("<u8").tofile(f)
        if iso:
            f.write(b"\x01")
        else:
            g.weights.astype("<f8").tofile(f)


def read_grb(path, directed, mapping=None):
    with open(path, "rb") as f:
        f.read(GRB_HEADER_LEN)
        fmt, kind, _hyper, nrows, _ncols, _nonempty, nvec, nvals, typecode, typesize = struct.unpack(
            "<iidQQqQQiQ", f.read(4 + 4 + 8 + 8 + 8 + 8 + 8 + 8 + 4 + 8))
        iso = kind > 100
        kind -= 100 if iso else 0
        if fmt != 0 or kind not in (0, 2):
            raise NotImplementedError("only sparse CSR .grb files are supported")
        rowptr = np.fromfile(f, dtype="<u8", count=nvec + 1)
        colidx = np.fromfile(f, dtype="<u8", count=nvals)
        weights = None
        if not iso:
            if typecode != 10:
                raise NotImplementedError("non-iso values must be FP64")
            weights = np.fromfile(f, dtype="<f8", count=nvals)
    return HostGraph(nrows, rowptr, colidx.astype(np.uint32), weights, directed, mapping)


# --------------------------------------------------------------------------- results

def read_result(path, kind):
    """Parse a ``<id> <value>`` output/golden file -> (ids uint64[], values).

    kind: 'int' (BFS/WCC/CDLP) or 'float' (PR/LCC/SSSP, accepts `infinity`)."""
    ids, vals = [], []
    with open(path) as f:
        for line in f:
            p = line.split()
            if len(p) != 2:
                continue
            ids.append(int(p[0]))
            if kind == "int":
                vals.append(int(p[1]))
            else:
                vals.append(float("inf") if p[1].lower().startswith("inf") else float(p[1]))
    return (np.array(ids, dtype=np.uint64),
            np.array(vals, dtype=np.int64 if kind == "int" else np.float64))
