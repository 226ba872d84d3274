"""Synthetic Graph500 RMAT inputs, host (numpy) edition -- SURVEY.md 8(d).

Bit-for-bit the same edge stream as the device generator in csrc/rmat.cu
(counter-based splitmix64 keyed by seed, edge index and level pair), so small
scales can be cross-checked on the CPU and big scales generated in HBM.
"""
import numpy as np

from .graphio import HostGraph, csr_from_edges

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def default_seed(scale):
    return 0x5EED0000 + int(scale)


def splitmix64(x):
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        z = x
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return z ^ (z >> np.uint64(31))


def scramble(v, scale, seed):
    """Bijection on [0, 2^scale): two rounds of odd multiply + xorshift."""
    v = np.asarray(v, dtype=np.uint64)
    mask = np.uint64((1 << scale) - 1)
    sh = np.uint64(scale // 2 + 1)
    k0 = splitmix64(np.uint64(seed) ^ np.uint64(0xA5A5A5A5))
    k1 = splitmix64(np.uint64(seed) ^ np.uint64(0x5A5A5A5A5A))
    with np.errstate(over="ignore"):
        x = v & mask
        x = (x * np.uint64(0x9E3779B97F4A7C15) + k0) & mask
        x ^= x >> sh
        x = (x * np.uint64(0xD1B54A32D192ED03) + k1) & mask
        x ^= x >> sh
    return x


def rmat_edges(scale, seed=None, first=0, count=None, edgefactor=16):
    seed = default_seed(scale) if seed is None else seed
    if count is None:
        count = edgefactor << scale
    i = np.arange(first, first + count, dtype=np.uint64)
    tA = np.uint32(int(0.57 * 4294967296.0))
    tAB = np.uint32(int(0.76 * 4294967296.0))
    tABC = np.uint32(int(0.95 * 4294967296.0))
    s = np.zeros(count, dtype=np.uint64)
    d = np.zeros(count, dtype=np.uint64)
    h = None
    with np.errstate(over="ignore"):
        for l in range(scale):
            if l % 2 == 0:
                h = splitmix64(np.uint64(seed) + (i << np.uint64(5)) + np.uint64(l >> 1))
            r = ((h >> np.uint64(32)) if (l & 1) else (h & np.uint64(0xFFFFFFFF))).astype(np.uint32)
            q = (r >= tA).astype(np.uint64) + (r >= tAB) + (r >= tABC)
            s = (s << np.uint64(1)) | (q >> np.uint64(1))
            d = (d << np.uint64(1)) | (q & np.uint64(1))
    return scramble(s, scale, seed), scramble(d, scale, seed)


def edge_weights(a, b, seed):
    """FP64 weight in (0,1], symmetric in (a,b); keyed on ORIGINAL ids."""
    a = np.asarray(a, dtype=np.uint64)
    b = np.asarray(b, dtype=np.uint64)
    lo, hi = np.minimum(a, b), np.maximum(a, b)
    with np.errstate(over="ignore"):
        k = splitmix64(np.uint64(seed) ^ np.uint64(0x57E1687))
        h = splitmix64(k ^ ((lo * np.uint64(0x100000001B3) + hi) & _M64))
    return ((h >> np.uint64(11)) + np.uint64(1)).astype(np.float64) * (1.0 / 9007199254740992.0)


def rmat_graph(scale, directed, weighted=False, seed=None, edgefactor=16):
    """Clean Graphalytics-style graph: self-loops and duplicates removed,
    isolated ids dropped, dense id = rank of the original (scrambled) id."""
    seed = default_seed(scale) if seed is None else seed
    src, dst = rmat_edges(scale, seed, edgefactor=edgefactor)
    keep = src != dst
    src, dst = src[keep], dst[keep]
    ids = np.unique(np.concatenate([src, dst]))
    s = np.searchsorted(ids, src)
    d = np.searchsorted(ids, dst)
    w = edge_weights(src, dst, seed) if weighted else None
    return csr_from_edges(ids.size, s, d, w, directed, mapping=ids)


def max_out_degree_vertex(g):
    """BFS/SSSP source convention: max out-degree, ties -> smallest dense id."""
    deg = np.diff(g.rowptr.astype(np.int64))
    return int(np.argmax(deg)) if deg.size else 0
