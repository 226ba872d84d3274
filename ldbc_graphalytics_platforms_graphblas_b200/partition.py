"""1-D row partition rule of the multi-GPU path, host-side mirror of
csrc/comm.cu (k_partition / make_even_partition): contiguous row blocks
balanced by entry count, interior boundaries rounded down to multiples of 32
so that BFS bitmap words have exactly one owner."""
import numpy as np


def balanced_bounds(rowptr, nranks, rowptr2=None):
    off = np.asarray(rowptr, dtype=np.uint64)
    if rowptr2 is not None:
        off = off + np.asarray(rowptr2, dtype=np.uint64)
    n = off.size - 1
    total = int(off[n])
    b = [0]
    for r in range(1, nranks):
        target = total * r // nranks
        lo = int(np.searchsorted(off[:n], np.uint64(target), side="left"))
        b.append(max(b[-1], lo & ~31))
    b.append(n)
    return b


def even_bounds(count, nranks, align=1):
    return [0] + [count * r // nranks // align * align for r in range(1, nranks)] + [count]


def upload_slices(count, nranks, align=64):
    """Mirror of graph.cu upload_array: rank r uploads [r * per, (r + 1) * per) over PCIe, every rank uploads the
    tail [main, count), and the equal slices are exchanged by one all-gather.  Returns (per, main)."""
    per = count // nranks // align * align
    return per, per * nranks


def column_slices(n, nranks):
    """Mirror of graph.cu transpose_partitioned: the vertex (column) range whose in-edges rank r sorts."""
    return even_bounds(n, nranks, align=32)
