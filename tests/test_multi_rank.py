"""N > 1 path.  CPU part (gloo, world_size 2): the partition rule and the
exchange pattern (owners compute their row block, all-gather / min-reduce the
replicated state) reproduce the oracle.  GPU part: launches
tests/multi_gpu_check.py under torchrun when the box has >= 2 GPUs."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from ldbc_graphalytics_platforms_graphblas_b200 import partition, rmat

WORKER = r'''
import os, sys
import numpy as np
import torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import oracle
from ldbc_graphalytics_platforms_graphblas_b200 import partition, rmat

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
g = rmat.rmat_graph(11, directed=True, weighted=True)
n, rp, ci = g.n, g.rowptr.astype(np.int64), g.colidx.astype(np.int64)
trp, tci = oracle.transpose(n, g.rowptr, g.colidx)
trp = trp.astype(np.int64); tci = tci.astype(np.int64)

def allgatherv(vec, b):
    parts = [None] * world
    dist.all_gather_object(parts, vec[b[rank]:b[rank + 1]].copy())
    return np.concatenate(parts)

# PageRank: pull over the owned block of in-edge rows, all-gather w, all-reduce the sink mass
b = partition.balanced_bounds(trp, world)
lo, hi = b[rank], b[rank + 1]
d = float(np.float32(0.85))
outdeg = np.diff(rp)
r = np.full(n, 1.0 / n)
for it in range(10):
    sink = torch.tensor([r[lo:hi][outdeg[lo:hi] == 0].sum()], dtype=torch.float64)
    dist.all_reduce(sink)
    w = np.where(outdeg > 0, r / np.where(outdeg > 0, outdeg / d, 1.0), 0.0)
    tele = (1 - d) / n + d * float(sink) / n
    mine = np.array([tele + w[tci[trp[v]:trp[v + 1]]].sum() for v in range(lo, hi)])
    r = r.copy(); r[lo:hi] = mine
    r = allgatherv(r, b)
ref = oracle.pagerank(n, g.rowptr, g.colidx, 0.85, 10)
assert np.max(np.abs(r - ref) / ref) < 1e-12, "partitioned PageRank"

# WCC: hook with the owned rows on a replica, min-reduce the parents, stop together
bo, bi = partition.balanced_bounds(rp, world), partition.balanced_bounds(trp, world)
f = np.arange(n)
while True:
    before = f.copy()
    for (ptr, idx, bb) in ((rp, ci, bo), (trp, tci, bi)):
        for u in range(bb[rank], bb[rank + 1]):
            nb = idx[ptr[u]:ptr[u + 1]]
            if nb.size:
                mn = f[f[nb]].min()
                f[f[u]] = min(f[f[u]], mn); f[u] = min(f[u], mn)
    t = torch.from_numpy(f.copy()); dist.all_reduce(t, op=dist.ReduceOp.MIN); f = t.numpy()
    f = np.minimum(f, f[f])
    ch = torch.tensor([int((f != before).any())]); dist.all_reduce(ch, op=dist.ReduceOp.MAX)
    if not int(ch):
        break
assert np.array_equal(f.astype(np.uint64), oracle.wcc(n, g.rowptr, g.colidx, True)), "partitioned WCC"

# SSSP: relax the owned frontier vertices on a replica, min-reduce, next frontier = what dropped
src = rmat.max_out_degree_vertex(g)
dist_v = np.full(n, np.inf); dist_v[src] = 0.0
frontier = np.array([src])
while frontier.size:
    prev = dist_v.copy()
    for u in frontier:
        if bo[rank] <= u < bo[rank + 1]:
            for e in range(rp[u], rp[u + 1]):
                dist_v[ci[e]] = min(dist_v[ci[e]], prev[u] + g.weights[e])
    t = torch.from_numpy(dist_v.copy()); dist.all_reduce(t, op=dist.ReduceOp.MIN); dist_v = t.numpy()
    frontier = np.nonzero(dist_v < prev)[0]
assert np.array_equal(dist_v, oracle.sssp(n, g.rowptr, g.colidx, g.weights, src)), "partitioned SSSP"

# SSSP with delta-stepping on several GPUs (algo_sssp.cu sssp_multi_delta): every rank keeps a full-length
# distance array that is authoritative for its own block and a filter elsewhere; an improvement that passes the
# local minimum is forwarded to the owner (here: the owner's block takes the minimum over the ranks' copies at the
# end of a round, the other entries stay local); two marks per owned vertex say at which distance it was last
# light- / heavy-expanded; {queue size, smallest waiting distance} are max-/min-reduced every round
delta = 8.0 * g.weights.mean() * n / ci.size
lo_, hi_ = bo[rank], bo[rank + 1]
dl = np.full(n, np.inf); dl[src] = 0.0
ldone = np.full(n, np.inf); hdone = np.full(n, np.inf)
def exchange():
    t = torch.from_numpy(dl.copy()); dist.all_reduce(t, op=dist.ReduceOp.MIN)
    dl[lo_:hi_] = t.numpy()[lo_:hi_]
def round_(heavy, T):
    done = hdone if heavy else ldone
    own = np.arange(lo_, hi_)
    due = own[(dl[own] < done[own]) & (dl[own] < T)]
    wait = own[(dl[own] < done[own]) & (dl[own] >= T)]
    done[due] = dl[due]
    for u in due:
        du = dl[u]
        for e in range(rp[u], rp[u + 1]):
            we = g.weights[e]
            if (we > delta) == heavy:
                dl[ci[e]] = min(dl[ci[e]], du + we)
    exchange()
    q = torch.tensor([float(due.size)]); dist.all_reduce(q, op=dist.ReduceOp.MAX)
    fm = torch.tensor([dl[wait].min() if (wait.size and not heavy) else np.inf], dtype=torch.float64)
    dist.all_reduce(fm, op=dist.ReduceOp.MIN)
    return int(q), float(fm)
T, expanded = delta, False
while True:
    while True:
        qn, fmin = round_(False, T)
        if not qn:
            break
        expanded = True
    if expanded:
        round_(True, T); expanded = False
        continue
    if not np.isfinite(fmin):
        break
    T += delta
    if fmin >= T:
        T = fmin + delta
got = allgatherv(dl, bo)
assert np.array_equal(got, oracle.sssp(n, g.rowptr, g.colidx, g.weights, src)), "multi-GPU delta-stepping SSSP"

# Transposition split by column range (graph.cu transpose_partitioned): a rank keeps the (column, row) pairs of
# its vertex slice in row-major order, sorts them stably by column, and the slices concatenate to A'
cb = partition.column_slices(n, world)
rows = np.repeat(np.arange(n), np.diff(rp))
keep = (ci >= cb[rank]) & (ci < cb[rank + 1])
order = np.argsort(ci[keep], kind="stable")
my_cols, my_rows = ci[keep][order], rows[keep][order]
parts = [None] * world
dist.all_gather_object(parts, (my_cols, my_rows))
all_cols = np.concatenate([p[0] for p in parts]); all_rows = np.concatenate([p[1] for p in parts])
assert np.array_equal(all_rows, tci), "column-split transposition: entries"
assert np.array_equal(np.searchsorted(all_cols, np.arange(n + 1)), trp), "column-split transposition: offsets"

# Block-local transposition (graph.cu transpose_block_local): in-degrees from a histogram of 1/world of the entries,
# summed over the ranks -> offsets of all in-edge rows and the nnz-balanced row blocks, known everywhere before any
# sort; a rank then keeps only the in-edge rows of its own block
es = partition.even_bounds(ci.size, world)
hist = torch.from_numpy(np.bincount(ci[es[rank]:es[rank + 1]], minlength=n).astype(np.int64)); dist.all_reduce(hist)
my_trp = np.concatenate([[0], np.cumsum(hist.numpy())])
assert np.array_equal(my_trp, trp), "block-local transposition: offsets from the in-degree histogram"
bl = partition.balanced_bounds(my_trp, world)
keep = (ci >= bl[rank]) & (ci < bl[rank + 1])
order = np.argsort(ci[keep], kind="stable")
assert np.array_equal(rows[keep][order], tci[trp[bl[rank]]:trp[bl[rank + 1]]]), "block-local transposition: own block"

# PageRank, compact form (algo_pr.cu build_compact / k_pr_push_lists): a rank keeps w only for its own vertices and for
# the sources its rows gather from.  The owner derives what a consumer needs from the replicated out-adjacency (v has an
# out-entry into the consumer's row block), the consumer from its in-edge rows: the same ascending lists on both sides,
# so the exchange is contiguous ranges and only a count matrix is communicated
b = partition.balanced_bounds(trp, world)
lo, hi = b[rank], b[rank + 1]
owner_of = np.searchsorted(np.asarray(b[1:]), np.arange(n), side="right")
send = {c: np.array([v for v in range(lo, hi) if c in set(owner_of[ci[rp[v]:rp[v + 1]]].tolist())], dtype=np.int64)
        for c in range(world) if c != rank}                               # owner side
mark = np.zeros(n, dtype=bool); mark[lo:hi] = True
mark[tci[trp[lo]:trp[hi]]] = True                                         # consumer side
keep_ids = np.nonzero(mark)[0]
local_of = np.full(n, -1); local_of[keep_ids] = np.arange(keep_ids.size)
counts = [None] * world
dist.all_gather_object(counts, {c: int(v.size) for c, v in send.items()})
for o in range(world):
    if o != rank:
        assert counts[o][rank] == int(mark[b[o]:b[o + 1]].sum()), "compact PageRank: owner and consumer disagree"
w_loc = np.zeros(keep_ids.size)
r_own = np.full(hi - lo, 1.0 / n)
r_all = np.full(n, 1.0 / n)
for it in range(10):
    sink = torch.tensor([r_own[outdeg[lo:hi] == 0].sum()], dtype=torch.float64); dist.all_reduce(sink)
    w_own = np.where(outdeg[lo:hi] > 0, r_own / np.where(outdeg[lo:hi] > 0, outdeg[lo:hi] / d, 1.0), 0.0)
    w_loc[local_of[lo:hi]] = w_own
    got = [None] * world
    dist.all_gather_object(got, {c: w_own[v - lo] for c, v in send.items()})   # the pushes of this iteration
    for o in range(world):
        if o != rank:
            w_loc[local_of[b[o]:b[o + 1]][mark[b[o]:b[o + 1]]]] = got[o][rank]   # one contiguous local range per owner
    tele = (1 - d) / n + d * float(sink) / n
    r_own = np.array([tele + w_loc[local_of[tci[trp[v]:trp[v + 1]]]].sum() for v in range(lo, hi)])
r_all[lo:hi] = r_own
r_all = allgatherv(r_all, b)
assert np.max(np.abs(r_all - ref) / ref) < 1e-12, "compact PageRank"

# Upload (graph.cu upload_array): equal slices + a common tail cover the array exactly once
for count in (0, 5, 64 * world, 64 * world + 3, 1000003):
    per, main = partition.upload_slices(count, world)
    got = np.zeros(count, dtype=np.int64)
    got[rank * per:(rank + 1) * per] += 1            # own PCIe slice
    t = torch.from_numpy(got); dist.all_reduce(t); got = t.numpy()   # the all-gather
    got[main:] += 1                                   # tail, uploaded by every rank
    assert (got == 1).all(), "upload slices"

# CDLP with active rows (algo_cdlp.cu): owners recompute only rows with a changed neighbour, owners mark the
# neighbours of their changed rows, the maps are OR-ed; labels equal the oracle's after every iteration
ug = rmat.rmat_graph(10, directed=False)
un, urp, uci = ug.n, ug.rowptr.astype(np.int64), ug.colidx.astype(np.int64)
ub = partition.balanced_bounds(urp, world)
lab = np.arange(un)
active = np.ones(un, dtype=bool)
for it in range(1, 9):
    new = lab.copy()
    for v in range(ub[rank], ub[rank + 1]):
        if active[v] and urp[v + 1] > urp[v]:
            vals, cnts = np.unique(lab[uci[urp[v]:urp[v + 1]]], return_counts=True)
            new[v] = vals[np.argmax(cnts)]             # smallest label among the most frequent
    new = allgatherv(new, ub)
    changed = new != lab
    mark = np.zeros(un, dtype=np.int64)
    for v in range(ub[rank], ub[rank + 1]):
        if changed[v]:
            mark[uci[urp[v]:urp[v + 1]]] = 1
    t = torch.from_numpy(mark); dist.all_reduce(t, op=dist.ReduceOp.MAX); active = t.numpy().astype(bool)
    lab = new
    assert np.array_equal(lab.astype(np.uint64), oracle.cdlp(un, ug.rowptr, ug.colidx, False, it)), f"active-row CDLP, iteration {it}"

# WCC with the sampled start (algo_wcc.cu): union with the first two entries of every row (replicated), the
# giant tree S is frozen, only rows outside S are hooked -- both ways per entry -- by their owners
def find(f, x):
    while f[x] != x:
        x = f[x]
    return x
f = np.arange(n)
for r_ in range(2):
    for u in range(n):
        if rp[u] + r_ < rp[u + 1]:
            a, b2 = find(f, u), find(f, ci[rp[u] + r_])
            if a != b2:
                f[max(a, b2)] = min(a, b2)
    f = np.array([find(f, v) for v in range(n)])
vals, cnts = np.unique(f[(np.arange(1024) * n) // 1024], return_counts=True)
giant = vals[np.argmax(cnts)]
rest = np.nonzero(f != giant)[0]
gp = f.copy()
while True:
    before = f.copy()
    for (ptr, idx, bb) in ((rp, ci, bo), (trp, tci, bo)):
        for u in rest:
            if not (bb[rank] <= u < bb[rank + 1]):
                continue
            for v in idx[ptr[u]:ptr[u + 1]]:
                for (x, y) in ((u, v), (v, u)):        # hook both ways: the S side never looks at the edge again
                    if gp[y] < f[x]:
                        f[f[x]] = min(f[f[x]], gp[y]); f[x] = min(f[x], gp[y])
    t = torch.from_numpy(f.copy()); dist.all_reduce(t, op=dist.ReduceOp.MIN); f = t.numpy()
    f = np.minimum(f, f[f]); gp = f[f]
    ch = torch.tensor([int((f != before).any())]); dist.all_reduce(ch, op=dist.ReduceOp.MAX)
    if not int(ch):
        break
assert np.array_equal(f.astype(np.uint64), oracle.wcc(n, g.rowptr, g.colidx, True)), "sampled WCC"

# PageRank's per-iteration sink-mass sum through the peer mailboxes (k_pr_tele_mail / peer_mail_sum_f64): every rank
# deposits its partial, all ranks add the partials IN RANK ORDER -- the same bits everywhere (an all-reduce may
# associate differently per rank count), and the exchange doubles as the iteration's barrier
rng = np.random.default_rng(100 + rank)
for it in range(20):
    part = float(rng.random() * 10.0 ** rng.integers(-12, 3))
    box = [None] * world
    dist.all_gather_object(box, part)              # slot [r] of every rank's mailbox
    total = 0.0
    for r_ in range(world):
        total += box[r_]
    bits = [None] * world
    dist.all_gather_object(bits, np.float64(total).tobytes())
    assert all(b_ == bits[0] for b_ in bits), "rank-ordered mailbox sum differs between ranks"
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_partition_rule():
    g = rmat.rmat_graph(12, directed=True)
    for nr in (1, 2, 3, 4, 8):
        b = partition.balanced_bounds(g.rowptr, nr)
        assert b[0] == 0 and b[-1] == g.n and len(b) == nr + 1 and all(x <= y for x, y in zip(b, b[1:]))
        assert all(x % 32 == 0 for x in b[1:-1])
        loads = np.diff(g.rowptr.astype(np.int64)[b])
        # balanced by entries up to one 32-row block and the heaviest row
        slack = np.diff(g.rowptr.astype(np.int64)).max() * 33
        assert loads.max() - loads.min() <= 2 * slack + g.nnz // nr // 4
    assert partition.even_bounds(10, 4) == [0, 2, 5, 7, 10]
    assert partition.even_bounds(100, 2, align=32) == [0, 32, 100]
    assert partition.upload_slices(1000, 4) == (192, 768) and partition.upload_slices(100, 4) == (0, 0)
    assert partition.column_slices(100, 2) == [0, 32, 100]


def test_world_size_2_gloo_exchange(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script), ROOT],
                       capture_output=True, text=True, timeout=600, env={**os.environ, "OMP_NUM_THREADS": "2"})
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert r.stdout.count("ok") == 2


@pytest.mark.gpu
def test_multi_gpu_matches_oracle():
    from ldbc_graphalytics_platforms_graphblas_b200 import capi
    ngpu = capi.device_count()
    if ngpu < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if ngpu < 4 else 4
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
                        "--master-addr", "127.0.0.1", "--master-port", "29542",
                        os.path.join(ROOT, "tests", "multi_gpu_check.py"), "16"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "MULTI_GPU_CHECK OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
