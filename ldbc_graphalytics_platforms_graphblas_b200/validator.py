"""Graphalytics output validation rules without the JVM.

The validator itself lives in graphalytics-core (not in the reference tree);
the three rules are restated from the Graphalytics specification:
exact match (BFS), equivalence of partitions (WCC, CDLP) and epsilon match
(PR, LCC, SSSP; relative 1e-4, infinities must agree).
"""
import numpy as np

EPSILON = 1e-4
RULE = {"bfs": "exact", "wcc": "equivalence", "cdlp": "equivalence",
        "pr": "epsilon", "lcc": "epsilon", "sssp": "epsilon"}


def exact(out, ref):
    return bool(np.array_equal(np.asarray(out), np.asarray(ref)))


def equivalence(out, ref):
    """Same partition: a bijection between the label sets maps out onto ref."""
    out = np.asarray(out)
    ref = np.asarray(ref)
    if out.shape != ref.shape:
        return False
    fwd, bwd = {}, {}
    for a, b in zip(out.tolist(), ref.tolist()):
        if fwd.setdefault(a, b) != b or bwd.setdefault(b, a) != a:
            return False
    return True


def epsilon(out, ref, eps=EPSILON):
    out = np.asarray(out, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if out.shape != ref.shape:
        return False
    inf_o, inf_r = np.isinf(out), np.isinf(ref)
    if not np.array_equal(inf_o, inf_r):
        return False
    fin = ~inf_r
    o, r = out[fin], ref[fin]
    scale = np.maximum(np.abs(r), np.abs(o))
    ok = np.abs(o - r) <= eps * np.where(scale == 0, 1.0, scale)
    return bool(np.all(ok))


def validate(algorithm, out, ref):
    rule = RULE[algorithm.lower()]
    return {"exact": exact, "equivalence": equivalence, "epsilon": epsilon}[rule](out, ref)


def canonical_min_labels(labels):
    """Relabel a partition so every member carries the smallest dense index of
    its block -- the representative LAGraph's FastSV returns (wcc.cpp:31-34)."""
    labels = np.asarray(labels)
    _, inv = np.unique(labels, return_inverse=True)
    mins = np.full(inv.max() + 1 if inv.size else 0, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(mins, inv, np.arange(labels.size, dtype=np.int64))
    return mins[inv].astype(np.uint64)
