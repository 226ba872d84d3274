#!/usr/bin/env python
"""CDLP of this repo against the reference's own CUDA CDLP on the same B200 and the same graph.

    python tools/cdlp_vs_reference.py --scale 22 [--iters 10] [--check]

The reference kernel (oracle/_ref/libcdlp_ref.so = /root/reference/.../cdlp_kernel.cu compiled unchanged for sm_100a
behind oracle/refgpu_shim/graphio.h) is timed over the reference's own window -- cudaMalloc, H2D, kernels, D2H and
the setElement loop, cdlp_cuda.cu:241-243 -- and gx_cdlp over the same window (upload from host arrays + run + download,
what bin/exe/cdlp puts between its Processing lines) as well as device-resident.  Undirected RMAT only: the reference
kernel ignores in-edges (cdlp_kernel.cu:171-180), so it is wrong on directed graphs.  One JSON line."""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=int, default=22)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="also compare both with the CPU oracle")
    args = ap.parse_args()
    so = os.path.join(ROOT, "oracle", "_ref", "libcdlp_ref.so")
    if not os.path.exists(so):
        print(json.dumps({"unavailable": "oracle/_ref/libcdlp_ref.so was not built (needs /root/reference at build time)"}))
        return
    ref = ctypes.CDLL(so)
    u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
    ref.ref_cdlp_gpu.argtypes = [u64p, u64p, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, u64p,
                                 ctypes.POINTER(ctypes.c_double)]
    ref.ref_cdlp_gpu.restype = ctypes.c_int
    capi.init(0)
    g = capi.Graph.rmat(args.scale, directed=False, want_mapping=False)
    n, nnz = g.n, g.nnz
    rp, ci, _ = g.download()
    ev = n + nnz // 2
    # gx_cdlp, device-resident (plan built by the first call)
    g.cdlp(args.iters, out=False)
    gx_dev = min(_timed(lambda: g.cdlp(args.iters, out=False))[0] for _ in range(args.reps))
    gx_labels = g.cdlp(args.iters)
    g.free()
    # gx_cdlp over the reference's window: host CSR -> upload -> run (plan included) -> labels on the host
    pin_rp = capi.PinnedArray((n + 1,), np.uint64); pin_rp.array[:] = rp
    pin_ci = capi.PinnedArray((nnz,), np.uint32); pin_ci.array[:] = ci
    out = np.empty(n, dtype=np.uint64)

    def gx_window():
        h = capi.Graph.from_csr(n, pin_rp.array, pin_ci.array, None, False)
        h.cdlp(args.iters, out=out)
        h.free()
    gx_window()
    gx_win = min(_timed(gx_window)[0] for _ in range(args.reps))
    # the reference kernel, its own window
    aj = ci.astype(np.uint64)
    ref_labels = np.empty(n, dtype=np.uint64)
    ms = ctypes.c_double()
    ref_runs = []
    for _ in range(args.reps + 1):
        rc = ref.ref_cdlp_gpu(rp, aj, n, nnz, 1, args.iters, ref_labels, ctypes.byref(ms))
        if rc != 0:
            print(json.dumps({"error": f"reference cdlp_gpu failed with {rc}"}))
            return
        ref_runs.append(ms.value)
    ref_win = min(ref_runs[1:])
    line = {"graph": f"RMAT-{args.scale} undirected", "n": n, "nnz": nnz, "iterations": args.iters,
            "ref_cuda_cdlp_window_ms": round(ref_win, 3), "gx_cdlp_same_window_ms": round(gx_win, 3),
            "gx_cdlp_device_ms": round(gx_dev, 3), "vs_ref_cuda_cdlp": round(ref_win / gx_win, 2),
            "vs_ref_cuda_cdlp_device_only": round(ref_win / gx_dev, 2),
            "ref_evps": ev / (ref_win * 1e-3), "gx_evps_same_window": ev / (gx_win * 1e-3),
            "labels_equal_gx_vs_ref": bool(np.array_equal(gx_labels, ref_labels)),
            "labels_differing": int((gx_labels != ref_labels).sum()),
            "window": "cudaMalloc + H2D + kernels + D2H (+ the reference's setElement loop), cdlp_cuda.cu:241-243"}
    if args.check:
        import oracle
        oracle.set_threads(os.cpu_count() or 1)
        o = oracle.cdlp(n, rp, ci, False, args.iters)
        line["gx_equals_oracle"] = bool(np.array_equal(gx_labels, o))
        line["ref_differs_from_oracle"] = int((ref_labels != o).sum())
    print(json.dumps(line), flush=True)


def _timed(f):
    capi.sync()
    t0 = time.perf_counter()
    f()
    capi.sync()
    return (time.perf_counter() - t0) * 1e3, None


if __name__ == "__main__":
    main()
