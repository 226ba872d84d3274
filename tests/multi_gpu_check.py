"""Multi-GPU parity check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/multi_gpu_check.py [scale] [--algos bfs,pr,...] [--graphs directed,undirected]
Every rank runs the algorithms on the 1-D partitioned path; the results must be identical on all ranks
(SHA-256 of the result arrays) and equal to the oracle on rank 0 (bit-exact for BFS/WCC/CDLP/SSSP, PR within
1e-6, LCC within 1e-9).  The defaults (scale 16, all six algorithms, both graph kinds, upload path) are the quick
check; the BASELINE.json config sizes are run as e.g.
    ... multi_gpu_check.py 24 --algos wcc,cdlp --graphs undirected        (config [2], N = 2, 4, 8)
    ... multi_gpu_check.py 26 --algos sssp --graphs undirected            (config [4], N = 8)
and print one JSON line with the timings next to the verdict."""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch.distributed as dist

os.environ.setdefault("GX_TRANSPOSE_SPLIT", "1")  # exercise the column-split transposition at any world size

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import oracle  # noqa: E402
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("scale", nargs="?", type=int, default=16)
    ap.add_argument("--algos", default="bfs,pr,wcc,cdlp,lcc,sssp")
    ap.add_argument("--graphs", default="directed,undirected")
    ap.add_argument("--no-upload", action="store_true", help="skip the host-array upload path (large scales)")
    ap.add_argument("--reps", type=int, default=1, help="timed repetitions after the checked run")
    args = ap.parse_args()
    scale, algos = args.scale, args.algos.split(",")
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    dist.init_process_group("gloo")
    capi.init(local)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    capi.comm_init(rank, world, uid[0])
    oracle.set_threads(os.cpu_count() or 1)
    failures, report = [], []
    big = scale > 20
    for kind in args.graphs.split(","):
        directed = kind == "directed"
        g = capi.Graph.rmat(scale, directed, weighted="sssp" in algos, want_mapping=False)
        src = g.max_degree_vertex()
        run = {"bfs": lambda: g.bfs(src), "pr": lambda: g.pagerank(0.85, 10), "wcc": lambda: g.wcc(), "cdlp": lambda: g.cdlp(10),
               "lcc": lambda: g.lcc(), "sssp": lambda: g.sssp(src)}
        res, timing = {}, {}
        for alg in algos:
            res[alg] = run[alg]()
            best = None
            for _ in range(args.reps):
                dist.barrier()
                t0 = time.perf_counter()
                {"bfs": lambda: g.bfs(src, out=False), "pr": lambda: g.pagerank(0.85, 10, out=False), "wcc": lambda: g.wcc(out=False),
                 "cdlp": lambda: g.cdlp(10, out=False), "lcc": lambda: g.lcc(out=False), "sssp": lambda: g.sssp(src, out=False)}[alg]()
                dt = time.perf_counter() - t0
                t = capi.last_timing()
                if best is None or dt < best[0]:
                    best = (dt, t["kernel_ms"], t["iterations"])
            if best:
                timing[alg] = {"wall_ms": round(best[0] * 1e3, 3), "kernel_ms": round(best[1], 3), "iterations": best[2]}
        rp_h = ci_h = w_h = None
        if not args.no_upload and not big:
            # the upload path on several GPUs: every rank pushes 1/world of the host arrays, the rest arrives by
            # all-gather over NVLink -- the device copy must equal the host arrays on every rank (u32 and u64 ids)
            rp_h, ci_h, w_h = g.download()
            for ids in (ci_h, ci_h.astype(np.uint64)):
                h = capi.Graph.from_csr(g.n, rp_h, ids, w_h, directed)
                rp2, ci2, w2 = h.download()
                same_w = (w2 is None and w_h is None) or np.array_equal(w2, w_h)
                if not (np.array_equal(rp2, rp_h) and np.array_equal(ci2, ci_h) and same_w):
                    failures.append((kind, rank, "upload", str(ids.dtype)))
                if "bfs" in res and not np.array_equal(h.bfs(src), res["bfs"]):
                    failures.append((kind, rank, "bfs after upload", str(ids.dtype)))
                h.free()
        # all ranks hold the same bits (PageRank included: a row is never split across ranks)
        digests = {alg: hashlib.sha256(np.ascontiguousarray(out).tobytes()).hexdigest() for alg, out in res.items()}
        all_d = [None] * world
        dist.all_gather_object(all_d, digests)
        for alg in algos:
            if any(d[alg] != all_d[0][alg] for d in all_d):
                failures.append((kind, rank, alg, "ranks differ"))
        if rank == 0:
            rp, ci, w = (rp_h, ci_h, w_h) if rp_h is not None else g.download()
            n = g.n
            T = oracle.transpose(n, rp, ci) if directed and any(a in algos for a in ("pr", "wcc", "cdlp", "lcc")) else None
            for alg in algos:
                t0 = time.perf_counter()
                out = res[alg]
                if alg == "bfs":
                    ok = np.array_equal(out, oracle.bfs(n, rp, ci, src))
                elif alg == "pr":
                    ref = oracle.pagerank(n, rp, ci, 0.85, 10, transposed=T)
                    ok = np.max(np.abs(out - ref) / ref) <= 1e-6
                elif alg == "wcc":
                    ok = np.array_equal(out, oracle.wcc(n, rp, ci, directed, transposed=T))
                elif alg == "cdlp":
                    ok = np.array_equal(out, oracle.cdlp(n, rp, ci, directed, 10, transposed=T))
                elif alg == "lcc":
                    if big:  # the oracle is quadratic in hub degrees: a sample incl. the top hubs
                        deg = np.diff(rp.astype(np.int64))
                        sample = np.unique(np.concatenate([np.argsort(deg)[-16:], np.random.default_rng(1).integers(0, n, 3000)]))
                        ref = oracle.lcc(n, rp, ci, directed, transposed=T, subset=sample.astype(np.uint64))
                        ok = np.allclose(out[sample], ref[sample], rtol=1e-9, atol=0)
                    else:
                        ok = np.allclose(out, oracle.lcc(n, rp, ci, directed, transposed=T), rtol=1e-9, atol=0)
                else:
                    ok = np.array_equal(out, oracle.sssp(n, rp, ci, w, src))
                if not ok:
                    failures.append((kind, 0, alg, "differs from the oracle"))
                report.append({"graph": f"RMAT-{scale} {kind}", "n": n, "nnz": int(g.nnz), "alg": alg, "ranks": world,
                               "match_oracle": bool(ok), "oracle_s": round(time.perf_counter() - t0, 2), **timing.get(alg, {})})
        g.free()
    all_fail = [None] * world
    dist.all_gather_object(all_fail, failures)
    failures = [f for fl in all_fail for f in fl]
    if rank == 0:
        for line in report:
            line["ranks_identical"] = not any(f[2] == line["alg"] and f[3] == "ranks differ" for f in failures if len(f) > 3)
            print(json.dumps(line), flush=True)
        print("MULTI_GPU_CHECK", "FAIL " + str(failures) if failures else f"OK world={world} scale={scale} algos={args.algos}",
              flush=True)
    dist.barrier()
    capi.comm_destroy()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
