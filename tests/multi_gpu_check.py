"""Multi-GPU parity check, launched as
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P tests/multi_gpu_check.py [scale]
Every rank runs the six algorithms on the 1-D partitioned path; the results must be
identical on all ranks and equal to the oracle (bit-exact for BFS/WCC/CDLP/SSSP, PR
within 1e-6, LCC within 1e-9)."""
import os
import sys

import numpy as np
import torch.distributed as dist

os.environ.setdefault("GX_TRANSPOSE_SPLIT", "1")  # exercise the column-split transposition at any world size

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import oracle  # noqa: E402
from ldbc_graphalytics_platforms_graphblas_b200 import capi  # noqa: E402
from ldbc_graphalytics_platforms_graphblas_b200.graphio import HostGraph  # noqa: E402


def main():
    scale = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    rank, world, local = (int(os.environ[k]) for k in ("RANK", "WORLD_SIZE", "LOCAL_RANK"))
    dist.init_process_group("gloo")
    capi.init(local)
    uid = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    capi.comm_init(rank, world, uid[0])
    failures = []
    for directed in (True, False):
        g = capi.Graph.rmat(scale, directed, weighted=True)
        src = g.max_degree_vertex()
        res = {"bfs": g.bfs(src), "pr": g.pagerank(0.85, 10), "wcc": g.wcc(), "cdlp": g.cdlp(10), "lcc": g.lcc(),
               "sssp": g.sssp(src)}
        # the upload path on several GPUs: every rank pushes 1/world of the host arrays, the rest arrives by
        # all-gather over NVLink -- the device copy must equal the host arrays on every rank (u32 and u64 ids)
        rp_h, ci_h, w_h = g.download()
        for ids in (ci_h, ci_h.astype(np.uint64)):
            h = capi.Graph.from_csr(g.n, rp_h, ids, w_h, directed)
            rp2, ci2, w2 = h.download()
            if not (np.array_equal(rp2, rp_h) and np.array_equal(ci2, ci_h) and np.array_equal(w2, w_h)):
                failures.append((directed, rank, "upload", str(ids.dtype)))
            if not np.array_equal(h.bfs(src), res["bfs"]):
                failures.append((directed, rank, "bfs after upload", str(ids.dtype)))
            h.free()
        gathered = [None] * world if rank == 0 else None
        dist.gather_object(res, gathered, dst=0)
        if rank == 0:
            rp, ci, w = g.download()
            n = g.n
            T = oracle.transpose(n, rp, ci) if directed else None
            ref = {"bfs": oracle.bfs(n, rp, ci, src), "pr": oracle.pagerank(n, rp, ci, 0.85, 10, transposed=T),
                   "wcc": oracle.wcc(n, rp, ci, directed, transposed=T),
                   "cdlp": oracle.cdlp(n, rp, ci, directed, 10, transposed=T),
                   "lcc": oracle.lcc(n, rp, ci, directed, transposed=T), "sssp": oracle.sssp(n, rp, ci, w, src)}
            for r, out in enumerate(gathered):
                for alg in ref:
                    if alg == "pr":
                        ok = np.max(np.abs(out[alg] - ref[alg]) / ref[alg]) <= 1e-6
                    elif alg == "lcc":
                        ok = np.allclose(out[alg], ref[alg], rtol=1e-9, atol=0)
                    else:
                        ok = np.array_equal(out[alg], ref[alg])
                    same = np.array_equal(out[alg], gathered[0][alg]) if alg != "pr" else np.allclose(out[alg], gathered[0][alg], rtol=1e-12)
                    if not (ok and same):
                        failures.append((directed, r, alg, bool(ok), bool(same)))
        g.free()
    all_fail = [None] * world
    dist.all_gather_object(all_fail, failures)
    failures = [f for fl in all_fail for f in fl]
    if rank == 0:
        print("MULTI_GPU_CHECK", "FAIL " + str(failures) if failures else f"OK world={world} scale={scale}", flush=True)
    dist.barrier()
    capi.comm_destroy()
    dist.destroy_process_group()
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
