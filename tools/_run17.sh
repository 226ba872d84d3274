set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
GX_PR_COMPACT=1 $TR --master-port 29511 --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench_n2_compact.json 2> gpurun_out/r2r_bench_n2_compact.err; cut -c1-250 gpurun_out/r2r_bench_n2_compact.json
GX_PR_FUSED=0 $TR --master-port 29512 --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2r_bench_n2_push.json 2> gpurun_out/r2r_bench_n2_push.err; cut -c1-250 gpurun_out/r2r_bench_n2_push.json
