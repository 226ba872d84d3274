set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --master-port 29501 --nproc-per-node 2 tests/multi_gpu_check.py 16 > gpurun_out/r2o_mgc2_s16.log 2>&1; tail -1 gpurun_out/r2o_mgc2_s16.log
$TR --master-port 29502 --nproc-per-node 4 tests/multi_gpu_check.py 16 > gpurun_out/r2o_mgc4_s16.log 2>&1; tail -1 gpurun_out/r2o_mgc4_s16.log
$TR --master-port 29503 --nproc-per-node 4 tests/multi_gpu_check.py 22 --algos pr,bfs,cdlp --no-upload > gpurun_out/r2o_mgc4_s22.log 2>&1; tail -4 gpurun_out/r2o_mgc4_s22.log | cut -c1-300
$TR --master-port 29504 --nproc-per-node 2 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2o_bench_n2.json 2> gpurun_out/r2o_bench_n2.err; cut -c1-250 gpurun_out/r2o_bench_n2.json
$TR --master-port 29505 --nproc-per-node 4 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2o_bench_n4.json 2> gpurun_out/r2o_bench_n4.err; cut -c1-250 gpurun_out/r2o_bench_n4.json
GX_PR_MAIL=0 $TR --master-port 29506 --nproc-per-node 4 bench.py --gpus 4 --steps 5 --warmup 3 > gpurun_out/r2o_bench_n4_nccl.json 2> gpurun_out/r2o_bench_n4_nccl.err; cut -c1-250 gpurun_out/r2o_bench_n4_nccl.json
$TR --master-port 29507 --nproc-per-node 4 tools/multi_gpu_run.py --algos cdlp,sssp --scale 24 --undirected --hash > gpurun_out/r2o_multi4_rmat24.jsonl 2>&1; tail -2 gpurun_out/r2o_multi4_rmat24.jsonl | cut -c1-400
