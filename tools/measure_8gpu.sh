set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
$TR --master-port 29521 --nproc-per-node 8 tests/multi_gpu_check.py 16 > gpurun_out/r2s_mgc8_s16.log 2>&1; tail -1 gpurun_out/r2s_mgc8_s16.log
$TR --master-port 29522 --nproc-per-node 8 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/r2s_bench_n8.json 2> gpurun_out/r2s_bench_n8.err; cut -c1-250 gpurun_out/r2s_bench_n8.json
$TR --master-port 29523 --nproc-per-node 8 tools/multi_gpu_run.py --algos sssp,wcc,cdlp --scale 26 --undirected --hash > gpurun_out/r2s_multi8_rmat26.jsonl 2> gpurun_out/r2s_multi8_rmat26.err; cut -c1-330 gpurun_out/r2s_multi8_rmat26.jsonl
