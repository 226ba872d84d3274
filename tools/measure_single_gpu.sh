set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2p_gputests.log 2>&1; tail -2 gpurun_out/r2p_gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2p_bench_n1.json 2> gpurun_out/r2p_bench_n1.err; cut -c1-300 gpurun_out/r2p_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2p_reference_arm.json 2> gpurun_out/r2p_reference_arm.err; cut -c1-300 gpurun_out/r2p_reference_arm.json
python tools/cold_tproc.py --scale 22 > gpurun_out/r2p_cold_tproc_d22.jsonl 2> gpurun_out/r2p_cold_d22.err; cut -c1-260 gpurun_out/r2p_cold_tproc_d22.jsonl
python tools/cold_tproc.py --scale 22 --undirected > gpurun_out/r2p_cold_tproc_u22.jsonl 2> gpurun_out/r2p_cold_u22.err; cut -c1-260 gpurun_out/r2p_cold_tproc_u22.jsonl
python tools/cdlp_vs_reference.py --scale 22 --check > gpurun_out/r2p_cdlp_vs_ref_cuda_rmat22.json 2> gpurun_out/r2p_cdlp_ref22.err; tail -1 gpurun_out/r2p_cdlp_vs_ref_cuda_rmat22.json | cut -c1-600
python tools/cdlp_vs_reference.py --scale 24 --check > gpurun_out/r2p_cdlp_vs_ref_cuda_rmat24.json 2> gpurun_out/r2p_cdlp_ref24.err; tail -1 gpurun_out/r2p_cdlp_vs_ref_cuda_rmat24.json | cut -c1-600
# per-kernel DRAM bytes of the undirected four (CDLP kernels changed)
python tools/bench_algos.py --algos wcc,cdlp,lcc,sssp --scale 22 --reps 1 --undirected > gpurun_out/r2p_plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2p_dram_u22.csv python tools/bench_algos.py --algos wcc,cdlp,lcc,sssp --scale 22 --reps 1 --undirected > gpurun_out/r2p_ncu_c.log 2>&1
python tools/bench_algos.py --algos cdlp --scale 24 --reps 1 --undirected > gpurun_out/r2p_plain_d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_cdlp --csv --log-file gpurun_out/r2p_dram_cdlp_u24.csv python tools/bench_algos.py --algos cdlp --scale 24 --reps 1 --undirected > gpurun_out/r2p_ncu_d.log 2>&1
# all six on the config graphs, checked
python tools/bench_algos.py --algos bfs,pr,wcc,cdlp,sssp --scale 22 --reps 3 --check --cache-at > gpurun_out/r2p_algos_d22.jsonl 2>&1
python tools/bench_algos.py --algos bfs,wcc,cdlp,sssp --scale 24 --reps 3 --undirected --check > gpurun_out/r2p_algos_u24.jsonl 2>&1; tail -4 gpurun_out/r2p_algos_u24.jsonl | cut -c1-200
python tools/bench_algos.py --algos lcc --scale 22 --reps 3 --undirected > gpurun_out/r2p_lcc_u22.jsonl 2>&1
# RMAT-26 on one GPU: SSSP against Dijkstra + hashes to compare with the 8-GPU runs
python tools/bench_algos.py --algos sssp,wcc,cdlp --scale 26 --reps 1 --undirected --check --hash > gpurun_out/r2p_algos_u26_checked.jsonl 2>&1; tail -3 gpurun_out/r2p_algos_u26_checked.jsonl | cut -c1-400
python tools/cold_tproc.py --scale 24 --undirected --algos wcc,cdlp,sssp > gpurun_out/r2p_cold_tproc_u24.jsonl 2> gpurun_out/r2p_cold_u24.err; cut -c1-260 gpurun_out/r2p_cold_tproc_u24.jsonl
