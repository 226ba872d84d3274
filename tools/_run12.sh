set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2g_gputests.log 2>&1; tail -2 gpurun_out/r2g_gputests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/r2g_bench_n1.json 2> gpurun_out/r2g_bench_n1.err; cut -c1-300 gpurun_out/r2g_bench_n1.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2g_reference_arm.json 2> gpurun_out/r2g_reference_arm.err; cut -c1-300 gpurun_out/r2g_reference_arm.json
# launch list of the bench command (plain run first, then the same command under ncu)
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_plain_a.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2g_launches_bench_rmat22.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_ncu_a.log 2>&1
# per-kernel DRAM bytes of all six algorithms
python tools/bench_algos.py --algos bfs,pr --scale 22 --reps 1 --cache-at > gpurun_out/r2g_plain_b.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2g_dram_d22.csv python tools/bench_algos.py --algos bfs,pr --scale 22 --reps 1 --cache-at > gpurun_out/r2g_ncu_b.log 2>&1
python tools/bench_algos.py --algos wcc,cdlp,lcc,sssp --scale 22 --reps 1 --undirected > gpurun_out/r2g_plain_c.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r2g_dram_u22.csv python tools/bench_algos.py --algos wcc,cdlp,lcc,sssp --scale 22 --reps 1 --undirected > gpurun_out/r2g_ncu_c.log 2>&1
# full captures of the top kernels
python tools/one_shot.py --algos pr --scale 22 --twice > gpurun_out/r2g_plain_d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_pr_tile -s 14 -c 4 -o gpurun_out/r2g_prof_pr python tools/one_shot.py --algos pr --scale 22 --twice > gpurun_out/r2g_ncu_d.log 2>&1
python tools/one_shot.py --algos lcc --scale 22 --undirected --twice > gpurun_out/r2g_plain_e.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_lcc_count -s 1 -c 1 -o gpurun_out/r2g_prof_lcc python tools/one_shot.py --algos lcc --scale 22 --undirected --twice > gpurun_out/r2g_ncu_e.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2g_*.csv
