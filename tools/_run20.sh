set -x
P=ldbc_graphalytics_platforms_graphblas_b200/lib/libgxb200_prev.so
python -m pytest tests/test_gpu_parity.py -x -q -k "lcc or cdlp or hub_rows or golden or rmat" > gpurun_out/r2t_gputests.log 2>&1; tail -2 gpurun_out/r2t_gputests.log
GX_LIB=$P python tools/bench_algos.py --algos lcc --scale 22 --reps 3 --undirected > gpurun_out/r2t_lcc_prev_u22.jsonl 2>&1; tail -1 gpurun_out/r2t_lcc_prev_u22.jsonl | cut -c1-300
python tools/bench_algos.py --algos lcc --scale 22 --reps 3 --undirected > gpurun_out/r2t_lcc_new_u22.jsonl 2>&1; tail -1 gpurun_out/r2t_lcc_new_u22.jsonl | cut -c1-300
python tools/bench_algos.py --algos lcc --scale 22 --reps 3 > gpurun_out/r2t_lcc_new_d22.jsonl 2>&1; tail -1 gpurun_out/r2t_lcc_new_d22.jsonl | cut -c1-300
python tools/bench_algos.py --algos cdlp --scale 24 --reps 3 --undirected --check > gpurun_out/r2t_cdlp_u24.jsonl 2>&1; tail -1 gpurun_out/r2t_cdlp_u24.jsonl | cut -c1-600
python tools/bench_algos.py --algos cdlp --scale 22 --reps 3 --check > gpurun_out/r2t_cdlp_d22.jsonl 2>&1; tail -1 gpurun_out/r2t_cdlp_d22.jsonl | cut -c1-300
python tools/iter_profile.py --scale 24 --undirected --cdlp 10 > gpurun_out/r2t_iter_u24.jsonl 2>&1
