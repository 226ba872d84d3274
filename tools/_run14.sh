set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "cdlp or hub_rows or golden or rmat" > gpurun_out/r2n_gputests_cdlp.log 2>&1; tail -2 gpurun_out/r2n_gputests_cdlp.log
for s in "22 --undirected" "22" "24 --undirected"; do
  tag=$(echo $s | tr -d ' -')
  python tools/bench_algos.py --algos cdlp --scale $s --reps 3 --check > gpurun_out/r2n_cdlp_new_$tag.jsonl 2>&1
  tail -1 gpurun_out/r2n_cdlp_new_$tag.jsonl | cut -c1-700
done
python tools/iter_profile.py --scale 24 --undirected --cdlp 10 > gpurun_out/r2n_iter_u24.jsonl 2>&1; tail -12 gpurun_out/r2n_iter_u24.jsonl | cut -c1-400
