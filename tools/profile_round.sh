#!/bin/bash
# Developer tool (GPU box): the measurement set behind profiles/rNN_* — GPU tests, both bench arms, the other BASELINE
# configurations, the ncu launch list of the bench command and full captures of k_viterbi2 and k_candidates.
#   tools/profile_round.sh <tag>
tag=${1:-r02}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.log; cut -c1-300 gpurun_out/bench_ref_$tag.json
python bench.py > gpurun_out/bench_ours_$tag.json 2> gpurun_out/bench_ours_$tag.log; cut -c1-300 gpurun_out/bench_ours_$tag.json
for c in 2 4 5; do
  python bench.py --config $c --steps 5 --warmup 3 > gpurun_out/bench_ours_${tag}_config$c.json 2> gpurun_out/bench_ours_${tag}_config$c.log; cut -c1-200 gpurun_out/bench_ours_${tag}_config$c.json
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}_config5.csv \
    python bench.py --config 5 --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_${tag}_config5.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_viterbi2 -s 2 -c 1 -f -o gpurun_out/prof_viterbi_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_candidates -s 1 -c 1 -f -o gpurun_out/prof_candidates_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_k2_$tag.log 2>&1
ls -la gpurun_out/*$tag*
