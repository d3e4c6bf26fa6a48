#!/bin/bash
# Developer tool (GPU box): the measurement set behind profiles/rNNx_* — GPU tests, both bench arms, the ncu launch
# list of the bench command and one full capture of k_viterbi.   tools/profile_round.sh <tag>
tag=${1:-r01x}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -2 gpurun_out/pytest_gpu_$tag.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2> gpurun_out/bench_ref_$tag.log; cat gpurun_out/bench_ref_$tag.json
python bench.py > gpurun_out/bench_ours_$tag.json 2> gpurun_out/bench_ours_$tag.log; cat gpurun_out/bench_ours_$tag.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_viterbi -s 2 -c 1 -f -o gpurun_out/prof_viterbi_$tag \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_$tag.log 2>&1
ls -la gpurun_out/*$tag*
