#!/bin/bash
# Developer tool (GPU box): one ncu --set full capture of a kernel of the default library.
#   tools/run_prof.sh <tag> <kernel regex> [ab_multi config]
cd "$(dirname "$0")/.."
tag=$1; rx=$2; cfg=${3:-viterbi_kernel=1}
L=vibrato_b200/libvibrato_b200
timeout 800 ncu --set full --clock-control none --import-source on -k regex:$rx -s 2 -c 1 -f -o gpurun_out/prof_$tag \
  python tools/ab_multi.py --reps 1 "base=$L.so:$cfg" > gpurun_out/ncu_$tag.log 2>&1
tail -2 gpurun_out/ncu_$tag.log
