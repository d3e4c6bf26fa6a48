#!/bin/bash
cd "$(dirname "$0")/.."
L=vibrato_b200/libvibrato_b200
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu_r02c.log 2>&1; tail -3 gpurun_out/pytest_gpu_r02c.log
C="viterbi_kernel=1,lanes_per_sentence=8/viterbi_kernel=2,lanes_per_sentence=8"
timeout 900 python tools/ab_multi.py --check \
  "base=$L.so:viterbi_kernel=0,lanes_per_sentence=8/$C" \
  "b2f0=${L}_b2f0.so:$C" "b2f1=${L}_b2f1.so:$C" "b4f0=${L}_b4f0.so:$C" "b4f1mb12=${L}_b4f1mb12.so:$C" \
  "b4f0mb12=${L}_b4f0mb12.so:$C" "b2f0mb12=${L}_b2f0mb12.so:$C" 2>&1 | tail -20
