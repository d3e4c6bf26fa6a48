#!/bin/bash
cd "$(dirname "$0")/.."
L=vibrato_b200/libvibrato_b200
timeout 900 python tools/ab_multi.py --check \
  "base=$L.so:viterbi_kernel=1/viterbi_kernel=1,sort_by_length=2/viterbi_kernel=1,sort_by_length=2,lanes_per_sentence=16/viterbi_kernel=1,sort_by_length=1" \
  "ep=${L}_ep.so:viterbi_kernel=1/viterbi_kernel=1,sort_by_length=2" 2>&1 | tail -20
bash tools/run_prof.sh v2d k_viterbi2
