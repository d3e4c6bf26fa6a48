#!/bin/bash
cd "$(dirname "$0")/.."
L=vibrato_b200/libvibrato_b200
C="viterbi_kernel=1,lanes_per_sentence=8/viterbi_kernel=2,lanes_per_sentence=8"
timeout 900 python tools/ab_multi.py --check \
  "base=$L.so:viterbi_kernel=0,lanes_per_sentence=8/$C/viterbi_kernel=1,lanes_per_sentence=16/viterbi_kernel=1,lanes_per_sentence=4" \
  "b4f0=${L}_b4f0.so:$C" "b4f1=${L}_b4f1.so:$C" "b2f1=${L}_b2f1.so:$C" \
  "b4f0mb12=${L}_b4f0mb12.so:$C" "pf0=${L}_pf0.so:viterbi_kernel=1" "pf16=${L}_pf16.so:viterbi_kernel=1" 2>&1 | tail -20
