#!/bin/bash
cd "$(dirname "$0")/.."
L=vibrato_b200/libvibrato_b200
timeout 900 python tools/ab_multi.py --check "base=$L.so:viterbi_kernel=1" "b4=${L}_b4.so:viterbi_kernel=1" "mb12=${L}_mb12.so:viterbi_kernel=1" "b4mb12=${L}_b4mb12.so:viterbi_kernel=1" "f1=${L}_f1.so:viterbi_kernel=1" "pf0=${L}_pf0.so:viterbi_kernel=1" "pf16=${L}_pf16.so:viterbi_kernel=1" 2>&1 | tail -8
