#!/bin/bash
cd "$(dirname "$0")/.."
L=vibrato_b200/libvibrato_b200
timeout 400 python tools/ab_multi.py --check "base=$L.so:viterbi_kernel=1/viterbi_kernel=2" "bulk=${L}_bulk.so:viterbi_kernel=1/viterbi_kernel=2" "base2=$L.so:viterbi_kernel=1" "bulk2=${L}_bulk.so:viterbi_kernel=1" 2>&1 | tail -7
