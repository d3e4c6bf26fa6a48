#!/bin/bash
# Developer tool (GPU box): GPU tests + stage times of the default build.
cd "$(dirname "$0")/.."
L=vibrato_b200/libvibrato_b200
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python tools/ab_multi.py "base=$L.so:viterbi_kernel=1" 2>&1 | tail -3
