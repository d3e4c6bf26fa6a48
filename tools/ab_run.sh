#!/bin/bash
# Developer tool (GPU box): parity + timing of the k_viterbi variants built by tools/build_variants.sh
cd "$(dirname "$0")/.."
R=$PWD/vibrato_b200
out=gpurun_out/ab.txt; : > $out
for v in "" _shfl16 _st2_16; do
  VBT_SO=$R/libvibrato_b200$v.so python -m pytest tests/test_gpu_parity.py -x -q -m gpu 2>&1 | tail -1 | sed "s/^/parity[$v] /" >> $out
done
for cfg in "shfl16 1" "st1_16 1" "st2_16 1" "shfl12 1" "st1_12 1" "st2_12 1" "shfl16 0" "st2_16 0"; do
  set -- $cfg
  VBT_SWEEP_TAG="$1/L$2" VBT_MATRIX_LAYOUT=$2 VBT_SO=$R/libvibrato_b200_$1.so python tools/sweep.py synth-unidic 1000000 2>&1 | grep lanes >> $out
done
cat $out
