#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python tools/multi_bench.py --out gpurun_out/multi_bench_r02.json 2>&1 | grep -v Warning | tail -9
