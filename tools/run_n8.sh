#!/bin/bash
# Developer tool (8-GPU box): bench.py under torchrun at N=8 on a reduced batch (sanity of the multi-rank path)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
VBT_BENCH_BATCH=${1:-400000} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_n8_sanity.json 2> gpurun_out/bench_n8_sanity.log
tail -3 gpurun_out/bench_n8_sanity.log | cut -c1-300; cut -c1-600 gpurun_out/bench_n8_sanity.json
