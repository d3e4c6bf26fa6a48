#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02f_n1.json 2> gpurun_out/bench_r02f_n1.log; tail -3 gpurun_out/bench_r02f_n1.log; cat gpurun_out/bench_r02f_n1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench_r02f_n2.json 2> gpurun_out/bench_r02f_n2.log; tail -3 gpurun_out/bench_r02f_n2.log; cat gpurun_out/bench_r02f_n2.json
