#!/bin/bash
# Developer tool (GPU box): compute-sanitizer memcheck and racecheck over tests/probes/sanitize_probe.py
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool memcheck python tests/probes/sanitize_probe.py > gpurun_out/sanitizer_memcheck_r02.log 2>&1; tail -4 gpurun_out/sanitizer_memcheck_r02.log
timeout 500 compute-sanitizer --tool racecheck python tests/probes/sanitize_probe.py > gpurun_out/sanitizer_racecheck_r02.log 2>&1; tail -4 gpurun_out/sanitizer_racecheck_r02.log
