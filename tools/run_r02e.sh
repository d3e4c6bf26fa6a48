#!/bin/bash
cd "$(dirname "$0")/.."
nvidia-smi -L | head -4
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
