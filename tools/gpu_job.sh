#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== rollout"; timeout 600 python tools/rollout_bench.py --steps 300 2>&1 | tail -8
