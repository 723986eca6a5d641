#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== training tests"; timeout 1500 python -m pytest tests/test_gpu_training.py -x -q -m gpu 2>&1 | tail -5
echo "=== bc bench"; timeout 900 python tools/bc_bench.py --ops 2>&1 | tail -30
