#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== rollout parity test"; timeout 900 python -m pytest tests/test_gpu_policy.py -x -q -m gpu -k "rollout_shape" -s 2>&1 | tail -12
