#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== kernel+policy tests"; timeout 1800 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_policy.py -x -q -m gpu 2>&1 | tail -5
echo "=== rollout sections"; timeout 600 python tools/rollout_sections.py 2>&1 | tail -10
