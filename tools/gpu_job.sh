#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== conv tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q -m gpu -k "conv3x3_zp or two_norm" 2>&1 | tail -15
echo "=== conv bench"; timeout 600 python tools/conv_bench.py 2>&1 | tail -20
