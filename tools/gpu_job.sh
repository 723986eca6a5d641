#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== idm tests"; timeout 1500 python -m pytest tests/test_idm.py tests/test_gpu_precise.py -x -q -m gpu 2>&1 | tail -3
echo "=== idm bench"; timeout 900 python tools/idm_bench.py 2>&1 | grep "IDM 4x\|instrumented\|conv3d_t5\|conv3x3_zp\|tensor-core\|maxpool"
