#!/bin/bash
# One gpurun call = one batch of GPU work (the script the builder edits between calls; it travels with the repo snapshot).
# Final-validation form:  /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash tools/gpu_job.sh'
mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== bench.py (all extras)"; timeout 1200 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 600 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
