#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== ncu launch list (one bench step, B=128, no extras)"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 340 -c 420 --csv --log-file gpurun_out/launches_r2b.csv python bench.py --steps 1 --warmup 1 --no-extras --no-cpu-baseline > gpurun_out/ncu_launch_b.log 2>&1
tail -2 gpurun_out/ncu_launch_b.log | cut -c1-300; wc -l gpurun_out/launches_r2b.csv
