#!/bin/bash
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 900 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -4
echo "=== pool ncu"; timeout 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:"maxpool" -s 2 -c 2 --csv --log-file gpurun_out/pool.csv python bench.py --batch 16 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1; grep -E "gpu__time" gpurun_out/pool.csv | awk -F'","' '{print $5, $NF}' | head
echo "=== bench"; timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2g.json"))
print("ms/step", d["ms_per_step"], "fps", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step_frac_of_flop_roofline"], "clk", d["clocks"]["sm_mhz"])
PY
