#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== bench.py (no extras)"; timeout 900 python bench.py --no-extras --no-cpu-baseline > gpurun_out/bench_r2i.json 2> gpurun_out/bench_r2i.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2i.json"))
print("ms/step", d["ms_per_step"], "fps", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step_frac_of_flop_roofline"], "clk", d["clocks"], "launches", d["gpu_launches"])
for s in d["roofline"]["by_shape"][:5]: print(s)
PY
tail -3 gpurun_out/bench_r2i.err
