#!/bin/bash
# One gpurun call = one batch of GPU work (this is the script the builder edits between calls; it travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
echo "=== bench.py (all extras)"; timeout 1200 python bench.py > gpurun_out/bench_r2k.json 2> gpurun_out/bench_r2k.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2k.json"))
print("ms/step", d["ms_per_step"], "fps", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step_frac_of_flop_roofline"], "clk", d["clocks"], "launches", d["gpu_launches"])
for k in ("configs","bc"): print(k, json.dumps(d.get(k))[:800])
PY
tail -3 gpurun_out/bench_r2k.err
