#!/bin/bash
mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "=== smoke"; timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8
echo "=== pool ncu"; timeout 300 ncu --metrics gpu__time_duration.sum,launch__registers_per_thread --clock-control none -k regex:"maxpool" -s 2 -c 2 --csv --log-file gpurun_out/pool.csv python bench.py --batch 16 --steps 1 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1; grep -E "gpu__time|registers" gpurun_out/pool.csv | cut -d, -f5,13,15 | head
echo "=== bench.py (all extras)"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2f.json"))
print("ms/step", d["ms_per_step"], "fps", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step_frac_of_flop_roofline"], "clk", d["clocks"], "launches", d["gpu_launches"])
for r in d["roofline"]["by_shape"][:5]: print(r)
for k in ("sample_agreement","gpu_eager_baseline","configs","bc"): print(k, json.dumps(d[k])[:700])
PY
tail -3 gpurun_out/bench_r2f.err
