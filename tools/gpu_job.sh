#!/bin/bash
mkdir -p gpurun_out
echo "=== 1x by shape"; timeout 600 python bench.py --width 1x --batch 64 --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_1x.json 2> gpurun_out/bench_1x.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_1x.json"))
print("ms/step", d["ms_per_step"], "fps", d["value"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step_frac_of_flop_roofline"], "kernel ms", d["roofline"]["kernel_ms_per_step"])
for r in d["roofline"]["by_shape"][:8]: print(r)
PY
echo "=== 1x launch list"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 320 --csv --log-file gpurun_out/launches_1x.csv python bench.py --width 1x --batch 64 --steps 2 --warmup 1 --no-extras --no-cpu-baseline > /dev/null 2>&1; python tools/summarize_launches.py gpurun_out/launches_1x.csv | head -16
