#!/bin/bash
# Batch of GPU work for one gpurun call (edited between calls; the script travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== kernel tests"; timeout 600 python -m pytest tests/test_gpu_kernels.py -x -q 2>&1 | tail -8
echo "=== firstconv bench"; timeout 200 python tools/firstconv_bench.py 2>&1 | tail -20
echo "=== conv bench"; timeout 200 python tools/conv_bench.py 2>&1 | tail -20
echo "=== full gpu tests"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== bench.py (no extras)"; timeout 600 python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2b.json"))
print("ms/step", d["ms_per_step"], "fps", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step_frac_of_flop_roofline"])
for r in d["roofline"]["by_shape"][:5]: print(r)
PY
tail -3 gpurun_out/bench_r2b.err
