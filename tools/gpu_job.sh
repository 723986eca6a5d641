#!/bin/bash
# Batch of GPU work for one gpurun call (edited between calls; the script travels with the repo snapshot).
mkdir -p gpurun_out
echo "=== full gpu tests"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
echo "=== side bounds"; timeout 600 python -m pytest tests/test_gpu_policy.py -q -s -k "small_config or fullsize or chunk_invariance" 2>&1 | grep -E "vpred|chunk invariance|passed|failed" | head -20
echo "=== bench.py (no extras), fold on"; timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_r2d.json"))
print("ms/step", d["ms_per_step"], "fps", d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "whole", d["roofline"]["whole_step_frac_of_flop_roofline"], "clk", d["clocks"]["sm_mhz"])
for r in d["roofline"]["by_shape"][:5]: print(r)
PY
tail -3 gpurun_out/bench_r2d.err
echo "=== fold off (A/B)"; timeout 600 python - <<'PY' 2>&1 | tail -5
import sys, torch
sys.path.insert(0, ".")
import vpt_b200
from video_pre_training_b200 import ops
torch.manual_seed(0)
pol = vpt_b200.MinecraftAgentPolicy(vpt_b200.minecraft_action_space(), vpt_b200.policy_kwargs("2x"), vpt_b200.PI_HEAD_KWARGS).cuda()
B, T = 128, 128
img = torch.randint(0, 256, (B, T, 128, 128, 3), dtype=torch.uint8, device="cuda")
first = torch.zeros(B, T, dtype=torch.bool, device="cuda")
for fold in (True, False, True, False):
    pol.net.fold_stack_norm = fold
    st = pol.initial_state(B)
    for _ in range(2):
        (_, _, _), st = pol({"img": img}, first, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(4):
        (_, _, _), st = pol({"img": img}, first, st)
    e1.record(); torch.cuda.synchronize()
    print(f"fold_stack_norm={fold}: {e0.elapsed_time(e1)/4:.1f} ms/step")
PY
