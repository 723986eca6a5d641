#!/usr/bin/env python
"""Rollout-latency path (SURVEY f-1): per-step latency of MinecraftAgentPolicy.act at B small, T=1 (agent.py:190-206),
eager launches vs one captured CUDA graph per step."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat

ap = argparse.ArgumentParser()
ap.add_argument("--width", default="2x")
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--steps", type=int, default=200)
a = ap.parse_args()
torch.manual_seed(0)
pol = vpt_b200.MinecraftAgentPolicy(vpt_b200.minecraft_action_space(), vpt_b200.policy_kwargs(a.width), vpt_b200.PI_HEAD_KWARGS).cuda()
B = a.batch
img = torch.randint(0, 256, (B, 128, 128, 3), dtype=torch.uint8, device="cuda")
first = torch.zeros(B, dtype=torch.bool, device="cuda")
st = pol.initial_state(B)
for _ in range(5):
    ac, st, res = pol.act({"img": img}, first, st)
torch.cuda.synchronize(); nat.device_check()
t0 = time.perf_counter()
for _ in range(a.steps):
    ac, st, res = pol.act({"img": img}, first, st)
    _ = ac["buttons"].cpu()  # the env needs the action on the host every step (agent.py:151-164)
t1 = time.perf_counter()
print(f"eager : {a.width} B={B} T=1: {(t1-t0)/a.steps*1e3:.3f} ms/step  ({B*a.steps/(t1-t0):.0f} frames/s)")
for pdl in (False, True):
    step = pol.make_graphed_act(B, pdl=pdl)
    st = pol.initial_state(B)
    for _ in range(5):
        ac, st, res = step({"img": img}, first, st)
    torch.cuda.synchronize(); nat.device_check()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        ac, st, res = step({"img": img}, first, st)
        _ = ac["buttons"].cpu()
    t1 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        ac, st, res = step({"img": img}, first, st)
    e1.record(); torch.cuda.synchronize()
    print(f"graph pdl={int(pdl)}: device {e0.elapsed_time(e1)/a.steps:.3f} ms/step")
    print(f"graph pdl={int(pdl)}: {a.width} B={B} T=1: {(t1-t0)/a.steps*1e3:.3f} ms/step  ({B*a.steps/(t1-t0):.0f} frames/s)")
    # parity: graphed step == eager step on the same inputs (deterministic action)
    st_a, st_b = pol.initial_state(B), pol.initial_state(B)
    for _ in range(3):
        ac_a, st_a, res_a = pol.act({"img": img}, first, st_a, stochastic=False, return_pd=True)
        ac_b, st_b, res_b = step({"img": img}, first, st_b, stochastic=False, return_pd=True)
    print("graph == eager:", bool(torch.equal(res_a["pd"]["buttons"], res_b["pd"]["buttons"])), bool(torch.equal(ac_a["camera"], ac_b["camera"])))
