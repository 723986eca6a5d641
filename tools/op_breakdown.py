#!/usr/bin/env python
"""Per-op CUDA-event time breakdown of one policy forward step (diagnostic; not part of the bench contract).
    python tools/op_breakdown.py [--width 2x] [--batch 128] [--timesteps 128]"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import vpt_b200  # noqa: E402
from video_pre_training_b200 import ops  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--width", default="2x")
ap.add_argument("--batch", type=int, default=128)
ap.add_argument("--timesteps", type=int, default=128)
ap.add_argument("--cluster", type=int, default=0)
a = ap.parse_args()

dev = torch.device("cuda", 0)
if a.cluster:
    ops.set_default_cluster(a.cluster)
torch.manual_seed(0)
pol = vpt_b200.MinecraftAgentPolicy(vpt_b200.minecraft_action_space(), vpt_b200.policy_kwargs(a.width), vpt_b200.PI_HEAD_KWARGS).to(dev)
B, T = a.batch, a.timesteps
img = torch.randint(0, 256, (B, T, 128, 128, 3), dtype=torch.uint8, device=dev)
first = torch.zeros(B, T, dtype=torch.bool, device=dev)
st = pol.initial_state(B)
for _ in range(2):
    _, st = pol({"img": img}, first, st)
torch.cuda.synchronize()

records = []
names = ["gemm", "conv3x3_zp", "stats_finalize", "firstconv_pool", "maxpool3s2", "affine_norm", "affine_norm_zp", "copy_rows", "state_mask_update", "attention", "log_softmax"]
orig = {n: getattr(ops, n) for n in names}


def wrap(n):
    f = orig[n]

    def g(*args, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        r = f(*args, **kw)
        e1.record()
        tag = n
        if n == "gemm":
            M, N, K = args[3], args[4], args[5]
            tag = f"gemm {'conv' if kw.get('conv') is not None else 'lin'} N={N} K={K}" + (f" HW={kw['conv'][0]}" if kw.get("conv") is not None else "")
            records.append((tag, e0, e1, 2.0 * M * N * K))
        elif n == "conv3x3_zp":
            x, Wb, H, W = args[0], args[1], args[2], args[3]
            tag = f"conv_zp N={Wb.shape[0]} K={Wb.shape[1]} HW={H}" + (" +res" if kw.get("residual") is not None else "")
            records.append((tag, e0, e1, 2.0 * x.shape[0] * H * W * Wb.shape[0] * Wb.shape[1]))
        else:
            records.append((tag, e0, e1, 0.0))
        return r
    return g


# inner calls (stats_finalize inside firstconv_pool etc.) are attributed to the outer op as well as themselves -> wrap leaf ops only
for n in names:
    setattr(ops, n, wrap(n))
s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s0.record()
_, st = pol({"img": img}, first, st)
s1.record()
torch.cuda.synchronize()
tot = collections.defaultdict(float)
cnt = collections.Counter()
fl = collections.defaultdict(float)
for tag, e0, e1, f in records:
    tot[tag] += e0.elapsed_time(e1)
    cnt[tag] += 1
    fl[tag] += f
step = s0.elapsed_time(s1)
print(f"step {step:.1f} ms  ({B*T/step*1000:.0f} frames/s) ; note: firstconv_pool/maxpool3s2/affine_norm*/conv3x3_zp include their stats_finalize")
for k, v in sorted(tot.items(), key=lambda x: -x[1]):
    extra = f"  {fl[k]/v/1e9:8.0f} TFLOP/s" if fl[k] else ""
    print(f"{v:9.2f} ms {100*v/step:5.1f}%  n={cnt[k]:4d}  {k}{extra}")
