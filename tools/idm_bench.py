#!/usr/bin/env python
"""IDM throughput (BASELINE configs[4]: 4x IDM, T=128): frames/s through InverseActionPolicy.predict on synthetic frames."""
import argparse, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import torch
import vpt_b200
import vpt_oracle as O
from video_pre_training_b200 import _native as nat

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=3)
a = ap.parse_args()
kw = vpt_b200.idm_net_kwargs()
torch.manual_seed(0)
t0 = time.time()
pol = vpt_b200.InverseActionPolicy(vpt_b200.idm_action_space(), dict(temperature=2.0), kw).cuda()
print(f"params {sum(p.numel() for p in pol.parameters())/1e6:.1f} M, init {time.time()-t0:.1f}s")
B, T = a.batch, 128
img = torch.randint(0, 256, (B, T, 128, 128, 3), dtype=torch.uint8, device="cuda")
first = torch.zeros(B, T, dtype=torch.bool, device="cuda")
for _ in range(2):
    ac, st, res = pol.predict({"img": img}, first=first, state_in=pol.initial_state(B))
torch.cuda.synchronize()
nat.device_check()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    ac, st, res = pol.predict({"img": img}, first=first, state_in=pol.initial_state(B))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.steps
cfg = O.Cfg(conv3d=True, **{k: v for k, v in kw.items() if k != "conv3d_params"})
print(f"IDM 4x B={B} T={T}: {ms:.1f} ms/step, {B*T/ms*1000:.0f} frames/s, peak mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB; "
      f"log_prob finite: {bool(torch.isfinite(res['log_prob']).all())}")
