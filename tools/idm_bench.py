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
# per-shape table of the tensor-core launches of one step (CUDA events around every GEMM / conv launch)
from video_pre_training_b200 import ops
import collections
ops.GEMM_PROFILE = []
ac, st, res = pol.predict({"img": img}, first=first, state_in=pol.initial_state(B))
torch.cuda.synchronize()
agg = collections.OrderedDict()
for (s0, s1, fl, tag, shape) in ops.GEMM_PROFILE:
    a_ = agg.setdefault((tag, tuple(shape)), [0, 0.0, 0.0])
    a_[0] += 1; a_[1] += s0.elapsed_time(s1); a_[2] += fl
ops.GEMM_PROFILE = None
tot = sum(v[1] for v in agg.values())
print(f"tensor-core launches: {tot:.1f} ms of the step")
for (tag, shape), (n, ms_, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {tag:6s} {str(shape):28s} n={n:3d} {ms_:8.2f} ms  {fl / ms_ / 1e9:7.0f} TFLOP/s")
# per-op breakdown (CUDA events around every ops.* call of one step; nested calls -- stats_finalize inside conv3x3_zp -- count in the outer op)
import types
recs, depth = [], [0]
orig = {n: f for n, f in vars(ops).items() if isinstance(f, types.FunctionType) and not n.startswith("_")}
def wrap(n, f):
    def g(*args, **kw):
        if depth[0]:
            return f(*args, **kw)
        depth[0] += 1
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        try:
            return f(*args, **kw)
        finally:
            e1.record(); depth[0] -= 1
            recs.append((n, e0, e1))
    return g
for n, f in orig.items():
    setattr(ops, n, wrap(n, f))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ac, st, res = pol.predict({"img": img}, first=first, state_in=pol.initial_state(B))
e1.record(); torch.cuda.synchronize()
for n, f in orig.items():
    setattr(ops, n, f)
agg = collections.OrderedDict()
for n, a0, a1 in recs:
    v = agg.setdefault(n, [0, 0.0]); v[0] += 1; v[1] += a0.elapsed_time(a1)
print(f"instrumented step {e0.elapsed_time(e1):.1f} ms; sum of ops {sum(v[1] for v in agg.values()):.1f} ms")
for n, (c, ms_) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
    print(f"  {ms_:8.2f} ms  n={c:4d}  {n}")
