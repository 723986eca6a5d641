"""Micro-benchmark: cost of the two-norm-composition epilogue modes (Ef table / affine residual) of vpt_conv3x3_zp vs the plain ones."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat, ops
g = torch.Generator().manual_seed(0)
for (HW, Cin, N, F_) in [(64, 128, 128, 2048), (32, 256, 256, 2048)]:
    x = torch.zeros(F_, HW + 1, HW + 1, Cin, dtype=torch.bfloat16, device="cuda")
    x[:, :HW, :HW] = torch.randn(F_, HW, HW, Cin, device="cuda").to(torch.bfloat16)
    xr = x.relu()  # pooled-like input (non-negative, half zeros)
    r = torch.zeros(F_, HW + 1, HW + 1, N, dtype=torch.bfloat16, device="cuda")
    r[:, :HW, :HW] = torch.randn(F_, HW, HW, N, device="cuda").to(torch.bfloat16)
    Wb = (torch.randn(N, 9 * Cin, device="cuda") * (9 * Cin) ** -0.5).to(torch.bfloat16)
    mr = torch.stack([torch.randn(F_) * 0.1, torch.rand(F_) + 0.5], 1).cuda()
    mrE = torch.stack([torch.zeros(F_), torch.rand(F_) + 0.5], 1).cuda()
    S1 = torch.randn(9, N, device="cuda"); S2 = torch.randn(9, N, device="cuda")
    Ef = torch.randn(F_, 9, N, device="cuda")
    rs, rb = torch.randn(F_, N, device="cuda"), torch.randn(F_, N, device="cuda")
    fl = 2.0 * F_ * HW * HW * N * 9 * Cin
    cases = [("plain          ", dict(x=x, mr=mr, S1=S1, S2=S2)), ("plain relu-in  ", dict(x=xr, mr=mr, S1=S1, S2=S2)),
             ("Ef table       ", dict(x=xr, mr=mrE, Ef=Ef)),
             ("residual       ", dict(x=x, mr=mr, S1=S1, S2=S2, residual=r)), ("affine residual", dict(x=x, mr=mr, S1=S1, S2=S2, residual=r, res_scale=rs, res_shift=rb))]
    for name, kw in cases:
        xx = kw.pop("x")
        for _ in range(2):
            ops.conv3x3_zp(xx, Wb, HW, HW, relu=1, want_stats=True, **kw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(8):
            ops.conv3x3_zp(xx, Wb, HW, HW, relu=1, want_stats=True, **kw)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 8
        nat.device_check()
        print(f"HW={HW} {Cin}->{N} F={F_}: {name}: {ms:7.3f} ms  {fl/ms/1e9:7.0f} TFLOP/s", flush=True)
