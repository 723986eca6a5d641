"""Micro-benchmark of vpt_conv3x3_zp over the 2x-width layer shapes under the MMA-issue debug variants."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat, ops
l = nat.lib()
g = torch.Generator().manual_seed(0)
shapes = [(64, 128, 128, 2048)]
variants = [(0, 0), (0, 1), (0, 2)]  # (pair mode, swap mode)  # (pair mode, swap mode)
for (HW, Cin, N, F_) in shapes:
    x = torch.zeros(F_, HW + 1, HW + 1, Cin, dtype=torch.bfloat16, device="cuda")
    x[:, :HW, :HW] = torch.randn(F_, HW, HW, Cin, device="cuda").to(torch.bfloat16)
    Wb = (torch.randn(N, 9 * Cin, device="cuda") * (9 * Cin) ** -0.5).to(torch.bfloat16)
    mr = torch.stack([torch.zeros(F_), torch.ones(F_)], 1).cuda()
    S1 = torch.zeros(9, N, device="cuda"); S2 = torch.zeros(9, N, device="cuda")
    fl = 2.0 * F_ * HW * HW * N * 9 * Cin
    ref = None
    for (nsplit, korder) in variants:
        l.vpt_set_conv_pair_mode(nsplit); l.vpt_set_conv_swap_mode(korder)
        for _ in range(2):
            out, _ = ops.conv3x3_zp(x, Wb, HW, HW, mr=mr, S1=S1, S2=S2, relu=1, want_stats=False)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out, _ = ops.conv3x3_zp(x, Wb, HW, HW, mr=mr, S1=S1, S2=S2, relu=1, want_stats=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nat.device_check()
        if ref is None: ref = out.float()
        err = ((out.float() - ref).norm() / ref.norm()).item()
        print(f"HW={HW} Cin={Cin} N={N} F={F_}: pair={nsplit} swap={korder}: {ms:7.3f} ms  {fl/ms/1e9:7.0f} TFLOP/s (algorithmic)  diff vs first {err:.1e}")
l.vpt_set_conv_pair_mode(1); l.vpt_set_conv_swap_mode(1)
