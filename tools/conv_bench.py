"""Micro-benchmark of vpt_conv3x3_zp on the stack-0 layer shape (128 -> 128 @ 64x64): A/B of the Cout == 128 kernel's epilogue variants.
  swap mode 1 = two-phase transposing epilogue on 16 warps (default), 6 = the same on 8 warps, 2 = no epilogue (MMA-rate experiment), 4 = fragment epilogue, 5 = channel-major
  single-pass epilogue (v3); bits 8..15 cap the weight pipeline depth, bits 20..23 switch parts of v3 off (timing experiments only).
  pair mode bit 8 (0x100) = round-1 per-thread global-store epilogue instead of the TMA-store one (pair kernel)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat, ops
l = nat.lib()
g = torch.Generator().manual_seed(0)
shapes = [(64, 128, 128, 2048, False), (64, 128, 128, 2048, True)]
variants = [("2phase 16w", 1, 1), ("2phase 8w", 1, 6), ("mma-only", 1, 2), ("frag", 1, 4), ("v3", 1, 5)]
for (HW, Cin, N, F_, res) in shapes:
    x = torch.zeros(F_, HW + 1, HW + 1, Cin, dtype=torch.bfloat16, device="cuda")
    x[:, :HW, :HW] = torch.randn(F_, HW, HW, Cin, device="cuda").to(torch.bfloat16)
    r = None
    if res:
        r = torch.zeros(F_, HW + 1, HW + 1, N, dtype=torch.bfloat16, device="cuda")
        r[:, :HW, :HW] = torch.randn(F_, HW, HW, N, device="cuda").to(torch.bfloat16)
    Wb = (torch.randn(N, 9 * Cin, device="cuda") * (9 * Cin) ** -0.5).to(torch.bfloat16)
    mr = torch.stack([torch.randn(F_) * 0.1, torch.rand(F_) + 0.5], 1).cuda()
    S1 = torch.randn(9, N, device="cuda"); S2 = torch.randn(9, N, device="cuda")
    fl = 2.0 * F_ * HW * HW * N * 9 * Cin
    ref = None
    for (name, mode, swap) in variants:
        l.vpt_set_conv_pair_mode(mode)
        l.vpt_set_conv_swap_mode(swap)
        for _ in range(2):
            out, st = ops.conv3x3_zp(x, Wb, HW, HW, mr=mr, S1=S1, S2=S2, relu=1, residual=r, want_stats=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out, _ = ops.conv3x3_zp(x, Wb, HW, HW, mr=mr, S1=S1, S2=S2, relu=1, residual=r, want_stats=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nat.device_check()
        if ref is None: ref = (out.float(), st)
        err = (out.float() - ref[0]).abs().max().item()
        serr = (st - ref[1]).abs().max().item()
        print(f"HW={HW} Cin={Cin} N={N} F={F_} res={int(res)}: {name:13s}: {ms:7.3f} ms  {fl/ms/1e9:7.0f} TFLOP/s (algorithmic)  max diff vs first {err:.1e} stats {serr:.1e}", flush=True)
l.vpt_set_conv_pair_mode(1); l.vpt_set_conv_swap_mode(1)
