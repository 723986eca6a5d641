"""A/B of the weight-gradient kernels on the 3x-width BC shapes (2048 frames): vpt_set_wgrad_mode(0) = one GEMM tile per tap,
1 = tap-pairing kernel (csrc/wgrad_tc.cuh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat, ops
l = nat.lib()
F_ = int(os.environ.get("F", "2048"))
shapes = [("stack0 192->192 @64", 192, 192, 64), ("stack1 first 192->384 @64", 384, 192, 64), ("stack1 384->384 @32", 384, 384, 32),
          ("stack2 384->384 @16", 384, 384, 16), ("2x stack0 128->128 @64", 128, 128, 64), ("2x stack1 256->256 @32", 256, 256, 32)]
for name, M, N, HW in shapes:
    Wp = HW + 1
    R = F_ * Wp * Wp
    a = torch.randn(R, M, device="cuda").to(torch.bfloat16)
    b = torch.randn(R, N, device="cuda").to(torch.bfloat16)
    shifts = [(ky - 1) * Wp + (kx - 1) for ky in range(3) for kx in range(3)]
    fl = 2.0 * M * N * 9 * R
    outs = {}
    for rep in range(2):
        for mode in (0, 1):
            l.vpt_set_wgrad_mode(mode)
            for _ in range(2):
                out = ops.wgrad(a, b, shifts)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                out = ops.wgrad(a, b, shifts)
            e1.record(); torch.cuda.synchronize()
            nat.device_check()
            ms = e0.elapsed_time(e1) / 5
            outs[mode] = out
            print(f"{name:28s} mode {mode}: {ms:7.3f} ms  {fl/ms/1e9:7.0f} TFLOP/s", flush=True)
    d = (outs[0] - outs[1]).abs().max().item() / outs[0].abs().max().item()
    print(f"{name:28s} max rel diff between the kernels {d:.2e}", flush=True)
    del a, b
l.vpt_set_wgrad_mode(1)
