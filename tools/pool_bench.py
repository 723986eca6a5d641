"""maxpool3s2 (stacks 1-2 of the 2x model, 2048 frames): ms and achieved HBM bandwidth (algorithmic bytes = read the input once + write the output)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat, ops
F_ = int(os.environ.get("F", "2048"))
for (HW, C) in ((64, 256), (32, 256)):
    x = torch.zeros(F_, HW + 1, HW + 1, C, dtype=torch.bfloat16, device="cuda")
    x[:, :HW, :HW] = torch.rand(F_, HW, HW, C, device="cuda").to(torch.bfloat16)
    for chan in (False, True):
        for _ in range(2):
            out = ops.maxpool3s2(x, zp=True, want_chan=chan)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            out = ops.maxpool3s2(x, zp=True, want_chan=chan)
        e1.record(); torch.cuda.synchronize()
        nat.device_check()
        ms = e0.elapsed_time(e1) / 10
        gb = (x.numel() + out[0].numel()) * 2 / 1e9
        print(f"pool {HW}x{HW}x{C} F={F_} chan={int(chan)}: {ms:.3f} ms (incl. stats finalize)  {gb/ms:.2f} TB/s", flush=True)
