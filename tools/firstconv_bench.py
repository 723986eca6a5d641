"""A/B micro-benchmark of vpt_firstconv_pool: mma.sync kernel (mode 0) vs tcgen05 kernel (mode 1), plus agreement of the two."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat, ops
l = nat.lib()
g = torch.Generator().manual_seed(0)
for (F_, H, W, C0) in [(2048, 128, 128, 128), (2048, 128, 128, 64), (1024, 128, 128, 192), (1, 128, 128, 128)]:
    img = torch.randint(0, 256, (F_, H, W, 3), dtype=torch.uint8, generator=g).cuda()
    w = (torch.randn(C0, 27, generator=g) / 255.0 * 0.3).cuda()
    b = (torch.randn(C0, generator=g) * 0.1).cuda()
    res = {}
    for mode in (0, 1):
        l.vpt_set_firstconv_mode(mode)
        for _ in range(2):
            out, mr = ops.firstconv_pool(img, w, b, C0, zp=True)
        torch.cuda.synchronize()
        nat.device_check()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            out, mr = ops.firstconv_pool(img, w, b, C0, zp=True)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        nat.device_check()
        res[mode] = (out.float(), mr)
        gb = F_ * (H * W * 3 + (H // 2 + 1) * (W // 2 + 1) * C0 * 2) / 1e9
        print(f"F={F_} {H}x{W} C0={C0} mode={mode}: {ms:8.3f} ms (incl. stats finalize)  {gb/ms*1e3:7.0f} GB/s algorithmic", flush=True)
    d = (res[0][0] - res[1][0]).abs()
    print(f"   out diff: max {d.max().item():.3e}  mismatching {int((d > 0).sum())} / {d.numel()}  stats diff {(res[0][1]-res[1][1]).abs().max().item():.3e}")
l.vpt_set_firstconv_mode(1)
