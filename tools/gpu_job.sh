#!/bin/bash
mkdir -p gpurun_out
echo "=== training tests"; timeout 1800 python -m pytest tests/test_gpu_training.py -x -q -s 2>&1 | grep -vE "^rel-L2|^cosine" | tail -12
echo "=== firstconv bwd at size"; timeout 300 python - <<'PY' 2>&1 | tail -6
import sys, torch
sys.path[:0] = [".", "tests", "oracle"]
import vpt_b200
from video_pre_training_b200 import ops, _native as nat
g = torch.Generator().manual_seed(0)
for C0, F_ in ((128, 2048), (192, 2048)):
    img = torch.randint(0, 256, (F_, 128, 128, 3), dtype=torch.uint8, generator=g).cuda()
    w = (torch.randn(C0, 27, generator=g) * 0.2 / 255.0).cuda(); b = (torch.randn(C0, generator=g) * 0.1).cuda()
    dy = torch.zeros(F_, 65, 65, C0, dtype=torch.bfloat16, device="cuda"); dy[:, :64, :64] = torch.randn(F_, 64, 64, C0, device="cuda").to(torch.bfloat16)
    res = {}
    for mode in (1, 0):
        nat.lib().vpt_set_firstconv_mode(mode)
        for _ in range(2): dW, db = ops.firstconv_bwd(img, w, b, dy, C0)
        torch.cuda.synchronize(); nat.device_check()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): dW, db = ops.firstconv_bwd(img, w, b, dy, C0)
        e1.record(); torch.cuda.synchronize()
        res[mode] = (dW, db)
        print(f"firstconv_bwd C0={C0} F={F_} mode={mode}: {e0.elapsed_time(e1)/3:.2f} ms")
    print("  tc vs cuda-core: rel dW", ((res[1][0]-res[0][0]).norm()/res[0][0].norm()).item(), "rel db", ((res[1][1]-res[0][1]).norm()/res[0][1].norm()).item())
nat.lib().vpt_set_firstconv_mode(1)
PY
echo "=== bc ops"; timeout 600 python tools/bc_bench.py --ops 2>&1 | tail -28
