"""Where the B=1 rollout step goes: sections of the forward captured as separate CUDA graphs and replayed (2x width, one frame)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import vpt_b200
from video_pre_training_b200 import _native as nat, ops

torch.manual_seed(0)
pol = vpt_b200.MinecraftAgentPolicy(vpt_b200.minecraft_action_space(), vpt_b200.policy_kwargs("2x"), vpt_b200.PI_HEAD_KWARGS).cuda()
net, cfg = pol.net, pol.net.cfg
prep = net.prepared(); pol._heads_prepared()
B = 1
img = torch.randint(0, 256, (B, 128, 128, 3), dtype=torch.uint8, device="cuda")
first = torch.zeros(B, dtype=torch.bool, device="cuda")
st = pol.initial_state(B)
Hf, Wf = cfg.final_hw; C2 = cfg.chans[-1]
cnn_out = torch.empty((B, Hf + 1, Wf + 1, C2), dtype=torch.bfloat16, device="cuda")
first_u8 = first.view(B, 1).contiguous().view(torch.uint8)

def time_graph(name, fn, n=300):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(3): fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g), torch.no_grad():
        out = fn()
    for _ in range(5): g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): g.replay()
    e1.record(); torch.cuda.synchronize()
    nat.device_check()
    print(f"{name:34s} {e0.elapsed_time(e1) / n * 1000:8.1f} us", flush=True)
    return out

with torch.no_grad():
    _, mr_c = net._cnn_chunk(img, prep, cnn_out)
    Kd = (Hf + 1) * (Wf + 1) * C2
    xd, mr_d = net._linear(cnn_out.view(B, Kd), prep.dense, cfg.cnn_outsize, mr=mr_c, relu=1, want_stats=True)
    x, mr_x = net._linear(xd, prep.linear, cfg.hidsize, mr=mr_d, relu=1, want_stats=True)
    x1, mr_x1, s1 = net._block(0, x, mr_x, first_u8, st[0], B, 1, prep, last=False)
    lat, _, _ = ops.affine_norm(x1, mr_x1, prep.fin_g, prep.fin_b, rows_per_group=1, want_f32=True)

time_graph("empty graph (one tiny kernel)", lambda: ops.affine_norm(x1, mr_x1, prep.fin_g, prep.fin_b, rows_per_group=1))
time_graph("CNN (firstconv + 3 stacks)", lambda: net._cnn_chunk(img, prep, cnn_out))
time_graph("dense + linear (2 GEMV + stats)", lambda: net._linear(net._linear(cnn_out.view(B, Kd), prep.dense, cfg.cnn_outsize, mr=mr_c, relu=1, want_stats=True)[0], prep.linear, cfg.hidsize, mr=mr_d, relu=1, want_stats=True))
time_graph("one transformer block", lambda: net._block(0, x, mr_x, first_u8, st[0], B, 1, prep, last=False))
time_graph("heads (pi GEMV + log-softmax + v)", lambda: pol._heads(lat, B, 1))
pd, v = pol._heads(lat, B, 1)
def samp():
    ac = pol.sample(pd)
    return ac, pol.logprob(ac, pd), pol.denormalize(v)
time_graph("sample + logprob + denormalise", samp)
time_graph("whole act()", lambda: pol.act({"img": img}, first, st))
# single kernels
Wc, _, bc = prep.layers[0]["qkvr"]
q = torch.empty((1, cfg.hidsize), dtype=torch.bfloat16, device="cuda")
time_graph("GEMV 6304 x 2048 (qkvr weights)", lambda: net._linear(x, prep.layers[0]["mlp0"], cfg.hidsize * cfg.pointwise_ratio, mr=mr_x, relu=1))
