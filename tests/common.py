"""Shared helpers for the parity tests."""
import torch

import vpt_b200
import vpt_oracle as O

SMALL = dict(img_shape=[32, 32, 3], hidsize=256, attention_heads=2, timesteps=8, attention_memory_size=16, n_recurrence_layers=2)


def small_kwargs(**over):
    kw = vpt_b200.policy_kwargs("1x", **SMALL)
    kw.update(over)
    return kw


def perturb(pol, seed=1):
    """Random-init blind-spot breaker: randomise every norm affine / bias, q weights x30."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in pol.named_parameters():
            if ".norm." in n or n.endswith(".bias") or "_ln." in n or ".n." in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            if "q_layer.weight" in n:
                p.mul_(30.0)


def make_policy(kw, seed=0, pert=True):
    torch.manual_seed(seed)
    pol = vpt_b200.MinecraftAgentPolicy(vpt_b200.minecraft_action_space(), kw, vpt_b200.PI_HEAD_KWARGS)
    if pert:
        perturb(pol)
    sd = {k: v.detach().clone() for k, v in pol.state_dict().items()}
    return pol, sd, O.Cfg(**kw)


def rel_err(a, b):
    return ((a.float() - b.float()).abs() / b.float().abs().clamp(min=1e-30)).max().item()


def l2_err(a, b):
    return ((a.float() - b.float()).norm() / b.float().norm().clamp(min=1e-30)).item()


def run_chunks(pol, sd, cfg, B, chunks, dev, first_at=None, seed=0, taps=False):
    """Feeds the same synthetic chunks to the CUDA policy (`dev`) and the oracle (CPU); yields per-chunk outputs."""
    g = torch.Generator().manual_seed(seed)
    st = pol.initial_state(B)
    st_o = O.initial_state(cfg, B)
    H, W, _ = cfg.img_shape
    out = []
    for ci, T in enumerate(chunks):
        img = torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8, generator=g)
        first = torch.zeros(B, T, dtype=torch.bool)
        if first_at is not None and ci == first_at[0]:
            first[first_at[1], 0] = True
        to = {} if taps else None
        if taps:
            pol.net.debug_taps = {}
        (pd, v, _), st = pol({"img": img.to(dev)}, first.to(dev), st)
        with torch.no_grad():
            (pd_o, v_o, _), st_o = O.agent_policy_forward(sd, cfg, img, first, st_o, taps=to)
        out.append(dict(pd=pd, v=v, st=st, pd_o=pd_o, v_o=v_o, st_o=st_o, taps=dict(pol.net.debug_taps or {}), taps_o=to))
    pol.net.debug_taps = None
    return out


import contextlib


@contextlib.contextmanager
def emulation():
    """Temporarily route video_pre_training_b200.ops through the test-only torch emulation (CPU tensors)."""
    import emu_ops
    from video_pre_training_b200 import ops

    saved = {}
    for name in dir(emu_ops):
        if not name.startswith("_") and callable(getattr(emu_ops, name)) and hasattr(ops, name):
            saved[name] = getattr(ops, name)
            setattr(ops, name, getattr(emu_ops, name))
    try:
        yield
    finally:
        for name, fn in saved.items():
            setattr(ops, name, fn)
