"""Generates tests/golden/*.pt from the UNMODIFIED reference (run in the build container, where /root/reference exists):

    python oracle/make_golden.py

Each fixture = {policy_kwargs, temperature, state_dict (fp32, reference schema), chunks: [{img u8, first bool,
camera log-probs, buttons log-probs of the chunk's last frame, vpred, state_out K/V of layer 0, state masks}], sample: indices under manual_seed(1234)}.
`tiny_*` use the smallest config the unmodified reference accepts so the files stay small; they pin
`oracle/vpt_oracle.py` on machines where the reference is absent (the GPU box).  The `perturbed` variant randomises every
norm affine / bias and scales q weights x30 so that layout mistakes that plain init hides (gamma=1, beta=0, near-uniform
attention) show up.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import refshim  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def perturb(pol, seed=1):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for n, p in pol.named_parameters():
            if ".norm." in n or n.endswith(".bias") or "_ln." in n or ".n." in n:
                p.add_(torch.randn(p.shape, generator=g) * 0.1)
            if "q_layer.weight" in n:
                p.mul_(30.0)


def make(name, pkw, chunks, B, pert, seed=0):
    pol = refshim.make_reference_agent_policy(pkw, temperature=2.0, seed=seed)
    if pert:
        perturb(pol)
    g = torch.Generator().manual_seed(seed + 100)
    H, W, _ = pkw["img_shape"]
    st = pol.initial_state(B)
    rec = []
    with torch.no_grad():
        for ci, T in enumerate(chunks):
            img = torch.randint(0, 256, (B, T, H, W, 3), dtype=torch.uint8, generator=g)
            first = torch.zeros(B, T, dtype=torch.bool)
            if ci == 2:
                first[B - 1, 0] = True
            (pd, v, _), st = pol({"img": img}, first, st)
            rec.append(dict(img=img, first=first, camera=pd["camera"].clone(), buttons_last=pd["buttons"][:, -1:].clone(), vpred=v.clone(),
                            k0=st[0][1][0].clone(), v0=st[0][1][1].clone(), masks=[s[0].clone() for s in st]))
        torch.manual_seed(1234)
        ac = pol.pi_head.sample(pd)
        lp = pol.pi_head.logprob(ac, pd)
    fx = dict(policy_kwargs=pkw, temperature=2.0, B=B, state_dict={k: v.clone() for k, v in pol.state_dict().items()},
              chunks=rec, sample={k: v.clone() for k, v in ac.items()}, sample_logprob=lp.clone())
    os.makedirs(OUT, exist_ok=True)
    torch.save(fx, os.path.join(OUT, name + ".pt"))
    print(name, os.path.getsize(os.path.join(OUT, name + ".pt")) // 1024, "KiB")


if __name__ == "__main__":
    tiny = refshim.policy_kwargs("2x", impala_width=1, hidsize=32, attention_heads=2, img_shape=[32, 32, 3], timesteps=8,
                                 attention_memory_size=16, n_recurrence_layers=2)
    make("tiny_plain", tiny, [8, 3, 8, 1], B=2, pert=False)
    make("tiny_perturbed", tiny, [8, 3, 8, 1], B=2, pert=True)
