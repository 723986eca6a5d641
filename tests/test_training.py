"""BC step: the hand-written backward (video-pre-training_b200/training.py) against torch autograd
through the oracle.  On CPU the ops are the test-only torch emulation, so this checks the host-side chain rule, the
weight-layout round trips and the norm-fold algebra; tests/test_gpu_training.py repeats it through the CUDA kernels."""
import pytest
import torch

import vpt_b200
import vpt_oracle as O
from common import make_policy, small_kwargs
from test_host_logic import emulated  # noqa: F401  (fixture)
from video_pre_training_b200.training import BCTrainer


def oracle_grads(sd, cfg, img, first, state, actions):
    """d(-mean log-prob)/d(param) by autograd through the oracle (behavioural_cloning.py:101-123 for one batch)."""
    leaf = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    (pd, _, _), st = O.agent_policy_forward(leaf, cfg, img, first, state)
    lp = O.logprob(pd, actions)
    loss = -lp.mean()
    loss.backward()
    return loss.detach(), {k: v.grad for k, v in leaf.items()}, st


def run_case(dev, B=2, T=8, chunks=2, seed=0, reset_at=None):
    pol, sd, cfg = make_policy(small_kwargs())
    pol = pol.to(dev)
    g = torch.Generator().manual_seed(seed)
    tr = BCTrainer(pol)
    st, st_o = pol.initial_state(B), O.initial_state(cfg, B)
    out = []
    for c in range(chunks):
        img = torch.randint(0, 256, (B, T, 32, 32, 3), dtype=torch.uint8, generator=g)
        first = torch.zeros(B, T, dtype=torch.bool)
        if reset_at is not None and c == reset_at[0]:
            first[reset_at[1], 0] = True  # episode boundary: this row must not see (or back-propagate into) its old memory
        actions = {"camera": torch.randint(0, 121, (B, T, 1), generator=g), "buttons": torch.randint(0, 8641, (B, T, 1), generator=g)}
        for p in pol.parameters():
            p.grad = None
        loss, st = tr.loss_and_grad(img.to(dev), first.to(dev), st, {k: v.to(dev) for k, v in actions.items()})
        loss_o, grads_o, st_o = oracle_grads(sd, cfg, img, first, st_o, actions)
        st_o = [(m, (k.detach(), v.detach())) for (m, (k, v)) in st_o]
        out.append((loss, loss_o, {n: p.grad for n, p in pol.named_parameters()}, grads_o))
    return out


def check(out, tol_l2=5e-2):
    for loss, loss_o, grads, grads_o in out:
        assert abs(loss.item() - loss_o.item()) < 1e-2 * abs(loss_o.item())
        worst = {}
        for n, g_o in grads_o.items():
            if n.startswith("value_head"):
                assert g_o is None and grads[n] is None  # the BC loss never touches the value head
                continue
            if g_o is None:
                continue
            g = grads[n]
            assert g is not None, f"no gradient for {n}"
            assert g.shape == g_o.shape and g.dtype == torch.float32
            err = ((g.cpu() - g_o).norm() / g_o.norm().clamp(min=1e-12)).item()
            worst[n] = err
        bad = {n: e for n, e in worst.items() if e > tol_l2}
        assert not bad, f"gradient rel-L2 error above {tol_l2}: {sorted(bad.items(), key=lambda kv: -kv[1])[:8]}"


@pytest.fixture()
def exact(monkeypatch):
    """fp32 everywhere the kernels would store bf16: the emulated step is then the same function as the oracle."""
    import emu_ops
    from video_pre_training_b200 import policy, training

    for m in (emu_ops, policy, training):
        monkeypatch.setattr(m, "BF16", torch.float32)
    yield


def test_bc_backward_is_the_exact_gradient(emulated, exact):
    """With the bf16 rounding switched off the hand-written backward must reproduce autograd through the oracle: this pins the
    chain rule, the norm-fold algebra, the weight-layout round trips, the KV-memory detach and the loss scaling.  Tolerance:
    the folded forward differs from the oracle by ~1e-6, which flips the ReLU mask of an element that sits at zero once in a
    few million elements; one flip moves a frame's gradient by a few percent, everything else agrees to ~1e-6."""
    out = run_case("cpu") + run_case("cpu", chunks=3, seed=1, reset_at=(1, 1))  # second run: an episode reset in chunk 1, row 1
    check(out, tol_l2=5e-2)
    exact_params = 0
    for _, _, grads, grads_o in out:
        for n, g_o in grads_o.items():
            if g_o is not None and not n.startswith("net.img_process.cnn"):
                assert ((grads[n] - g_o).norm() / g_o.norm()).item() < 1e-3, n
                exact_params += 1
    assert exact_params > 80


def test_bc_step_with_bf16_rounding_points_emulated(emulated):
    """Same step with every bf16 rounding point of the kernels emulated.  A gradient is a discontinuous function of the forward
    activations (ReLU / max-pool masks): the ~1 % forward difference between a bf16 and an fp32 forward flips ~1 % of the
    masks, i.e. ~10 % gradient noise per ReLU layer, so against the fp32 oracle only the direction can be checked."""
    for loss, loss_o, grads, grads_o in run_case("cpu", chunks=1):
        assert abs(loss.item() - loss_o.item()) < 1e-2 * abs(loss_o.item())
        for n, g_o in grads_o.items():
            if g_o is None:
                continue
            g = grads[n]
            assert torch.isfinite(g).all()
            cos = (g * g_o).sum() / (g.norm() * g_o.norm()).clamp(min=1e-20)
            assert cos > 0.8, (n, cos.item())


def test_trainer_refuses_what_the_reference_does_not_train():
    pol, _, _ = make_policy(small_kwargs())
    with pytest.raises(TypeError):
        BCTrainer(pol.net)  # needs the agent policy (heads + loss), behavioural_cloning.py:54-62
    idm = vpt_b200.InverseActionPolicy(vpt_b200.idm_action_space(), dict(temperature=2.0),
                                       vpt_b200.idm_net_kwargs(img_shape=[32, 32, 128], hidsize=256, attention_heads=2, timesteps=8,
                                                               attention_memory_size=8, n_recurrence_layers=1, impala_width=4))
    with pytest.raises(TypeError):
        BCTrainer(idm)


def test_kv_memory_is_detached_between_steps(emulated):
    """behavioural_cloning.py:111: the state carried to the next chunk must not require grad or alias the tape."""
    pol, _, _ = make_policy(small_kwargs())
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (1, 8, 32, 32, 3), dtype=torch.uint8, generator=g)
    first = torch.zeros(1, 8, dtype=torch.bool)
    actions = {"camera": torch.randint(0, 121, (1, 8, 1), generator=g), "buttons": torch.randint(0, 8641, (1, 8, 1), generator=g)}
    _, st = BCTrainer(pol).loss_and_grad(img, first, pol.initial_state(1), actions)
    for mask, (k, v) in st:
        assert mask.dtype == torch.bool and not k.requires_grad and not v.requires_grad and k.dtype == torch.float32
        assert k.shape == (1, 8, 256)


def test_backward_matches_autograd_at_the_taped_operating_point(emulated):
    """The tight gradient pin: autograd through the FORCED replica (tests/forced_replica.py: every layer
    recomputed in fp32 from the parameters, values and ReLU / pool masks taken from the forward's tape) against the hand-written
    backward, with every bf16 rounding point active.  No mask can flip, so the bound is per parameter and tight."""
    from forced_replica import forced_loss

    pol, sd, cfg = make_policy(small_kwargs())
    g = torch.Generator().manual_seed(0)
    img = torch.randint(0, 256, (2, 8, 32, 32, 3), dtype=torch.uint8, generator=g)
    first = torch.zeros(2, 8, dtype=torch.bool)
    actions = {"camera": torch.randint(0, 121, (2, 8, 1), generator=g), "buttons": torch.randint(0, 8641, (2, 8, 1), generator=g)}
    tr = BCTrainer(pol)
    tr.keep_tape = True
    loss, _ = tr.loss_and_grad(img, first, pol.initial_state(2), actions)
    leaf = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    lf = forced_loss(leaf, cfg, tr.last_tape, img, first, actions)
    lf.backward()
    assert abs(loss.item() - lf.item()) < 1e-4 * abs(lf.item())
    for n, p in pol.named_parameters():
        if n.startswith("value_head"):
            assert p.grad is None
            continue
        e = ((p.grad - leaf[n].grad).norm() / leaf[n].grad.norm()).item()
        assert e < 2e-2, (n, e)
