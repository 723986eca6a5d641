"""Caller-side rows f-2 (frame ingest) and f-3 (action codec) + the MineRLAgent mirror.
CPU: the codec against the live reference (lib/action_mapping.py, lib/actions.py) where /root/reference exists, codec
properties everywhere, the resize oracle against cv2.  GPU: the resize kernel bit-exact against the oracle / cv2, agent smoke."""
import numpy as np
import pytest
import torch

import refshim
import resize_oracle
import vpt_b200
from video_pre_training_b200 import agent as A

try:
    import cv2
except Exception:  # pragma: no cover
    cv2 = None


def _random_factored(n, rng):
    btn = (rng.random((n, 20)) < 0.25).astype(np.int64)
    cam = rng.integers(0, 11, (n, 2))
    cam[rng.random(n) < 0.4] = 5
    return dict(buttons=btn, camera=cam)


@pytest.mark.skipif(not refshim.available(), reason="/root/reference not present (GPU box)")
def test_codec_matches_live_reference():
    import sys
    ns = refshim.load()
    import lib.actions as ref_actions  # noqa: E402  (importable once refshim.load() has set up sys.path + stubs)
    mapper = ns.action_mapping.CameraHierarchicalMapping(n_camera_bins=11)
    tr = ref_actions.ActionTransformer(**A.ACTION_TRANSFORMER_KWARGS)
    codec = A.ActionCodec(**A.ACTION_TRANSFORMER_KWARGS)
    assert codec.n_buttons_joint == len(mapper.BUTTONS_COMBINATIONS) == 8641
    assert np.array_equal(codec.idx_to_factored, mapper.BUTTON_IDX_TO_FACTORED)
    assert np.array_equal(codec.idx_camera_off, mapper.BUTTON_IDX_TO_CAMERA_META_OFF)
    rng = np.random.default_rng(0)
    joint = dict(buttons=rng.integers(0, 8641, (500, 1)), camera=rng.integers(0, 121, (500, 1)))
    a, b = codec.to_factored(joint), mapper.to_factored({k: v.copy() for k, v in joint.items()})
    assert np.array_equal(a["buttons"], b["buttons"]) and np.array_equal(a["camera"], b["camera"])
    fac = _random_factored(2000, rng)
    a, b = codec.from_factored(fac), mapper.from_factored({k: v.copy() for k, v in fac.items()})
    assert np.array_equal(a["buttons"], b["buttons"]) and np.array_equal(a["camera"], b["camera"])
    e1, e2 = codec.policy2env(fac), tr.policy2env({k: v.copy() for k, v in fac.items()})
    assert set(e1) == set(e2) and all(np.array_equal(e1[k], e2[k]) for k in e1)
    env = {"camera": rng.uniform(-15, 15, (300, 2)), "attack": rng.integers(0, 2, 300), "hotbar.3": rng.integers(0, 2, 300)}
    p1, p2 = codec.env2policy(env), tr.env2policy(env)
    assert np.array_equal(p1["camera"], p2["camera"]) and np.array_equal(p1["buttons"], p2["buttons"])
    assert codec.null_buttons_idx == mapper.get_zero_action()["buttons"] and codec.camera_null_idx == mapper.camera_null_idx


def test_codec_properties():
    codec = A.ActionCodec(**A.ACTION_TRANSFORMER_KWARGS)
    # every joint index survives joint -> factored -> joint, except that a non-null camera choice is dropped when the
    # button combination has the camera meta action off (lib/action_mapping.py:222-223)
    b = np.arange(8641)[:, None]
    for cam in (60, 0, 120, 37):
        fac = codec.to_factored(dict(buttons=b, camera=np.full_like(b, cam)))
        back = codec.from_factored(fac)
        off = codec.idx_camera_off[b[:, 0]]
        assert np.array_equal(back["camera"][off, 0], np.full(off.sum(), 60))
        assert np.array_equal(back["buttons"][off, 0], b[off, 0])          # meta-off combinations are fixed points
        on = ~off
        on[codec.inventory_idx] = False                                    # inventory is exclusive with the camera (:204-208)
        if cam != 60:  # (a camera-meta-ON combination whose camera choice is null maps back to its meta-OFF twin)
            assert np.array_equal(back["buttons"][on, 0], b[on, 0]) and np.all(back["camera"][on, 0] == cam)
        assert back["buttons"][codec.inventory_idx, 0] == codec.inventory_idx and back["camera"][codec.inventory_idx, 0] == 60
    # mu-law quantiser: bins 0..10 <-> [-10, 10], null bin 5 <-> 0, monotone, inverse on the bin centres
    centres = codec.undiscretize_camera(np.arange(11))
    assert centres[5] == 0 and np.all(np.diff(centres) > 0) and abs(centres[0] + 10) < 1e-9 and abs(centres[10] - 10) < 1e-9
    assert np.array_equal(codec.discretize_camera(centres), np.arange(11))


@pytest.mark.skipif(cv2 is None, reason="cv2 not importable")
def test_resize_oracle_is_bit_exact_with_cv2():
    rng = np.random.default_rng(0)
    for (H, W) in [(360, 640), (720, 1280), (128, 128), (200, 300), (431, 777), (129, 1399)]:
        img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
        ref = cv2.resize(img, (128, 128), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(resize_oracle.resize_linear_u8(img, 128, 128), ref), (H, W)


@pytest.mark.gpu
def test_resize_kernel_bit_exact():
    rng = np.random.default_rng(1)
    for (F_, H, W) in [(3, 360, 640), (1, 720, 1280), (2, 431, 777)]:
        img = rng.integers(0, 256, (F_, H, W, 3), dtype=np.uint8)
        got = A.resize_frames(torch.from_numpy(img).cuda()).cpu().numpy()
        for f in range(F_):
            assert np.array_equal(got[f], resize_oracle.resize_linear_u8(img[f], 128, 128))
            if cv2 is not None:
                assert np.array_equal(got[f], cv2.resize(img[f], (128, 128), interpolation=cv2.INTER_LINEAR))


@pytest.mark.gpu
def test_minerl_agent_rollout_smoke():
    kw = vpt_b200.policy_kwargs("1x", n_recurrence_layers=1)
    torch.manual_seed(0)
    agent = A.MineRLAgent(device="cuda", policy_kwargs=kw, pi_head_kwargs=vpt_b200.PI_HEAD_KWARGS)
    rng = np.random.default_rng(2)
    for _ in range(3):
        act = agent.get_action({"pov": rng.integers(0, 256, (360, 640, 3), dtype=np.uint8)})
        assert set(act) == set(A.BUTTONS) | {"camera"} and act["camera"].shape == (1, 2) and act["attack"].shape == (1,)
    agent.reset()
    back = agent._env_action_to_agent({k: (np.asarray(v) if k == "camera" else np.asarray(v)) for k, v in act.items()})
    assert back["buttons"].shape == (1, 1) and back["camera"].shape == (1, 1)
