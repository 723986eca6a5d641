"""ORACLE -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A CPU fp32 restatement (functional, weights come from a reference-schema ``state_dict``) of the VPT policy
forward pass of openai/Video-Pre-Training.  Every function cites the reference file:line it follows
(paths relative to /root/reference).  Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this module, and only as the checker or as the timed
CPU baseline -- the product path (``video-pre-training_b200``) never imports it and has no CPU fallback.

Parity pin: the reference ships no golden vectors or tests, so the oracle is pinned by
running the UNMODIFIED reference in the build container (``oracle/refshim.py``): ``tests/test_oracle.py``
compares this file against the live reference whenever /root/reference is present, and against fixtures in
``tests/golden/*.pt`` (made by ``oracle/make_golden.py`` from the live reference) everywhere else.

All arithmetic is torch CPU fp32 (the reference's own arithmetic library, requirements.txt:1), so it is
bit-for-bit the same op sequence where that matters (conv2d / group_norm / layer_norm / baddbmm / softmax /
log_softmax) and differs only in glue (mask built in closed form, no concat/unfold helpers).
"""
import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ----------------------------------------------------------------------------------------------------------
# config
# ----------------------------------------------------------------------------------------------------------
class Cfg:
    """The subset of lib/policy.py:96-126 MinecraftPolicy kwargs the transformer models use (agent.py:16-36)."""

    def __init__(self, impala_width=8, impala_chans=(16, 32, 32), hidsize=2048, attention_heads=16,
                 attention_memory_size=256, timesteps=128, n_recurrence_layers=4, img_shape=(128, 128, 3),
                 pointwise_ratio=4, temperature=2.0, attention_mask_style="clipped_causal",
                 first_conv_norm=False, conv3d=None, **unused):
        self.chans = tuple(int(impala_width * c) for c in impala_chans)  # policy.py:136
        self.hidsize = hidsize
        self.heads = attention_heads
        self.maxlen = attention_memory_size - timesteps  # masked_attention.py:137
        self.timesteps = timesteps
        self.n_layers = n_recurrence_layers
        self.img_shape = tuple(img_shape)
        self.pointwise_ratio = pointwise_ratio
        self.temperature = temperature
        self.mask_style = attention_mask_style
        self.first_conv_norm = first_conv_norm
        self.conv3d = conv3d
        if attention_mask_style == "none":
            self.maxlen = attention_memory_size - timesteps


def widths(name: str):
    return {"1x": dict(impala_width=4, hidsize=1024, attention_heads=8),
            "2x": dict(impala_width=8, hidsize=2048, attention_heads=16),
            "3x": dict(impala_width=12, hidsize=3072, attention_heads=24)}[name]


# ----------------------------------------------------------------------------------------------------------
# CNN
# ----------------------------------------------------------------------------------------------------------
def img_preprocess(img_u8: Tensor) -> Tensor:
    """lib/policy.py:39-45 -- u8 -> f32, divide by ob_scale=255."""
    return img_u8.to(torch.float32) / 255.0


def fanin_conv(x: Tensor, sd: Dict[str, Tensor], p: str) -> Tensor:
    """lib/util.py:75-82 (conv flavour): [GroupNorm(1 group) on the INPUT] -> conv3x3 pad1 -> ReLU.
    bias exists iff there is no norm (lib/util.py:65)."""
    if p + ".norm.weight" in sd:
        x = F.group_norm(x, 1, sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    x = F.conv2d(x, sd[p + ".layer.weight"], sd.get(p + ".layer.bias"), padding=1)
    return F.relu(x, inplace=True)  # in place like the reference (lib/util.py:81); also what keeps the CPU baseline fair


def fanin_linear(x: Tensor, sd: Dict[str, Tensor], p: str, relu: bool = True) -> Tensor:
    """lib/util.py:75-82 (linear flavour): [LayerNorm] -> Linear -> [ReLU]."""
    if p + ".norm.weight" in sd:
        x = F.layer_norm(x, (x.shape[-1],), sd[p + ".norm.weight"], sd[p + ".norm.bias"], eps=1e-5)
    x = F.linear(x, sd[p + ".layer.weight"], sd.get(p + ".layer.bias"))
    return F.relu(x, inplace=True) if relu else x


def cnn_basic_block(x: Tensor, sd, p: str, taps=None) -> Tensor:
    """lib/impala_cnn.py:50-52 -- x + conv1(conv0(x)); the ReLU of conv1 is applied BEFORE the add."""
    h = fanin_conv(x, sd, p + ".conv0")
    if taps is not None:
        taps[p + ".conv0"] = h
    y = x + fanin_conv(h, sd, p + ".conv1")
    if taps is not None:
        taps[p] = y
    return y


def cnn_down_stack(x: Tensor, sd, p: str, taps=None) -> Tensor:
    """lib/impala_cnn.py:114-121 -- firstconv -> max_pool2d(3,2,1) -> GroupNorm(1) -> 2 residual blocks."""
    x = fanin_conv(x, sd, p + ".firstconv")
    x = F.max_pool2d(x, kernel_size=3, stride=2, padding=1)
    if taps is not None:
        taps[p + ".pool"] = x
    x = F.group_norm(x, 1, sd[p + ".n.weight"], sd[p + ".n.bias"], eps=1e-5)
    if taps is not None:
        taps[p + ".n"] = x
    j = 0
    while f"{p}.blocks.{j}.conv0.layer.weight" in sd:
        x = cnn_basic_block(x, sd, f"{p}.blocks.{j}", taps)
        j += 1
    return x


def impala_cnn(x_bthwc: Tensor, sd, p: str, taps=None) -> Tensor:
    """lib/impala_cnn.py:187-195 -- (b,t)->bt, NHWC->NCHW, stacks, flatten in C,H,W order, dense."""
    b, t = x_bthwc.shape[:2]
    x = x_bthwc.reshape(b * t, *x_bthwc.shape[2:]).permute(0, 3, 1, 2)
    i = 0
    while f"{p}.stacks.{i}.firstconv.layer.weight" in sd:
        x = cnn_down_stack(x, sd, f"{p}.stacks.{i}", taps)
        i += 1
    x = x.reshape(b, t, -1)  # c*H*W + h*W + w (lib/torch_util.py:107-112)
    x = fanin_linear(x, sd, p + ".dense")
    if taps is not None:
        taps[p + ".dense"] = x
    return x


def conv3d_stage(x_bthwc: Tensor, sd, p: str) -> Tensor:
    """lib/policy.py:394-403 (IDM): conv3d k=(5,1,1) pad (2,0,0) 3->128 with bias, ReLU, per-sample loop."""
    x = x_bthwc.permute(0, 4, 1, 2, 3)  # b c t h w
    outs = []
    for mb in torch.split(x, 1):
        outs.append(F.relu(F.conv3d(mb, sd[p + ".layer.weight"], sd[p + ".layer.bias"], padding=(2, 0, 0))))
    x = torch.cat(outs)
    return x.permute(0, 2, 3, 4, 1)


# ----------------------------------------------------------------------------------------------------------
# transformer
# ----------------------------------------------------------------------------------------------------------
def allowed_mask(first_b: Tensor, state_mask: Optional[Tensor], t: int, maxlen: int) -> Tuple[Tensor, Tensor]:
    """Closed form of lib/masked_attention.py:11-44 (band) and :47-94 (first / state_mask logic).

    allowed[b,i,j] = [0 <= d < maxlen] and ( j >= T-t  or  (not first[b] and state_mask[b,j]) ),
    d = (T-t+i) - j, T = maxlen + t.  state_mask=None means all-False (:75-76).
    new state_mask = cat(state_mask[:, t:] & ~first, ones(min(t, T-t)))  (:86-92)."""
    T = maxlen + t
    b = first_b.shape[0]
    if state_mask is None:
        state_mask = torch.zeros((b, 1, T - t), dtype=torch.bool)
    i = torch.arange(t)[:, None]
    j = torch.arange(T)[None, :]
    d = (T - t + i) - j
    band = (d >= 0) & (d < maxlen) if maxlen < T else (d >= 0)
    m = band[None].repeat(b, 1, 1)
    not_first = ~first_b.reshape(b, 1, 1)
    m[:, :, : T - t] &= not_first
    m[:, :, : T - t] &= state_mask
    new_state_mask = torch.cat([state_mask[:, :, t:] & not_first,
                                torch.ones((b, 1, min(t, T - t)), dtype=torch.bool)], dim=-1)
    return m, new_state_mask


def split_heads(x_bte: Tensor, h: int) -> Tensor:
    """lib/xf.py:96-103 -- head-major columns: (b,t,h*q) -> (b*h,t,q)."""
    b, t, e = x_bte.shape
    return x_bte.reshape(b, t, h, e // h).permute(0, 2, 1, 3).reshape(b * h, t, e // h)


def rel_logits(xhat: Tensor, sd, p: str, heads: int, t: int, T: int) -> Tensor:
    """lib/xf.py:265-271 + lib/util.py:232-267: extra[bh,i,j] = sum_n R[bh,i,n] * b_nd[n, (T-t+i)-j],
    zero outside 0 <= (T-t+i)-j < maxlen."""
    R = F.linear(xhat, sd[p + ".r_layer.weight"], sd[p + ".r_layer.bias"]).float()
    R = split_heads(R, heads)  # (b*h, t, nbasis)
    b_nd = sd[p + ".b_nd"]
    nb, band = b_nd.shape
    i = torch.arange(t)[:, None]
    j = torch.arange(T)[None, :]
    d = (T - t + i) - j
    ok = (d >= 0) & (d < band)
    D = torch.zeros(nb, t, T)
    D[:, ok] = b_nd[:, d[ok]]
    return torch.einsum("btn,ntp->btp", R, D)


def self_attention_layer(xhat: Tensor, state, sd, p: str, heads: int, maxlen: int, mask: Optional[Tensor]):
    """lib/xf.py:334-391 -- Q(+bias),K,V projections, KV memory concat + trim, banded masked attention with
    muP 1/dh scale (xf.py:59), proj(+bias), residual onto the layer INPUT (xf.py:358-360)."""
    q = F.linear(xhat, sd[p + ".q_layer.weight"], sd[p + ".q_layer.bias"])
    k = F.linear(xhat, sd[p + ".k_layer.weight"])
    v = F.linear(xhat, sd[p + ".v_layer.weight"])
    mem_k, mem_v = state
    # update_state (xf.py:366-391), cache_keep_len == maxlen
    start = max(mem_k.shape[1] - maxlen, 0)
    full_k = torch.cat([mem_k[:, start:], k], dim=1)
    full_v = torch.cat([mem_v[:, start:], v], dim=1)
    new_state = (full_k[:, max(full_k.shape[1] - maxlen, 0):], full_v[:, max(full_v.shape[1] - maxlen, 0):])
    t, T = q.shape[1], full_k.shape[1]
    Q, K, V = split_heads(q, heads), split_heads(full_k, heads), split_heads(full_v, heads)
    e = Q.shape[2]
    if p + ".b_nd" in sd and sd[p + ".b_nd"].shape[1] > 0:
        extra = rel_logits(xhat, sd, p, heads, t, T)
    else:
        extra = torch.zeros(())
    if mask is not None:
        bias = (~mask).float().repeat_interleave(heads, dim=0) * -1e9  # xf.py:46
    else:
        bias = torch.zeros(())
    bias = bias + extra
    if bias.dim() == 0:
        bias = bias.expand(Q.shape[0], t, T)
    logit = torch.baddbmm(bias, Q.float(), K.float().transpose(-1, -2), alpha=1.0 / e)  # xf.py:55-60
    W = torch.softmax(logit, dim=2)
    A = torch.einsum("btp,bpe->bte", W, V)
    b = xhat.shape[0]
    A = A.reshape(b, heads, t, e).permute(0, 2, 1, 3).reshape(b, t, heads * e)  # xf.py:123-129
    out = F.linear(A, sd[p + ".proj_layer.weight"], sd[p + ".proj_layer.bias"])
    return xhat + out, new_state


def residual_recurrent_block(x: Tensor, first: Tensor, state, sd, p: str, cfg: Cfg, taps=None):
    """lib/util.py:193-211 (transformer branch) + lib/masked_attention.py:161-178."""
    xhat = F.layer_norm(x, (x.shape[-1],), sd[p + ".pre_r_ln.weight"], sd[p + ".pre_r_ln.bias"], eps=1e-5)
    state_mask, kv = state
    t = x.shape[1]
    if cfg.mask_style == "clipped_causal":
        mask, state_mask = allowed_mask(first[:, 0], state_mask, t, cfg.maxlen)  # only first[:,0] is read (:167)
    else:
        mask = None
    y, kv = self_attention_layer(xhat, kv, sd, p + ".r.orc_block", cfg.heads, cfg.maxlen, mask)
    if taps is not None:
        taps[p + ".attn"] = y
    h = fanin_linear(y, sd, p + ".mlp0")
    z = y + F.linear(h, sd[p + ".mlp1.layer.weight"], sd[p + ".mlp1.layer.bias"])  # no activation (util.py:166)
    if taps is not None:
        taps[p] = z
    return z, (state_mask, kv)


def initial_state(cfg: Cfg, batch: int):
    """lib/policy.py:220-224 -> lib/util.py:125-129 -> lib/masked_attention.py:153-159 -> lib/xf.py:393-397."""
    return [(None, (torch.zeros(batch, cfg.maxlen, cfg.hidsize), torch.zeros(batch, cfg.maxlen, cfg.hidsize)))
            for _ in range(cfg.n_layers)]


# ----------------------------------------------------------------------------------------------------------
# whole net
# ----------------------------------------------------------------------------------------------------------
def minecraft_policy_forward(sd, cfg: Cfg, img_u8: Tensor, first: Tensor, state_in, prefix="net", taps=None):
    """lib/policy.py:193-218 (and :374-392 for the IDM variant when cfg.conv3d is set)."""
    x = img_preprocess(img_u8)
    if cfg.conv3d is not None:
        x = conv3d_stage(x, sd, prefix + ".conv3d_layer")
    x = impala_cnn(x, sd, prefix + ".img_process.cnn", taps)
    x = fanin_linear(x, sd, prefix + ".img_process.linear")
    if taps is not None:
        taps[prefix + ".img_process"] = x
    state_out = []
    assert len(state_in) == cfg.n_layers  # lib/util.py:117-119
    for l in range(cfg.n_layers):
        x, s = residual_recurrent_block(x, first, state_in[l], sd, f"{prefix}.recurrent_layer.blocks.{l}", cfg, taps)
        state_out.append(s)
    x = F.relu(x)
    if cfg.conv3d is None:
        x = fanin_linear(x, sd, prefix + ".lastlayer")
    # IDM: lastlayer output is discarded (policy.py:390-391) -> final_ln(relu(x))
    x = F.layer_norm(x, (x.shape[-1],), sd[prefix + ".final_ln.weight"], sd[prefix + ".final_ln.bias"], eps=1e-5)
    if taps is not None:
        taps[prefix + ".latent"] = x
    return x, state_out


def categorical_head(latent: Tensor, sd, p: str, shape: Tuple[int, ...], n: int, temperature: float) -> Tensor:
    """lib/action_head.py:163-174 -- Linear -> reshape (..., *shape, n) -> /temperature -> fp32 log_softmax."""
    out = F.linear(latent, sd[p + ".linear_layer.weight"], sd[p + ".linear_layer.bias"])
    out = out.reshape(out.shape[:-1] + tuple(shape) + (n,)) / temperature
    return F.log_softmax(out.float(), dim=-1)


def agent_policy_forward(sd, cfg: Cfg, img_u8: Tensor, first: Tensor, state_in, taps=None):
    """lib/policy.py:252-269 -- ((pi_logits{camera,buttons}, vpred, None), state_out).
    Head order camera -> buttons (lib/action_mapping.py:228-231)."""
    latent, state_out = minecraft_policy_forward(sd, cfg, img_u8, first, state_in, "net", taps)
    ncam = sd["pi_head.camera.linear_layer.weight"].shape[0]
    nbut = sd["pi_head.buttons.linear_layer.weight"].shape[0]
    pd = {"camera": categorical_head(latent, sd, "pi_head.camera", (1,), ncam, cfg.temperature),
          "buttons": categorical_head(latent, sd, "pi_head.buttons", (1,), nbut, cfg.temperature)}
    vpred = F.linear(latent, sd["value_head.linear.weight"], sd["value_head.linear.bias"])  # scaled_mse_head.py:34-35
    return (pd, vpred, None), state_out


def idm_policy_forward(sd, cfg: Cfg, img_u8: Tensor, first: Tensor, state_in, taps=None):
    """lib/policy.py:432-446 -- IDM heads: buttons (20 x Discrete 2), camera (2 x Discrete 11)
    (lib/action_mapping.py:110-115; dict order buttons -> camera)."""
    latent, state_out = minecraft_policy_forward(sd, cfg, img_u8, first, state_in, "net", taps)
    pd = {"buttons": categorical_head(latent, sd, "pi_head.buttons", (20,), 2, cfg.temperature),
          "camera": categorical_head(latent, sd, "pi_head.camera", (2,), 11, cfg.temperature)}
    return (pd, None, None), state_out


def denormalize_value(sd, v: Tensor) -> Tensor:
    """lib/normalize_ewma.py:31-35,57-60 (norm_axes=2, epsilon=1e-5)."""
    deb = sd["value_head.normalizer.debiasing_term"].clamp(min=1e-5)
    mean = sd["value_head.normalizer.running_mean"] / deb
    mean_sq = sd["value_head.normalizer.running_mean_sq"] / deb
    var = (mean_sq - mean ** 2).clamp(min=1e-2)
    return v * torch.sqrt(var)[None, None] + mean[None, None]


def gumbel_sample(logits: Tensor, u: Tensor) -> Tensor:
    """lib/action_head.py:195-207 given the uniform draws u = rand_like(logits)."""
    u = u.clone()
    u[u == 1.0] = 0.999
    return torch.argmax(logits - torch.log(-torch.log(u)), dim=-1)


def sample(pd: Dict[str, Tensor], deterministic: bool = False) -> Dict[str, Tensor]:
    """lib/action_head.py:253-254 -- per head in dict order, consuming the global torch RNG like rand_like."""
    out = {}
    for k, lg in pd.items():
        out[k] = torch.argmax(lg, dim=-1) if deterministic else gumbel_sample(lg, torch.rand_like(lg))
    return out


def logprob(pd: Dict[str, Tensor], ac: Dict[str, Tensor]) -> Tensor:
    """lib/action_head.py:176-184 + :250-251 -- gather the log-pmf at the action, sum over sub-actions, heads."""
    tot = 0
    for k, lg in pd.items():
        r = lg.gather(-1, ac[k].long().unsqueeze(-1)).squeeze(-1)
        tot = tot + r.sum(dim=-1)
    return tot


def forward_flops_per_frame(cfg: Cfg, ncam=121, nbut=8641) -> float:
    """2*MAC, banded attention = maxlen keys per query."""
    H, W, _ = cfg.img_shape
    fl, cin = 0.0, 3
    for c in cfg.chans:
        fl += 2 * H * W * 9 * cin * c
        H, W = (H + 1) // 2, (W + 1) // 2
        fl += 4 * 2 * H * W * 9 * c * c
        cin = c
    h = cfg.hidsize
    fl += 2 * (cin * H * W) * 256 + 2 * 256 * h
    per_layer = 2 * h * h * 4 + 2 * h * 10 * cfg.heads + 2 * 2 * h * h * cfg.pointwise_ratio
    per_layer += 4 * cfg.maxlen * h + 2 * 10 * cfg.maxlen * cfg.heads
    fl += cfg.n_layers * per_layer + 2 * h * h + 2 * h * (ncam + nbut + 1)
    return fl
