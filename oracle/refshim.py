"""TEST INFRASTRUCTURE ONLY -- imports the UNMODIFIED reference (openai/Video-Pre-Training) from
/root/reference so that (a) the restatement in `oracle/vpt_oracle.py` can be validated against it and
(b) golden vectors can be generated (`oracle/make_golden.py`).

/root/reference does not exist on the GPU box, so nothing that runs there may import this module
(`available()` returns False there and callers skip).

The reference needs two third-party modules that are absent from this image; both only define *types*
(no arithmetic), so they are stubbed here:

* ``gym3.types`` -- ValType / Discrete / Real / TensorType / DictType, used by lib/action_head.py:9,263-275,
  lib/action_mapping.py:7,34-39,112-115,228-231 and lib/policy.py:7.  DictType iteration order is insertion
  order (this decides the RNG draw order camera -> buttons, lib/action_mapping.py:228-231).
* ``minerl.herobraine.hero.mc`` -- imported by lib/actions.py:2, only used for an item map we never touch.
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VPT_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "lib", "policy.py"))


def _install_stubs():
    if "gym3" not in sys.modules:
        gym3 = types.ModuleType("gym3")
        gtypes = types.ModuleType("gym3.types")

        class ValType:
            pass

        class Discrete(ValType):
            def __init__(self, n):
                self.n = n

            def __eq__(self, o):
                return isinstance(o, Discrete) and o.n == self.n

        class Real(ValType):
            pass

        class TensorType(ValType):
            def __init__(self, eltype, shape):
                self.eltype = eltype
                self.shape = tuple(shape)

            @property
            def size(self):
                n = 1
                for s in self.shape:
                    n *= s
                return n

        class DictType(ValType):
            def __init__(self, **kw):
                self._d = dict(kw)

            def items(self):
                return self._d.items()

            def keys(self):
                return self._d.keys()

            def __getitem__(self, k):
                return self._d[k]

        gtypes.ValType, gtypes.Discrete, gtypes.Real = ValType, Discrete, Real
        gtypes.TensorType, gtypes.DictType = TensorType, DictType
        gym3.types = gtypes
        sys.modules["gym3"] = gym3
        sys.modules["gym3.types"] = gtypes
    if "minerl" not in sys.modules:
        names = ["minerl", "minerl.herobraine", "minerl.herobraine.hero", "minerl.herobraine.hero.mc"]
        mods = [types.ModuleType(n) for n in names]
        for n, m in zip(names, mods):
            sys.modules[n] = m
        mods[0].herobraine = mods[1]
        mods[1].hero = mods[2]
        mods[2].mc = mods[3]
        mods[3].MINERL_ITEM_MAP = {}


_LOADED = None


def load():
    """Returns a namespace with the reference's `policy`, `action_mapping`, `torch_util` modules."""
    global _LOADED
    if _LOADED is not None:
        return _LOADED
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    _install_stubs()
    import warnings

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import lib.torch_util as tu  # noqa

        # lib/torch_util.py:36,44 evaluates th.has_cuda at import -> "cuda" on a CUDA build with no GPU.
        tu.set_default_torch_device("cpu")
        import lib.policy as policy  # noqa
        import lib.action_mapping as action_mapping  # noqa
    ns = types.SimpleNamespace(policy=policy, action_mapping=action_mapping, torch_util=tu,
                               DictType=sys.modules["gym3.types"].DictType)
    _LOADED = ns
    return ns


# agent.py:16-38 re-stated (agent.py itself needs gym + a MineRL env object, so it is not imported).
def policy_kwargs(width="2x", **over):
    w = {"1x": (4, 1024, 8), "2x": (8, 2048, 16), "3x": (12, 3072, 24)}[width] if isinstance(width, str) else width
    kw = dict(
        attention_heads=w[2], attention_mask_style="clipped_causal", attention_memory_size=256,
        diff_mlp_embedding=False, hidsize=w[1], img_shape=[128, 128, 3], impala_chans=[16, 32, 32],
        impala_kwargs={"post_pool_groups": 1}, impala_width=w[0],
        init_norm_kwargs={"batch_norm": False, "group_norm_groups": 1}, n_recurrence_layers=4,
        only_img_input=True, pointwise_ratio=4, pointwise_use_activation=False, recurrence_is_residual=True,
        recurrence_type="transformer", timesteps=128, use_pointwise_layer=True, use_pre_lstm_ln=False,
    )
    kw.update(over)
    return kw


TINY = dict(impala_width=1, hidsize=64, attention_heads=2, img_shape=[32, 32, 3], timesteps=8,
            attention_memory_size=16, n_recurrence_layers=2)


def make_reference_agent_policy(pkw, temperature=2.0, seed=0):
    """Builds lib.policy.MinecraftAgentPolicy exactly like agent.py:114-129 does (minus gym/env)."""
    import torch

    ns = load()
    mapper = ns.action_mapping.CameraHierarchicalMapping(n_camera_bins=11)
    action_space = ns.DictType(**mapper.get_action_space_update())
    torch.manual_seed(seed)
    pol = ns.policy.MinecraftAgentPolicy(action_space=action_space, policy_kwargs=pkw,
                                         pi_head_kwargs=dict(temperature=temperature))
    pol.eval()
    return pol
