"""Minimal action-space value types with the attribute surface the policy needs (`.items()`, `.shape`, `.eltype.n`).
The reference takes gym3.types objects (lib/action_head.py:263-275); anything duck-typing that surface -- including
real gym3 types -- is accepted by `MinecraftAgentPolicy`."""


class Discrete:
    def __init__(self, n):
        self.n = int(n)


class TensorType:
    def __init__(self, shape, eltype):
        self.shape = tuple(shape)
        self.eltype = eltype

    @property
    def size(self):
        n = 1
        for s in self.shape:
            n *= s
        return n


class DictType:
    def __init__(self, **spaces):
        self._spaces = dict(spaces)  # insertion order == sampling order (lib/action_head.py:253-254)

    def items(self):
        return self._spaces.items()

    def keys(self):
        return self._spaces.keys()

    def __getitem__(self, k):
        return self._spaces[k]


def minecraft_action_space():
    """lib/action_mapping.py:228-231 for CameraHierarchicalMapping(n_camera_bins=11): camera 11*11 joint bins,
    buttons 10*3*3*3*2*2*2*2*2 + 1 (inventory) joint combinations (:128-132)."""
    return DictType(camera=TensorType((1,), Discrete(121)), buttons=TensorType((1,), Discrete(8641)))


def idm_action_space():
    """lib/action_mapping.py:110-115 (IDMActionMapping)."""
    return DictType(buttons=TensorType((20,), Discrete(2)), camera=TensorType((2,), Discrete(11)))
