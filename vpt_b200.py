"""Import shim: the package directory is `video-pre-training_b200/` (hyphenated, not importable by name), so
`import vpt_b200` loads it under the module name `video_pre_training_b200` and aliases this module to it."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "video-pre-training_b200")
_name = "video_pre_training_b200"
if _name in sys.modules:
    _mod = sys.modules[_name]
else:
    _spec = importlib.util.spec_from_file_location(_name, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_name] = _mod
    _spec.loader.exec_module(_mod)
sys.modules[__name__] = _mod
