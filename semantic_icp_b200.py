"""Import shim: the package directory is named `semantic-icp_b200/` (not a valid Python identifier), so this module
loads its Python parts and re-exports them as `semantic_icp_b200.sicp` / `semantic_icp_b200.synth`."""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PY = os.path.join(_ROOT, "semantic-icp_b200", "python")


def _load(name):
    full = f"semantic_icp_b200.{name}"
    if full in sys.modules:
        return sys.modules[full]
    spec = importlib.util.spec_from_file_location(full, os.path.join(_PY, f"{name}.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[full] = mod
    spec.loader.exec_module(mod)
    return mod


sicp = _load("sicp")
synth = _load("synth")
shard = _load("shard")
PKG_ROOT = os.path.join(_ROOT, "semantic-icp_b200")
