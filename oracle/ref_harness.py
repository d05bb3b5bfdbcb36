"""Import the UNMODIFIED reference modules from /root/reference (TEST INFRASTRUCTURE ONLY).

Works only in the build container (the reference tree is not shipped to the GPU
box).  The reference's one third-party import, ``basicsr.ops.fused_act``
(models/networks.py:10), and the script-level third-party packages
(ultralytics, modelscope, imageio: test_sr.py:13-15, test_w.py:9) are absent
here; ``oracle/stubs`` provides stand-ins that are put first on sys.path.
"""
import importlib
import os
import sys

REFERENCE_ROOT = os.environ.get("MARCONET_REFERENCE", "/root/reference")
_STUBS = os.path.join(os.path.dirname(os.path.abspath(__file__)), "stubs")


def available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "networks.py"))


def import_reference_networks():
    """Returns the reference ``models.networks`` module (namespace package rooted at REFERENCE_ROOT)."""
    if not available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    for name in [m for m in sys.modules if m == "models" or m.startswith("models.")]:
        mod = sys.modules[name]
        f = getattr(mod, "__file__", None) or ""
        if not f.startswith(REFERENCE_ROOT):
            del sys.modules[name]
    for p in (_STUBS, REFERENCE_ROOT):
        if p in sys.path:
            sys.path.remove(p)
    sys.path.insert(0, REFERENCE_ROOT)
    sys.path.insert(0, _STUBS)
    return importlib.import_module("models.networks")


def build_reference_models(sds):
    """Construct the three reference classes with no args and strict-load the synthetic checkpoints
    (test_sr.py:42-52)."""
    nets = import_reference_networks()
    out = {}
    for key, cls in (("tspgan", nets.TSPGAN), ("encoder", nets.TextContextEncoderV2), ("sr", nets.TSPSRNet)):
        m = cls()
        m.load_state_dict(sds[key], strict=True)
        out[key] = m.eval()
    return out
