"""The reference's UNMODIFIED test_sr.py, executed (a) against the reference's own models on CPU (BASELINE configs[0]:
plumbing, proves the harness and stubs) and (b) against this repo's drop-in `models` package.

Needs the read-only reference tree, so it runs in the build container only (skipped on the GPU box, where the same call
sequence is exercised by tests/test_gpu_models.py::test_full_pipeline_vs_golden).  Without a GPU, (b) must get through
construction, strict checkpoint load, .eval(), .to(device) and the parameter banner, and then fail LOUDLY at the first
forward (no CPU fallback)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MARCONET_REFERENCE", "/root/reference")
STUBS = os.path.join(ROOT, "oracle", "stubs")

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def workdir(tmp_path_factory, checkpoints):
    if not os.path.isfile(os.path.join(REF, "test_sr.py")):
        pytest.skip("reference tree not present")
    import cv2
    d = tmp_path_factory.mktemp("marconet_run")
    os.makedirs(d / "checkpoints")
    for key, name in (("tspgan", "net_prior_generation.pth"), ("sr", "net_sr.pth"), ("encoder", "net_transformer_encoder.pth")):
        torch.save({"params": checkpoints[key]}, d / "checkpoints" / name)
    os.makedirs(d / "LQs")
    img = np.random.default_rng(0).integers(0, 256, (32, 256, 3), dtype=np.uint8)
    cv2.imwrite(str(d / "LQs" / "line0.png"), img)
    # A user swaps the reference's models/ directory for dropin/models.  Python puts the script's own directory first on
    # sys.path, so to emulate that checkout the byte-identical script is copied (at test time, into the temp dir only) next
    # to a symlink of the reference's utils/ -- the reference's models/ is then NOT importable from this layout.
    import shutil
    os.makedirs(d / "swapped")
    shutil.copy(os.path.join(REF, "test_sr.py"), d / "swapped" / "test_sr.py")
    os.symlink(os.path.join(REF, "utils"), d / "swapped" / "utils")
    return d


def _run(workdir, script, pythonpath, out):
    env = dict(os.environ, PYTHONPATH=os.pathsep.join(pythonpath), OMP_NUM_THREADS=str(os.cpu_count() or 1))
    return subprocess.run([sys.executable, script, "-i", "./LQs", "-o", out], cwd=workdir, env=env,
                          capture_output=True, text=True, timeout=900)


def test_reference_script_with_reference_models_cpu(workdir):
    r = _run(workdir, os.path.join(REF, "test_sr.py"), [STUBS], "./out_ref")
    assert r.returncode == 0, r.stderr[-2000:]
    assert "43.062275 M Parameters" in r.stdout and "27.970194 M Parameters" in r.stdout and "16.865923 M Parameters" in r.stdout
    assert len(os.listdir(workdir / "out_ref")) == 1


def test_reference_script_with_dropin_models(workdir):
    import filecmp
    assert filecmp.cmp(workdir / "swapped" / "test_sr.py", os.path.join(REF, "test_sr.py"), shallow=False)
    r = _run(workdir, str(workdir / "swapped" / "test_sr.py"), [os.path.join(ROOT, "dropin"), ROOT, STUBS], "./out_b200")
    # identical banner = identical parameter sets, after construction + strict load from ./checkpoints/*.pth
    assert "43.062275 M Parameters" in r.stdout and "27.970194 M Parameters" in r.stdout and "16.865923 M Parameters" in r.stdout, \
        r.stdout[-1500:] + r.stderr[-1500:]
    if torch.cuda.is_available():
        assert r.returncode == 0, r.stderr[-2000:]
        assert len(os.listdir(workdir / "out_b200")) == 1
    else:
        assert r.returncode != 0 and "no CPU fallback" in r.stderr, r.stderr[-1500:]
