"""Generates tests/golden/script_w.npz: samples of the 11 PNGs that the reference's UNMODIFIED test_w.py writes for two synthetic
LR images, using the reference's own models on the CPU, real cv2 / torchvision, the synthetic checkpoints (seed 0) and the
imageio stand-in of oracle/stubs (test_w.py:9,115 only writes the GIF with it).  OPENCV_IPP=disabled as in make_golden_script.py.

Build container only (needs /root/reference):  python -m oracle.make_golden_script_w
TEST INFRASTRUCTURE ONLY.
"""
import os
import subprocess
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MARCONET_REFERENCE", "/root/reference")
STUBS = os.path.join(ROOT, "oracle", "stubs")
OUT = os.path.join(ROOT, "tests", "golden", "script_w.npz")
SY, SX = 8, 16       # stored samples: png[::SY, ::SX]


def make_inputs(d):
    """The two style images of the test (shared with tests/test_dropin_scripts.py)."""
    import cv2
    os.makedirs(d, exist_ok=True)
    w1 = np.random.default_rng(1).integers(0, 256, (48, 400, 3), dtype=np.uint8)     # resized to 32 x 267
    w2 = np.random.default_rng(2).integers(0, 256, (32, 255, 3), dtype=np.uint8)
    cv2.imwrite(os.path.join(d, "w1.png"), w1)
    cv2.imwrite(os.path.join(d, "w2.png"), w2)


def write_checkpoints(d, sds, names=("tspgan", "sr", "encoder")):
    os.makedirs(os.path.join(d, "checkpoints"), exist_ok=True)
    files = dict(tspgan="net_prior_generation.pth", sr="net_sr.pth", encoder="net_transformer_encoder.pth")
    for key in names:
        torch.save({"params": sds[key]}, os.path.join(d, "checkpoints", files[key]))


def main():
    import cv2
    sys.path.insert(0, ROOT)
    from marconet_b200.testing import synth
    sds = synth.make_checkpoints(0)
    with tempfile.TemporaryDirectory() as d:
        write_checkpoints(d, sds, ("tspgan", "encoder"))
        make_inputs(os.path.join(d, "in"))
        env = dict(os.environ, PYTHONPATH=STUBS, OPENCV_IPP="disabled", OMP_NUM_THREADS=str(os.cpu_count() or 1))
        r = subprocess.run([sys.executable, os.path.join(REF, "test_w.py"), "-w1", "./in/w1.png", "-w2", "./in/w2.png", "-o", "./out"],
                           cwd=d, env=env, capture_output=True, text=True, timeout=3600)
        assert r.returncode == 0, r.stderr[-3000:]
        assert "Finishing interpolation." in r.stdout
        rec = {}
        for i in range(11):
            name = "w_{:.2f}.png".format(i / 10)
            png = cv2.imread(os.path.join(d, "out", name))
            assert png is not None and png.shape[0] == 128 and png.shape[1] % 128 == 0, name
            rec[f"png{i}"] = np.ascontiguousarray(png[::SY, ::SX])
            rec["width"] = np.array(png.shape[1])
        assert os.path.isfile(os.path.join(d, "out", "w.gif"))
    rec["sy"], rec["sx"] = np.array(SY), np.array(SX)
    np.savez_compressed(OUT, **rec)
    print("wrote", OUT, "chars", int(rec["width"]) // 128, os.path.getsize(OUT))


if __name__ == "__main__":
    main()
