"""Generates tests/golden/script_sr_row.npz: the SR row of the PNG that the reference's UNMODIFIED test_sr.py writes for one
synthetic LR image, using the reference's own models on the CPU, the real cv2 / torchvision pre- and post-processing, the
synthetic checkpoints (marconet_b200.testing.synth, seed 0) and the deterministic detector / OCR stand-ins of oracle/stubs.

OpenCV runs with OPENCV_IPP=disabled so that cv2.resize takes OpenCV's own cubic code path (see oracle/image_ops.py: the IPP
path is closed source, CPU dependent, and differs by +-1 grey level on ~5 % of the LQ pixels).

Build container only (needs /root/reference):  python -m oracle.make_golden_script
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
OUT = os.path.join(ROOT, "tests", "golden", "script_sr_row.npz")
STRIDE = 2      # the stored SR row is subsampled [::STRIDE, ::STRIDE] to keep the fixture small


def main():
    import cv2
    sys.path.insert(0, ROOT)
    from marconet_b200.testing import synth
    sds = synth.make_checkpoints(0)
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "checkpoints"))
        for key, name in (("tspgan", "net_prior_generation.pth"), ("sr", "net_sr.pth"), ("encoder", "net_transformer_encoder.pth")):
            torch.save({"params": sds[key]}, os.path.join(d, "checkpoints", name))
        os.makedirs(os.path.join(d, "LQs"))
        img = np.random.default_rng(0).integers(0, 256, (32, 256, 3), dtype=np.uint8)
        path = os.path.join(d, "LQs", "line0.png")
        cv2.imwrite(path, img)
        env = dict(os.environ, PYTHONPATH=STUBS, OPENCV_IPP="disabled", OMP_NUM_THREADS=str(os.cpu_count() or 1))
        r = subprocess.run([sys.executable, os.path.join(REF, "test_sr.py"), "-i", "./LQs", "-o", "./out"], cwd=d, env=env,
                           capture_output=True, text=True, timeout=1800)
        assert r.returncode == 0, r.stderr[-3000:]
        outs = os.listdir(os.path.join(d, "out"))
        assert len(outs) == 1, outs
        png = cv2.imread(os.path.join(d, "out", outs[0]))
        # detector / OCR stand-ins, through the reference's own helper, to record what the script fed the nets
        sys.path[:0] = [STUBS, REF]
        from ultralytics import YOLO
        from modelscope.pipelines import pipeline
        from utils.yolo_ocr_xloc import get_yolo_ocr_xloc
        from utils.alphabets import alphabet
        rgb, boxes, chars, _ = get_yolo_ocr_xloc(path, yolo_model=YOLO(None), ocr_pipeline=pipeline(None), num_cropped_boxes=5, expand_px=1,
                                                 expand_px_for_first_last_cha=12, yolo_iou=0.1, yolo_conf=0.07)
        labels = [alphabet.find(c) for c in chars]
        assert min(labels) >= 0
    assert png.shape[0] == 4 * 128, png.shape
    sr_row = png[256:384]                                    # rows: ShowLQ, ShowLocs, ShowSR, prior (test_sr.py:231)
    np.savez_compressed(OUT, image_rgb=np.ascontiguousarray(rgb), boxes=np.asarray(boxes, dtype=np.int64), labels=np.asarray(labels, dtype=np.int64),
                        sr_row=np.ascontiguousarray(sr_row[::STRIDE, ::STRIDE]), stride=np.array(STRIDE), text=np.array("".join(chars)))
    print("wrote", OUT, sr_row.shape, "chars", "".join(chars), "boxes", np.asarray(boxes).tolist()[:2], os.path.getsize(OUT))


if __name__ == "__main__":
    main()
