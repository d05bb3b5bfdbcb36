"""Generates tests/golden/resize_cubic.npz: inputs and outputs of cv2.resize(..., INTER_CUBIC) as the reference's
test_sr.py:98-99 calls it, produced by the real OpenCV (own code path: cv2.ipp.setUseIPP(False), see oracle/image_ops.py).
Run in the build container:  python -m oracle.make_golden_image
"""
import os

import cv2
import numpy as np

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "resize_cubic.npz")
CASES = [(48, 300, 0), (20, 260, 1), (79, 460, 2), (32, 512, 3), (131, 97, 4), (9, 33, 5)]   # (h, w, seed): down/up-scaling, identity, tiny


def make_image(h, w, seed):
    rng = np.random.default_rng(seed)
    if seed % 3 == 2:
        return (rng.integers(0, 2, (h, w, 3)) * 255).astype(np.uint8)     # hard edges: overshoot, saturation, rounding ties
    return rng.integers(0, 256, (h, w, 3), dtype=np.uint8)


def main():
    cv2.ipp.setUseIPP(False)
    data = {"opencv_version": np.array(cv2.__version__)}
    for h, w, seed in CASES:
        img = make_image(h, w, seed)
        data[f"out_{h}x{w}"] = cv2.resize(img, (0, 0), fx=32 / h, fy=32 / h, interpolation=cv2.INTER_CUBIC)
    np.savez_compressed(OUT, **data)
    print("wrote", OUT, {k: v.shape for k, v in data.items()})


if __name__ == "__main__":
    main()
