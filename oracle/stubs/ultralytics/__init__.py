"""Deterministic stand-in for ultralytics.YOLO (third-party detector, weights unreachable offline).
Returns 8 evenly spaced full-height boxes per image.  TEST INFRASTRUCTURE ONLY."""
import cv2
import torch


class _Boxes:
    def __init__(self, xyxy):
        self.xyxy = xyxy


class _Result:
    def __init__(self, xyxy):
        self.boxes = _Boxes(xyxy)


class YOLO:
    N_BOXES = 8

    def __init__(self, path=None):
        self.path = path

    def __call__(self, paths, imgsz=640, iou=0.1, conf=0.07):
        out = []
        for p in paths:
            img = cv2.imread(p)
            h, w = img.shape[:2]
            n = self.N_BOXES
            step = w / n
            xyxy = torch.tensor([[i * step + 0.1 * step, 0, (i + 1) * step - 0.1 * step, h] for i in range(n)])
            out.append(_Result(xyxy))
        return out
