"""Small-shape exercise of the kernels added late in round 1, meant to run under compute-sanitizer:
   compute-sanitizer --tool memcheck python tools/sanitize_new_kernels.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
flag = torch.zeros(1, dtype=torch.int32, device=dev)
# device-side checks
ops.check_labels(torch.tensor([-1, 5, 7000], device=dev), 6736, flag)
first = torch.tensor([0, 3, 3, 5], dtype=torch.int32, device=dev)
ops.char_windows(torch.rand(3, 8, device=dev), first, [3, 0, 2], 512, 16, flag)
ops.char_windows(torch.rand(3, 8, device=dev) * 1.4, first, [3, 0, 2], 1024, 32, flag)
# image ops
for h, w in ((9, 33), (48, 300), (32, 512), (131, 97)):
    img = torch.from_numpy(np.random.default_rng(h).integers(0, 256, (h, w, 3), dtype=np.uint8)).to(dev)
    ops.preprocess_lq(img, return_resized=True)
sr = torch.rand(2, 3, 16, 40, device=dev) * 2 - 1
ops.postprocess_sr(sr)
ops.postprocess_sr(sr.contiguous(memory_format=torch.channels_last))
# small-M linear: all tile shapes, cluster K slices, ragged N tile, gathered patch embedding
for m, k, n in ((16, 512, 512), (33, 64, 16), (64, 1024, 512), (17, 4096, 1040), (16, 512, 7168), (2, 512, 6736)):
    x, w, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev), torch.randn(n, device=dev)
    ops.linear(x, w, b, act=ops.ACT_GELU, residual=torch.randn(m, n, device=dev))
feat = torch.randn(2, 8, 512, 512, device=dev)
ops.patch_embed(feat, torch.randn(32768, 512, device=dev), torch.randn(512, device=dev), torch.randn(64, 512, device=dev))
# bilinear x2 (sliding window) incl. odd widths and 1-pixel maps, GroupNorm apply with ragged widths
for n, h, w, c in ((2, 4, 4, 64), (1, 1, 1, 32), (3, 5, 7, 128), (1, 8, 130, 64)):
    x = torch.randn(n, h, w, c, device=dev)
    ops.resample_modulate(x, torch.randn(n, c, device=dev), up=True)
    ops.resample_modulate(x, None, up=True)
x = torch.randn(3, 6, 10, 64, device=dev)
vw = torch.tensor([10, 4, 0], dtype=torch.int32, device=dev)
mr = ops.groupnorm_stats(x, valid_w=vw)
ops.groupnorm_apply(x, mr, torch.randn(64, device=dev), torch.randn(64, device=dev), valid_w=vw)
torch.cuda.synchronize()
print("sanitize_new_kernels: done")
