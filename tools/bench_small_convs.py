"""Micro-benchmark of the small / strided conv layers of one line (the ones the tcgen05 kernel does not take or barely fills):
CUDA events around back-to-back launches (warm L2, as inside the module graphs).  Developer tool.
    python tools/bench_small_convs.py
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import ops  # noqa: E402

SHAPES = [  # name, N, H, W, Cin, Cout, k, (sh, sw)
    ("encoder.conv1", 1, 32, 512, 3, 32, 3, (1, 1)),
    ("layer1.0.conv1", 1, 32, 512, 32, 32, 1, (1, 1)),
    ("layer1.0.conv2", 1, 32, 512, 32, 32, 3, (2, 1)),
    ("layer2.0.conv1", 1, 16, 512, 32, 64, 1, (1, 1)),
    ("layer2.0.conv2", 1, 16, 512, 64, 64, 3, (1, 1)),
    ("layer3.0.conv1", 1, 16, 512, 64, 128, 1, (1, 1)),
    ("layer3.0.ds", 1, 16, 512, 64, 128, 1, (2, 1)),
    ("layer3.0.conv2", 1, 16, 512, 128, 128, 3, (2, 1)),
    ("layer3.1.conv2", 1, 8, 512, 128, 128, 3, (1, 1)),
    ("layer4.1.conv2", 1, 8, 512, 256, 256, 3, (1, 1)),
    ("sr.first_32", 1, 32, 512, 3, 64, 3, (1, 1)),
    ("sr.first_16", 1, 32, 512, 64, 128, 3, (2, 2)),
    ("sr.first_8.0", 1, 16, 256, 128, 256, 3, (2, 2)),
    ("sr.first_8.2", 1, 8, 128, 256, 256, 3, (1, 1)),
    ("sr.body_16.0", 1, 16, 256, 384, 256, 3, (1, 1)),
    ("gen.conv1", 16, 4, 4, 512, 512, 3, (1, 1)),
    ("gen.convs.0", 16, 8, 8, 512, 512, 3, (1, 1)),
    ("gen.convs.2", 16, 16, 16, 512, 512, 3, (1, 1)),
]


def main():
    dev = torch.device("cuda:0")
    iters = 50
    only = sys.argv[1] if len(sys.argv) > 1 else None       # one shape, eager launches (for ncu)
    for name, n, h, w, cin, cout, k, st in SHAPES:
        if only and name != only:
            continue
        x = torch.randn(n, h, w, cin, device=dev)
        wt = ops.ConvWeight((torch.randn(k * k * cin, cout, device=dev) / (k * k * cin) ** 0.5).contiguous(), k * k)
        bias = torch.randn(cout, device=dev)
        l0 = ops.LAUNCHES
        y = ops.conv2d(x, wt, k, k, stride=st, pad=(k // 2, k // 2), bias=bias, act=ops.ACT_RELU)
        nl = ops.LAUNCHES - l0
        for _ in range(5):
            ops.conv2d(x, wt, k, k, stride=st, pad=(k // 2, k // 2), bias=bias, act=ops.ACT_RELU)
        if only:
            torch.cuda.synchronize()
            continue
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(iters):
                ops.conv2d(x, wt, k, k, stride=st, pad=(k // 2, k // 2), bias=bias, act=ops.ACT_RELU)
        g.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        fl = 2.0 * y.numel() * k * k * cin
        mb = (x.numel() + y.numel() + wt.w.numel()) * 4 / 1e6
        print(f"{name:16s} N{n} {h}x{w} {cin}->{cout} k{k} s{st}: {us:7.1f} us  {fl / us * 1e-6:7.2f} TF  {mb / us * 1e3:7.1f} GB/s  launches {nl}")


if __name__ == "__main__":
    main()
