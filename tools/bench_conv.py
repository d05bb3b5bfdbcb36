"""Micro-benchmark of one mn_conv2d_nhwc problem (CUDA events, L2 flushed between launches).
    python tools/bench_conv.py N H W Cin Cout K [precision] [iters]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import ops  # noqa: E402


def main():
    n, h, w, cin, cout, k = [int(v) for v in sys.argv[1:7]]
    prec = int(sys.argv[7]) if len(sys.argv) > 7 else ops.PREC_F16X3_TC
    iters = int(sys.argv[8]) if len(sys.argv) > 8 else 20
    dev = torch.device("cuda:0")
    x = torch.randn(n, h, w, cin, device=dev)
    wt = ops.ConvWeight((torch.randn(k * k * cin, cout, device=dev) / (k * k * cin) ** 0.5).contiguous(), k * k)
    bias = torch.randn(cout, device=dev)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    tot = 0.0
    for i in range(iters + 3):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.conv2d(x, wt, k, k, pad=(k // 2, k // 2), bias=bias, act=ops.ACT_LRELU02, gain=2 ** 0.5, precision=prec)
        e1.record()
        torch.cuda.synchronize()
        if i >= 3:
            tot += e0.elapsed_time(e1)
    ms = tot / iters
    fl = 2.0 * n * h * w * cout * k * k * cin
    print(f"conv N{n} {h}x{w} {cin}->{cout} k{k} prec{prec}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s (algorithmic)")


if __name__ == "__main__":
    main()
