"""Micro-benchmark of one res-block convolution with the GroupNorm+swish operand transform, fused (default) vs two-pass
(mn_groupnorm_apply + conv): CUDA events, L2 flushed between launches.  Developer tool.
    python tools/bench_conv_gn.py N H W Cin Cout [iters]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import ops  # noqa: E402


def main():
    n, h, w, cin, cout = [int(v) for v in sys.argv[1:6]]
    iters = int(sys.argv[6]) if len(sys.argv) > 6 else 20
    dev = torch.device("cuda:0")
    x = torch.randn(n, h, w, cin, device=dev)
    wt = ops.ConvWeight((torch.randn(9 * cin, cout, device=dev) / (9 * cin) ** 0.5).contiguous(), 9)
    bias, gamma, beta = torch.randn(cout, device=dev), torch.rand(cin, device=dev) + 0.5, torch.randn(cin, device=dev)
    mr = ops.groupnorm_stats(x)
    flush = torch.empty(64 << 20, dtype=torch.float32, device=dev)
    fl = 2.0 * n * h * w * cout * 9 * cin
    for fuse in (True, False):
        tot = 0.0
        for i in range(iters + 3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv2d(x, wt, 3, 3, pad=(1, 1), bias=bias, gn=(mr, gamma, beta), gn_fuse=fuse)
            e1.record()
            torch.cuda.synchronize()
            if i >= 3:
                tot += e0.elapsed_time(e1)
        ms = tot / iters
        print(f"conv+GN N{n} {h}x{w} {cin}->{cout} k3 {'fused transform' if fuse else 'two passes   '}: {ms * 1e3:.1f} us  {fl / ms / 1e9:.1f} TFLOP/s (algorithmic)")


if __name__ == "__main__":
    main()
