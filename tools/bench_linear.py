"""Steady-state cost of the small-M linear layers: LOOPS launches recorded into one CUDA graph (inputs L2-resident, programmatic
dependent launch overlapping the launch latency, no host launch floor -- like inside the module graphs), CUDA events around the replay.
MN_LIN_KS=1|2|4|8 forces the number of K slices per cluster.  Usage: python tools/bench_linear.py"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
LOOPS = 100
for m, k, n in ((16, 512, 512), (64, 512, 512), (64, 512, 1536), (64, 512, 1024), (64, 1024, 512), (64, 512, 6736), (16, 512, 7168), (16, 512, 1536)):
    x, w, b = torch.randn(m, k, device=dev), torch.randn(k, n, device=dev), torch.randn(n, device=dev)
    y = torch.empty(m, n, device=dev)
    for _ in range(5):
        ops.linear(x, w, b, out=y)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(LOOPS):
            ops.linear(x, w, b, out=y)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"ks": os.environ.get("MN_LIN_KS", "auto"), "shape": [m, k, n], "us_per_launch": round(e0.elapsed_time(e1) / LOOPS * 1e3, 2)}))
