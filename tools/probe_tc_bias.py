"""Measure the systematic (truncation) bias of the tcgen05 fp32 accumulation: signed relative error vs fp64."""
import math, os, sys
import torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import ops

dev = torch.device("cuda:0")
for (n, h, w, cin, cout, k) in [(2, 32, 32, 512, 512, 3), (2, 32, 32, 256, 256, 3), (2, 32, 32, 128, 128, 3), (2, 32, 32, 512, 256, 1)]:
    g = torch.Generator().manual_seed(1)
    x = torch.randn(n, cin, h, w, generator=g) * 1.3 + 0.1
    wt = torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)
    ref = F.conv2d(x.double(), wt.double(), padding=k // 2)
    cw = ops.ConvWeight(wt.permute(2, 3, 1, 0).reshape(k * k * cin, cout).contiguous().to(dev), k * k)
    xn = x.permute(0, 2, 3, 1).contiguous().to(dev)
    for prec, name in ((ops.PREC_F16X3_TC, "f16x3"), (ops.PREC_FP32_SIMT, "fp32")):
        y = ops.conv2d(xn, cw, k, k, pad=(k // 2, k // 2), precision=prec).permute(0, 3, 1, 2).cpu().double()
        err = y - ref
        big = ref.abs() > 0.5
        srel = (err * ref.sign() / ref.abs())[big]
        print(f"K={cin*k*k:5d} {name}: max|err| {err.abs().max():.2e}  mean signed rel err {srel.mean():+.3e}  std {srel.std():.2e} "
              f" pos {((err/ref.abs())[big & (ref>0)]).mean():+.2e} neg {((err/ref.abs())[big & (ref<0)]).mean():+.2e}  n_acc={cin*k*k//16}")
