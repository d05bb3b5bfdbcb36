"""Small-shape exercise of the kernels added or restructured in round 2, meant to run under compute-sanitizer:
   compute-sanitizer --tool memcheck  python tools/sanitize_round2_kernels.py
   compute-sanitizer --tool racecheck python tools/sanitize_round2_kernels.py
conv_tc2 with cta_group::2 pairs / dedicated epilogue warps / epilogue GroupNorm statistics / per-sample output pointers / range
guard, query-chunked attention, outer-K patch embedding, the standalone helper functions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from marconet_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
# tcgen05 conv: pair kernel (cs = 2), single-CTA (cs = 1), 64- and 128-wide tiles, masked windows, second output, GN statistics
for n, h, w, cin, cout in ((3, 32, 32, 64, 256), (1, 8, 16, 64, 64), (2, 16, 32, 128, 128), (5, 4, 4, 128, 128), (1, 64, 64, 64, 128)):
    x = torch.randn(n, h, w, cin, device=dev)
    cw = ops.ConvWeight((torch.randn(9 * cin, cout, device=dev) / (9 * cin) ** 0.5).contiguous(), 9, name=f"san.{n}.{h}.{cin}.{cout}")
    bias = torch.randn(cout, device=dev)
    vw = torch.randint(1, w + 1, (n,), dtype=torch.int32, device=dev)
    for prec in (ops.PREC_F16X3_TC, ops.PREC_BF16X3_TC):
        ops.conv2d(x, cw, 3, 3, pad=(1, 1), bias=bias, act=ops.ACT_LRELU02, gain=2 ** 0.5, valid_w=vw, precision=prec)
        ops.conv2d(x, cw, 3, 3, pad=(1, 1), bias=bias, out_scale=torch.rand(n, cout, device=dev), out2=True, y2_scale=torch.rand(n, cout, device=dev),
                   residual=torch.randn(n, h, w, cout, device=dev), act=ops.ACT_TANH, precision=prec)
    if h * w >= 128 and 2.0 * n * h * w * cout * 9 * cin >= ops.TC_MIN_FLOP:     # layers the tensor-core kernel runs (peer pointers need it)
        y1, mr = ops.conv2d(x, cw, 3, 3, pad=(1, 1), bias=bias, valid_w=vw, gn_stats=True)
        # fused GroupNorm+swish operand transform (four lanes per halo row, in-place overwrite ordered by __syncwarp) on that output
        cw2 = ops.ConvWeight((torch.randn(9 * cout, cout, device=dev) / (9 * cout) ** 0.5).contiguous(), 9, name=f"san.gn.{n}.{h}.{cout}")
        ops.conv2d(y1, cw2, 3, 3, pad=(1, 1), bias=bias, valid_w=vw, gn=(mr, torch.rand(cout, device=dev) + 0.5, torch.randn(cout, device=dev)), gn_fuse=True)
        # per-sample destination pointers of the second output (local memory standing in for a peer's)
        dst = torch.empty(n, h, w, cout, device=dev)
        ptrs = torch.tensor([dst[i].data_ptr() for i in reversed(range(n))], dtype=torch.int64, device=dev)
        ops.conv2d(x, cw, 3, 3, pad=(1, 1), bias=bias, out2_ptrs=ptrs)
# fp32 CUDA-core implicit GEMM (epilogue staged through shared memory, one rolled loop): vector and scalar epilogues, strides,
# split-K partials + reduce, residual / second output
for n, h, w, cin, cout, k, st in ((1, 16, 64, 32, 32, 1, (1, 1)), (1, 16, 64, 64, 128, 1, (2, 1)), (2, 9, 13, 3, 5, 3, (1, 1)),
                                  (1, 16, 32, 128, 128, 3, (2, 2)), (16, 4, 4, 64, 64, 3, (1, 1)), (1, 7, 11, 16, 130, 3, (1, 1))):
    x = torch.randn(n, h, w, cin, device=dev)
    wt = (torch.randn(k * k * cin, cout, device=dev) / (k * k * cin) ** 0.5).contiguous()
    y = ops.conv2d(x, wt, k, k, stride=st, pad=(k // 2, k // 2), bias=torch.randn(cout, device=dev), act=ops.ACT_RELU, precision=ops.PREC_FP32_SIMT)
    ops.conv2d(x, wt, k, k, stride=st, pad=(k // 2, k // 2), residual=torch.randn_like(y), out2=True, act=ops.ACT_TANH,
               precision=ops.PREC_FP32_SIMT, split_k=3 if k == 3 else 0)
# calibration / range guard path
with ops.calibration(dev) as cal:
    x = torch.randn(2, 16, 16, 64, device=dev) * 1e5
    cw = ops.ConvWeight((torch.randn(9 * 64, 64, device=dev) / 24).contiguous(), 9, name="san.range")
    ops.conv2d(x, cw, 3, 3, pad=(1, 1))
cal.results()
ops.poll_range(dev, reroute=False)
# attention (S = 64 and 16), patch embedding with outer K slices, helper functions
for b, s in ((1, 64), (3, 16), (2, 37)):
    ops.attention(torch.randn(b, s, 1536, device=dev))
feat = torch.randn(1, 8, 512, 512, device=dev)
ops.patch_embed(feat, torch.randn(32768, 512, device=dev), torch.randn(512, device=dev), torch.randn(64, 512, device=dev))
a, b2 = torch.randn(2, 7, 5, 9, device=dev), torch.randn(2, 7, 4, 6, device=dev)
ops.swish(a); ops.calc_mean_std_4d(a); ops.adaptive_instance_normalization(a, b2)
torch.cuda.synchronize()
print("sanitize_round2_kernels: done")
