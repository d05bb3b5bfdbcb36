"""ResNet-45 feature extractor of the LR encoder (reference models/resnet.py:11-74), as a parameter
container + a fixed sequence of NHWC implicit-GEMM convolutions with fused ReLU / residual epilogues.

Geometry: conv3x3(3->32) then blocks [3,4,6,6,3] at widths [32,64,128,256,512] with first-block
strides [(2,1),1,(2,1),1,1]; each block is 1x1 -> ReLU -> 3x3(stride) -> (+ 1x1 stride projection of
the input on the first block of a stage) -> ReLU; no BatchNorm, no biases.
"""
import math

import torch.nn as nn

from .. import ops
from ..ops import ACT_RELU

_STAGES = ((32, 3, (2, 1)), (64, 4, (1, 1)), (128, 6, (2, 1)), (256, 6, (1, 1)), (512, 3, (1, 1)))


def _pack(w, name=None):
    cout, cin, kh, kw = w.shape
    return ops.ConvWeight(w.detach().permute(2, 3, 1, 0).reshape(kh * kw * cin, cout).contiguous(), kh * kw, name=name)


def conv1x1(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=1, stride=stride, bias=False)


def conv3x3(in_planes, out_planes, stride=1):
    return nn.Conv2d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=False)


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = conv1x1(inplanes, planes)
        self.conv2 = conv3x3(planes, planes, stride)
        self.downsample = downsample
        self.stride = stride if isinstance(stride, tuple) else (stride, stride)

    def pack(self, name=None):
        nm = (lambda s: None) if name is None else (lambda s: f"{name}.{s}")
        return dict(c1=_pack(self.conv1.weight, nm("conv1")), c2=_pack(self.conv2.weight, nm("conv2")), stride=self.stride,
                    ds=None if self.downsample is None else _pack(self.downsample[0].weight, nm("downsample.0")))


class ResNet(nn.Module):
    def __init__(self, stages=_STAGES):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 32, kernel_size=3, stride=1, padding=1, bias=False)
        inplanes = 32
        for li, (planes, nblk, stride) in enumerate(stages, 1):
            blocks = []
            for bi in range(nblk):
                s = stride if bi == 0 else (1, 1)
                ds = None
                if bi == 0 and (s != (1, 1) or inplanes != planes):
                    ds = nn.Sequential(nn.Conv2d(inplanes, planes, kernel_size=1, stride=s, bias=False))
                blocks.append(BasicBlock(inplanes, planes, s, ds))
                inplanes = planes
            setattr(self, f"layer{li}", nn.Sequential(*blocks))
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                n = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
                m.weight.data.normal_(0, math.sqrt(2.0 / n))

    def pack(self, name="resnet"):
        blocks = []
        for li in range(1, 6):
            blocks += [b.pack(f"{name}.layer{li}.{bi}") for bi, b in enumerate(getattr(self, f"layer{li}"))]
        return dict(stem=_pack(self.conv1.weight, f"{name}.conv1"), blocks=blocks)

    @staticmethod
    def run(pk, x):
        """x: NHWC [B,32,512,3] -> NHWC [B,8,512,512]."""
        x = ops.conv2d(x, pk["stem"], 3, 3, pad=(1, 1), act=ACT_RELU)
        for b in pk["blocks"]:
            o = ops.conv2d(x, b["c1"], 1, 1, act=ACT_RELU)
            r = x if b["ds"] is None else ops.conv2d(x, b["ds"], 1, 1, stride=b["stride"])
            x = ops.conv2d(o, b["c2"], 3, 3, stride=b["stride"], pad=(1, 1), residual=r, act=ACT_RELU)
        return x

    def forward(self, x):
        """NCHW in / NCHW-shaped (channels_last) out, like the reference module."""
        if not x.is_cuda:
            raise RuntimeError("marconet_b200 ResNet runs only on CUDA (sm_100a) devices")
        return ops.as_nchw_view(self.run(self.pack(), ops.nchw_to_nhwc(x.float())))


def resnet45stride():
    return ResNet()
