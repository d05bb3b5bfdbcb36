"""ResNet-45 feature extractor of the LR encoder (reference models/resnet.py:11-74), as a parameter
container + a fixed sequence of NHWC implicit-GEMM convolutions with fused ReLU / residual epilogues.

Geometry: conv3x3(3->32) then blocks [3,4,6,6,3] at widths [32,64,128,256,512] with first-block
strides [(2,1),1,(2,1),1,1]; each block is 1x1 -> ReLU -> 3x3(stride) -> (+ 1x1 stride projection of
the input on the first block of a stage) -> ReLU; no BatchNorm, no biases.
"""
import math

import torch
import torch.nn as nn

from .. import ops
from ..ops import ACT_RELU

_STAGES = ((32, 3, (2, 1)), (64, 4, (1, 1)), (128, 6, (2, 1)), (256, 6, (1, 1)), (512, 3, (1, 1)))


_PADC = 64     # the tcgen05 conv kernels work on 64-channel blocks


def _pack(w, pad_in=False, pad_out=False):
    """[Cout, Cin, KH, KW] -> ops.ConvWeight.  ``pad_out`` / ``pad_in`` zero-pad a 32-channel side to 64 so that the 32-wide
    stage of the net runs on the tensor-core kernels: its activations are kept as 64-channel NHWC buffers whose upper half is
    exactly zero (zero weight columns write zeros, relu(0 + 0) = 0, zero weight rows ignore them)."""
    w = w.detach()
    cout, cin, kh, kw = w.shape
    if pad_out and cout < _PADC:
        w = torch.cat([w, w.new_zeros(_PADC - cout, cin, kh, kw)], dim=0)
        cout = _PADC
    if pad_in and cin < _PADC:
        w = torch.cat([w, w.new_zeros(cout, _PADC - cin, kh, kw)], dim=1)
        cin = _PADC
    return ops.ConvWeight(w.permute(2, 3, 1, 0).reshape(kh * kw * cin, cout).contiguous(), kh * kw)


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

    def pack(self):
        # stride-1 convs take the padded 64-channel buffer as it is (zero weight rows for the padding); strided convs run on
        # the fp32 kernel and read only the real channels through a channel-slice view, so their K is not padded.
        cin = self.conv1.in_channels
        unit = self.stride == (1, 1)
        return dict(c1=_pack(self.conv1.weight, pad_in=True, pad_out=True), c2=_pack(self.conv2.weight, pad_in=unit, pad_out=True),
                    c2_in=self.conv2.in_channels, stride=self.stride, cin=cin,
                    ds=None if self.downsample is None else _pack(self.downsample[0].weight, pad_in=unit, pad_out=True))


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

    def pack(self):
        blocks = []
        for li in range(1, 6):
            blocks += [b.pack() for b in getattr(self, f"layer{li}")]
        return dict(stem=_pack(self.conv1.weight, pad_out=True), blocks=blocks)

    @staticmethod
    def run(pk, x):
        """x: NHWC [B,32,512,3] -> NHWC [B,8,512,512].  Activations of the 32-wide stage are 64-channel buffers with a zero
        upper half (see _pack)."""
        def real(t, c):          # the first c channels as a channel-slice view (x_cs stays 64): input of a strided conv
            return t if t.shape[-1] == c else t[..., :c]

        x = ops.conv2d(x, pk["stem"], 3, 3, pad=(1, 1), act=ACT_RELU)
        for b in pk["blocks"]:
            unit = b["stride"] == (1, 1)
            o = ops.conv2d(x, b["c1"], 1, 1, act=ACT_RELU)
            r = x if b["ds"] is None else ops.conv2d(x if unit else real(x, b["cin"]), b["ds"], 1, 1, stride=b["stride"])
            x = ops.conv2d(o if unit else real(o, b["c2_in"]), b["c2"], 3, 3, stride=b["stride"], pad=(1, 1), residual=r, act=ACT_RELU)
        return x

    def forward(self, x):
        """NCHW in / NCHW-shaped (channels_last) out, like the reference module."""
        if not x.is_cuda:
            raise RuntimeError("marconet_b200 ResNet runs only on CUDA (sm_100a) devices")
        return ops.as_nchw_view(self.run(self.pack(), ops.nchw_to_nhwc(x.float())))


def resnet45stride():
    return ResNet()
