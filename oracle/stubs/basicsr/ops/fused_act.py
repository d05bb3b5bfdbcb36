"""Stand-in for basicsr.ops.fused_act (third-party CUDA extension, not installed here).
Published formula: y = leaky_relu(x + bias, 0.2) * sqrt(2). TEST INFRASTRUCTURE ONLY."""
import torch
import torch.nn as nn
import torch.nn.functional as F


def fused_leaky_relu(input, bias=None, negative_slope=0.2, scale=2 ** 0.5):
    if bias is not None:
        input = input + bias.view(1, -1, *([1] * (input.dim() - 2)))
    return F.leaky_relu(input, negative_slope) * scale


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale

    def forward(self, input):
        return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)
