"""Mirror of the reference's `models/module.py` surface (jdacs/models/module.py), B200 path underneath.

Same names and call signatures: ConvBnReLU, ConvBn, ConvBnReLU3D, ConvBn3D (parameter holders with the
reference's state-dict keys), homo_warping, depth_regression.  The 2-D blocks stay library (cuDNN) code —
they belong to FeatureNet, which is outside the plane-sweep path; everything 3-D / warp / regression runs in
libmvs_b200.  (The reference's unused BasicBlock / Hourglass3d are not mirrored.)
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops, regnet

ALIGN_CORNERS = False  # hazard H1: what F.grid_sample does on torch >= 1.3 at the reference call site


class ConvBnReLU(nn.Module):
    """jdacs/models/module.py:15-22 (2-D, library code)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        return F.relu(self.bn(self.conv(x)), inplace=True)


class ConvBn(nn.Module):
    """jdacs/models/module.py:25-32 (2-D, library code)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv2d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm2d(out_channels)

    def forward(self, x):
        return self.bn(self.conv(x))


class ConvBnReLU3D(nn.Module):
    """jdacs/models/module.py:35-42.  Holder of conv.weight / bn.*; forward runs mvs_conv3d_fwd.

    Accepts a C8 volume (internal use) or the reference's [B,C,D,H,W] tensor; returns the same kind."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        if kernel_size != 3 or pad != 1:
            raise ValueError("the plane-sweep path only has 3x3x3, pad-1 convolutions")
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)
        self._cache = regnet.PackCache()

    def train(self, mode=True):
        self._cache.clear()     # everything derived from the weights is re-packed on the next eval forward
        return super().train(mode)

    def forward(self, x, skip=None, algo=0, frozen_grad=False):
        plain = x.dim() == 5
        y = regnet.conv_bn_relu(regnet.as_c8(x, torch.float32), self.conv, self.bn, self.training, self._cache, skip, algo, frozen_grad)
        return regnet.unpack_c8_grad(y) if plain else y


class ConvBn3D(nn.Module):
    """jdacs/models/module.py:45-52 (parameter holder; unused by MVSNet itself, kept for import parity)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, pad=1):
        super().__init__()
        self.conv = nn.Conv3d(in_channels, out_channels, kernel_size, stride=stride, padding=pad, bias=False)
        self.bn = nn.BatchNorm3d(out_channels)

    def forward(self, x):
        return self.bn(self.conv(x))


def homo_warping(src_fea, src_proj, ref_proj, depth_values):
    """jdacs/models/module.py:105-140: [B,C,H,W] x [B,4,4] x [B,4,4] x [B,D] -> [B,C,D,H,W].

    Gradient reaches src_fea only (the reference builds the grid under no_grad)."""
    rt = ops.compose_proj(torch.stack((ref_proj, src_proj), dim=1))
    return ops.homo_warp(src_fea, rt[0], depth_values, ALIGN_CORNERS)


def depth_regression(p, depth_values):
    """jdacs/models/module.py:145-148: expectation of depth_values under the probability volume p [B,D,H,W].

    p is already normalised here (API parity); the fused softmax + regression the model uses is ops.soft_argmin."""
    depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return torch.sum(p * depth_values, 1)
