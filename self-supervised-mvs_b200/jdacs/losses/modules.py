"""API-parity shims for `losses/modules.py` (jdacs/losses/modules.py:17-90).

UnSupLoss itself does not use these any more: its SSIM / colour-gradient / smoothness / reconstruction terms live inside the
fused kernels of csrc/loss.cu (ops.unsup_loss).  The names stay importable because `from losses.modules import *` is part of the
reference's surface and its co-segmentation loss (unsup_seg_loss.py, outside the plane-sweep path) still composes them on
feature-sized maps.  Each helper is a short torch expression with the reference's semantics; NHWC maps throughout."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

_C1, _C2 = 0.01 ** 2, 0.03 ** 2


def _box3(t):
    """3x3 valid box filter over the two middle axes of an NHWC tensor."""
    return F.avg_pool2d(t.movedim(3, 1), kernel_size=3, stride=1).movedim(1, 3)


class SSIM(nn.Module):
    """Masked 3x3 SSIM dissimilarity, [B,H,W,C] x [B,H,W,C] x [B,H,W,1] -> [B,H-2,W-2,C] (modules.py:17-52)."""

    def forward(self, x, y, mask):
        # one pooling pass over the five moment images instead of five pooling modules
        c = x.shape[3]
        mom = _box3(torch.cat((x, y, x * x, y * y, x * y), dim=3))
        mu_x, mu_y, xx, yy, xy = (mom[..., i * c:(i + 1) * c] for i in range(5))
        var_x, var_y, cov = xx - mu_x * mu_x, yy - mu_y * mu_y, xy - mu_x * mu_y
        num = (2 * mu_x * mu_y + _C1) * (2 * cov + _C2)
        den = (mu_x * mu_x + mu_y * mu_y + _C1) * (var_x + var_y + _C2)
        return _box3(mask) * ((1 - num / den) * 0.5).clamp(0, 1)


def gradient_x(img):
    """img[x] - img[x+1] along W (modules.py:55-56)."""
    return -torch.diff(img, dim=2)


def gradient_y(img):
    """img[y] - img[y+1] along H (modules.py:58-59)."""
    return -torch.diff(img, dim=1)


def gradient(pred):
    """(forward difference along W, along H) (modules.py:61-64)."""
    return torch.diff(pred, dim=2), torch.diff(pred, dim=1)


def depth_smoothness(depth, img, lambda_wt=1):
    """Edge-aware first-order smoothness (modules.py:67-77): depth [B,H,W,1], img [B,H,W,C]."""
    total = 0
    for axis in (2, 1):
        edge = torch.diff(img, dim=axis).abs().mean(dim=3, keepdim=True)
        total = total + (torch.diff(depth, dim=axis) * torch.exp(-lambda_wt * edge)).abs().mean()
    return total


def compute_reconstr_loss(warped, ref, mask, simple=True):
    """Smooth-L1 of the masked colours, plus (simple=False) of their forward differences, 50/50 (modules.py:80-90)."""
    a, r = warped * mask, ref * mask
    photo = F.smooth_l1_loss(a, r)
    if simple:
        return photo
    edges = sum(F.smooth_l1_loss(torch.diff(a, dim=ax), torch.diff(r, dim=ax)) for ax in (2, 1))
    return 0.5 * photo + 0.5 * edges
