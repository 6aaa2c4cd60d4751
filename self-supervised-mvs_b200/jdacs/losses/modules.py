"""Mirror of `losses/modules.py` (jdacs/losses/modules.py): SSIM, gradients, depth_smoothness, compute_reconstr_loss.

Small element-wise / 3x3-pool terms on 128x160 maps; they stay PyTorch (out of the plane-sweep scope)."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F


class SSIM(nn.Module):
    """jdacs/losses/modules.py:17-52: masked 3x3 SSIM dissimilarity on NHWC maps."""

    def __init__(self):
        super().__init__()
        self.C1 = 0.01 ** 2
        self.C2 = 0.03 ** 2

    def forward(self, x, y, mask):
        x, y, mask = x.permute(0, 3, 1, 2), y.permute(0, 3, 1, 2), mask.permute(0, 3, 1, 2)
        pool = lambda t: F.avg_pool2d(t, 3, 1)
        mu_x, mu_y = pool(x), pool(y)
        sigma_x = pool(x ** 2) - mu_x ** 2
        sigma_y = pool(y ** 2) - mu_y ** 2
        sigma_xy = pool(x * y) - mu_x * mu_y
        n = (2 * mu_x * mu_y + self.C1) * (2 * sigma_xy + self.C2)
        d = (mu_x ** 2 + mu_y ** 2 + self.C1) * (sigma_x + sigma_y + self.C2)
        out = pool(mask) * torch.clamp((1 - n / d) / 2, 0, 1)
        return out.permute(0, 2, 3, 1)


def gradient_x(img):
    return img[:, :, :-1, :] - img[:, :, 1:, :]


def gradient_y(img):
    return img[:, :-1, :, :] - img[:, 1:, :, :]


def gradient(pred):
    D_dy = pred[:, 1:, :, :] - pred[:, :-1, :, :]
    D_dx = pred[:, :, 1:, :] - pred[:, :, :-1, :]
    return D_dx, D_dy


def depth_smoothness(depth, img, lambda_wt=1):
    """jdacs/losses/modules.py:67-77."""
    weights_x = torch.exp(-(lambda_wt * torch.mean(torch.abs(gradient_x(img)), 3, keepdim=True)))
    weights_y = torch.exp(-(lambda_wt * torch.mean(torch.abs(gradient_y(img)), 3, keepdim=True)))
    return torch.mean(torch.abs(gradient_x(depth) * weights_x)) + torch.mean(torch.abs(gradient_y(depth) * weights_y))


def compute_reconstr_loss(warped, ref, mask, simple=True):
    """jdacs/losses/modules.py:80-90."""
    if simple:
        return F.smooth_l1_loss(warped * mask, ref * mask, reduction='mean')
    alpha = 0.5
    ref_dx, ref_dy = gradient(ref * mask)
    warped_dx, warped_dy = gradient(warped * mask)
    photo_loss = F.smooth_l1_loss(warped * mask, ref * mask, reduction='mean')
    grad_loss = F.smooth_l1_loss(warped_dx, ref_dx, reduction='mean') + F.smooth_l1_loss(warped_dy, ref_dy, reduction='mean')
    return (1 - alpha) * photo_loss + alpha * grad_loss
