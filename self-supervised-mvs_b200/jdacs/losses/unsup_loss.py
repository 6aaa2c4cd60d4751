"""Mirror of `losses/unsup_loss.py` (jdacs/losses/unsup_loss.py:19-83): UnSupLoss as ONE fused forward and ONE fused backward
call of libmvs_b200.so (mvs_unsup_loss_fwd / _bwd, csrc/loss.cu): resize, photometric warp of every source view, masked
smooth-L1 of colours and colour gradients, SSIM, depth smoothness and the top-3 view selection."""
from __future__ import annotations

import torch.nn as nn

from ... import ops
from .modules import SSIM


class UnSupLoss(nn.Module):
    """forward(imgs [B,N,3,H,W], cams [B,N,2,4,4], depth [B,H/4,W/4]) -> scalar; also sets .reconstr_loss, .ssim_loss,
    .smooth_loss, .unsup_loss (read by jdacs/train.py:142,211-212).  Needs N >= 4: top-k with k=3 (hazard H5).

    downscale / smooth_weight / smooth_lambda select the variant: jdacs = (True, 0.18, args.smooth_lambda=1.0),
    jdacs-ms = (False, 0.05, 1.0).  The gradient reaches `depth`; the views are data."""

    def __init__(self, downscale=True, smooth_weight=0.18, smooth_lambda=1.0):
        super().__init__()
        self.ssim = SSIM()          # attribute kept for API parity (the reference builds it in __init__, :21-22)
        self.downscale = downscale
        self.smooth_weight = smooth_weight
        self.smooth_lambda = smooth_lambda

    def forward(self, imgs, cams, depth):
        assert imgs.shape[1] == cams.shape[1], "Different number of images and projection matrices"
        want = 4 if self.downscale else 1
        if imgs.shape[-2] // want != depth.shape[-2] or imgs.shape[-1] // want != depth.shape[-1]:
            raise ValueError("UnSupLoss(downscale=%s) needs views at %dx the depth map's size, got %s vs %s"
                             % (self.downscale, want, tuple(imgs.shape[-2:]), tuple(depth.shape[-2:])))
        out = ops.unsup_loss(imgs, cams, depth, self.smooth_lambda, self.smooth_weight)
        self.unsup_loss, self.reconstr_loss, self.ssim_loss, self.smooth_loss = out[0], out[1], out[2], out[3]
        return self.unsup_loss
