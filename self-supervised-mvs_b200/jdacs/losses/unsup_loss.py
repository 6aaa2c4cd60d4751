"""Mirror of `losses/unsup_loss.py` (jdacs/losses/unsup_loss.py:19-83): UnSupLoss with the photometric warp on the
B200 kernel; reconstruction / SSIM / smoothness / top-3 view selection stay PyTorch compositions."""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .homography import inverse_warping
from .modules import SSIM, compute_reconstr_loss, depth_smoothness


class UnSupLoss(nn.Module):
    """forward(imgs [B,N,3,H,W], cams [B,N,2,4,4], depth [B,H/4,W/4]) -> scalar; also sets .reconstr_loss, .ssim_loss,
    .smooth_loss, .unsup_loss (read by jdacs/train.py:142,211-212).  Needs N >= 4: top-k with k=3 (hazard H5).

    downscale / smooth_weight / smooth_lambda select the variant: jdacs = (True, 0.18, args.smooth_lambda=1.0),
    jdacs-ms = (False, 0.05, 1.0)."""

    def __init__(self, downscale=True, smooth_weight=0.18, smooth_lambda=1.0):
        super().__init__()
        self.ssim = SSIM()
        self.downscale = downscale
        self.smooth_weight = smooth_weight
        self.smooth_lambda = smooth_lambda

    def _prep(self, img):
        if self.downscale:
            img = F.interpolate(img, scale_factor=0.25, mode='bilinear')
        return img.permute(0, 2, 3, 1)

    def forward(self, imgs, cams, depth):
        imgs = torch.unbind(imgs, 1)
        cams = torch.unbind(cams, 1)
        assert len(imgs) == len(cams), "Different number of images and projection matrices"
        num_views = len(imgs)
        ref_img, ref_cam = self._prep(imgs[0]), cams[0]
        self.reconstr_loss = 0
        self.ssim_loss = 0
        self.smooth_loss = 0
        reprojection_losses = []
        for view in range(1, num_views):
            view_img = self._prep(imgs[view])
            warped_img, mask = inverse_warping(view_img, ref_cam, cams[view], depth)
            reconstr_loss = compute_reconstr_loss(warped_img, ref_img, mask, simple=False)
            reprojection_losses.append(reconstr_loss + 1e4 * (1 - mask))
            if view < 3:
                self.ssim_loss += torch.mean(self.ssim(ref_img, warped_img, mask))
        self.smooth_loss += depth_smoothness(depth.unsqueeze(dim=-1), ref_img, self.smooth_lambda)
        reprojection_volume = torch.stack(reprojection_losses).permute(1, 2, 3, 4, 0)
        top_vals, _ = torch.topk(torch.neg(reprojection_volume), k=3, sorted=False)
        top_vals = torch.neg(top_vals)
        top_vals = top_vals * (top_vals < 1e4).float()
        self.reconstr_loss = torch.mean(torch.sum(top_vals, dim=-1))
        self.unsup_loss = 12 * self.reconstr_loss + 6 * self.ssim_loss + self.smooth_weight * self.smooth_loss
        return self.unsup_loss
