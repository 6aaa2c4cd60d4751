"""Mirror of the two functions train.py imports from `models/augmentations.py` (jdacs/train.py:31; definitions
jdacs/models/augmentations.py:107-130): `random_image_mask` and `aug_loss`, the data-augmentation branch's mask and loss
(SURVEY 8f-2).  The PIL-based colour / blur transforms of that file are data-loader code and are not mirrored.

Both are written without a data-dependent shape, so the training batch can be captured in a CUDA graph (trainer.GraphedTrainStep):
the box corner may be a device tensor, and the loss is a masked mean instead of a boolean gather."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def random_image_mask(img, filter_size, box=None):
    """img [B,3,H,W]; zero a random fh x fw box -> (img * mask, mask [B,3,H,W]); (img, None) when the box is the whole image.
    The corner is drawn with np.random.randint like the reference unless `box` = (x, y) (ints or a device tensor) is given."""
    fh, fw = filter_size
    _, _, h, w = img.size()
    if fh == h and fw == w:
        return img, None
    if box is None:
        x = np.random.randint(0, w - fw)
        y = np.random.randint(0, h - fh)
    else:
        x, y = box[0], box[1]
    ys = torch.arange(h, device=img.device).view(h, 1) - y
    xs = torch.arange(w, device=img.device).view(1, w) - x
    filter_mask = (~((ys >= 0) & (ys < fh) & (xs >= 0) & (xs < fw))).to(img.dtype).expand_as(img)
    return img * filter_mask, filter_mask


def aug_loss(depth_est, depth_gt, mask):
    """smooth-L1 between two depth maps over the pixels where mask > 0.5 (mean over those pixels)."""
    m = (mask > 0.5).to(depth_est.dtype)
    return (F.smooth_l1_loss(depth_est, depth_gt, reduction="none") * m).sum() / m.sum()
