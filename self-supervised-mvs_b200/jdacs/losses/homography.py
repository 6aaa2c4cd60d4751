"""Mirror of `losses/homography.py`: inverse_warping (jdacs/losses/homography.py:186-238, jdacs-ms twin :186-243).

One kernel for the forward (camera composition + back-projection + projection + clamped bilinear gather + mask),
one for the backward (gradient to depth through the sampling coordinates; optional gradient to the image)."""
from __future__ import annotations

from ... import ops


def inverse_warping(img, left_cam, right_cam, depth):
    """img [B,H,W,C], left_cam/right_cam [B,2,4,4] ([:,0]=E, [:,1,:3,:3]=K), depth [B,H,W]
    -> (warped [B,H,W,C], mask [B,H,W,1]).  Keeps the reference's use of the LEFT intrinsics for both views (H6)
    and its mask / clamped-weight conventions (H7)."""
    return ops.inverse_warp(img, left_cam, right_cam, depth)
