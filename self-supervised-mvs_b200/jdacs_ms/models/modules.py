"""Mirror of the reference's jdacs-ms `models/modules.py` surface, B200 path underneath.

Same names / signatures: conv, conditionIntrinsics, calInitDepthInterval, calSweepingDepthHypo, homo_warping,
calDepthHypo, proj_cost, ConvBnReLU3D, depth_regression, depth_regression_refine.  No hard-coded `.cuda()`
(hazard H4): everything follows the device of its inputs.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import ops
from ...jdacs.models.module import ALIGN_CORNERS, ConvBn, ConvBn3D, ConvBnReLU, ConvBnReLU3D  # noqa: F401


def conv(in_planes, out_planes, kernel_size=3, stride=1, padding=1, dilation=1):
    """jdacs-ms/models/modules.py:15-19 (2-D, library code; FeaturePyramid block)."""
    return nn.Sequential(nn.Conv2d(in_planes, out_planes, kernel_size, stride, padding, dilation, bias=True),
                         nn.LeakyReLU(0.1))


def conditionIntrinsics(intrinsics, img_shape, fp_shapes):
    """jdacs-ms/models/modules.py:22-37: K[:2] / (image height / level height) per pyramid level -> [B,nScale,3,3]."""
    out = []
    for fp_shape in fp_shapes:
        k = intrinsics.clone()
        k[:, :2, :] = k[:, :2, :] / (img_shape[2] / fp_shape[2])
        out.append(k)
    return torch.stack(out).permute(1, 0, 2, 3)


def calInitDepthInterval(ref_in, src_in, ref_ex, src_ex, pixel_interval):
    """jdacs-ms/models/modules.py:40-41."""
    return 165


def calSweepingDepthHypo(ref_in, src_in, ref_ex, src_ex, depth_min, depth_max, nhypothesis_init=48):
    """jdacs-ms/models/modules.py:44-59: nhypothesis_init uniform planes over the FIRST batch item's range.

    The reference's deprecated inclusive torch.range can drop the last plane under fp32 rounding (hazard H3,
    47 planes for DTU's own 425..1065); the intended count is produced here: d_k = dmin + k (dmax-dmin)/(n-1)."""
    assert nhypothesis_init % 2 == 0
    step = (depth_max[0] - depth_min[0]) / (nhypothesis_init - 1)
    planes = depth_min[0] + step * torch.arange(nhypothesis_init, dtype=torch.float32, device=depth_min.device)
    return planes.unsqueeze(0).repeat(ref_in.shape[0], 1).to(ref_in.device)


def homo_warping(src_feature, ref_in, src_in, ref_ex, src_ex, depth_hypos):
    """jdacs-ms/models/modules.py:62-104: warp from (K, E) pairs; depth_hypos [B,D] -> [B,C,D,H,W]."""
    rt = ops.compose_proj_ke(ref_in, src_in.unsqueeze(1), ref_ex, src_ex.unsqueeze(1), 1.0)
    return ops.homo_warp(src_feature, rt[0], depth_hypos, ALIGN_CORNERS)


def calDepthHypo(netArgs, ref_depths, ref_intrinsics, src_intrinsics, ref_extrinsics, src_extrinsics, depth_min,
                 depth_max, level):
    """jdacs-ms/models/modules.py:107-206: [B,H,W] -> [B,8,H,W] hypotheses around the upsampled depth
    (fp64 epipolar construction against source view 0, one scalar interval per batch item)."""
    with torch.no_grad():
        return ops.depth_hypo_refine(ref_depths, ref_intrinsics, src_intrinsics[:, 0], ref_extrinsics,
                                     src_extrinsics[:, 0], 4)


def proj_cost(settings, ref_feature, src_feature, level, ref_in, src_in, ref_ex, src_ex, depth_hypos,
              volume_dtype=torch.float32, as_c8=False):
    """jdacs-ms/models/modules.py:209-261: per-pixel-hypothesis variance cost volume, one fused kernel.

    src_feature is the reference's list[nsrc] of pyramids (list of levels).  Returns [B,C,D,H,W] fp32 like the
    reference, or the C8 volume when as_c8 (what CostRegNet consumes directly).  The variance keeps the
    reference's aliasing of volume_sum with ref^2 (hazard H2)."""
    nsrc = settings.nsrc
    rt = ops.compose_proj_ke(ref_in, src_in[:, :nsrc], ref_ex, src_ex[:, :nsrc], 1.0)
    srcs = [src_feature[s][level] for s in range(nsrc)]
    vol = ops.warp_variance(ref_feature, srcs, rt, depth_hypos, volume_dtype, ALIGN_CORNERS, True)
    if as_c8:
        return vol
    from ... import regnet
    return regnet.unpack_c8_grad(vol)


def depth_regression(p, depth_values):
    """jdacs-ms/models/modules.py:324-327."""
    depth_values = depth_values.view(*depth_values.shape, 1, 1)
    return torch.sum(p * depth_values, 1)


def depth_regression_refine(prob_volume, depth_hypothesis):
    """jdacs-ms/models/modules.py:330-331."""
    return torch.sum(prob_volume * depth_hypothesis, 1)
