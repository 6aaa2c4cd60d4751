"""Deterministic synthetic DTU-shaped inputs for the plane-sweep hot path.

There is no dataset on the build or GPU boxes, so every test, the oracle fixtures and
bench.py draw their inputs from here.  Shapes and value ranges follow
what the reference data loaders hand to the model:

  * images are per-image standardised -> N(0,1)          (jdacs/datasets/dtu_yao.py:94-99)
  * proj_matrices[b, v] = [[K @ E[:3]], [0,0,0,1]] at FEATURE resolution
                                                          (jdacs/datasets/dtu_yao.py:273-275)
  * cams[b, v, 0] = E (4x4), cams[b, v, 1, :3, :3] = K   (jdacs/datasets/dtu_yao.py:266-272)
  * depth_values = depth_min + interval * arange(D)       (jdacs/datasets/dtu_yao.py:290, hazard H14)

Everything is generated on the CPU with a fixed seed and moved by the caller.
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch

# feature-resolution intrinsics of a 160x128 DTU depth map (fx, fy, cx, cy)
K_FEAT = (361.54125, 360.3975, 82.900625, 66.383875)
# (alpha [rot about y], beta [rot about x], translation mm) per source view; non-degenerate for
# the CVP depth-hypothesis construction (needs rotation / y-baseline)
SRC_POSES = (
    (0.10, 0.05, (-60.0, 25.0, 10.0)),
    (-0.12, -0.04, (70.0, -30.0, 5.0)),
    (0.06, -0.08, (-30.0, -55.0, 8.0)),
    (-0.05, 0.09, (40.0, 60.0, -6.0)),
    (0.08, 0.07, (55.0, 35.0, 4.0)),
    (-0.09, 0.03, (-45.0, 50.0, -5.0)),
)
DEPTH_MIN = 425.0
DEPTH_INTERVAL = 2.65


def intrinsics(width: int = 160, height: int = 128) -> np.ndarray:
    """3x3 K for a (height, width) map: 160x128 is the DTU feature map, 640x512 the image.

    Focal length and principal point scale with the map size so that odd test sizes stay centred."""
    fx, fy, cx, cy = K_FEAT
    sx = width / 160.0
    sy = height / 128.0
    return np.array([[fx * sx, 0.0, cx * sx], [0.0, fy * sy, cy * sy], [0.0, 0.0, 1.0]], dtype=np.float64)


def extrinsics(view: int) -> np.ndarray:
    """4x4 world->camera E of view `view` (0 = reference = identity)."""
    E = np.eye(4, dtype=np.float64)
    if view == 0:
        return E
    a, b, t = SRC_POSES[(view - 1) % len(SRC_POSES)]
    # later wraps get a slightly different pose so that views never coincide
    k = (view - 1) // len(SRC_POSES)
    a, b = a * (1.0 + 0.1 * k), b * (1.0 - 0.1 * k)
    ry = np.array([[math.cos(a), 0, math.sin(a)], [0, 1, 0], [-math.sin(a), 0, math.cos(a)]])
    rx = np.array([[1, 0, 0], [0, math.cos(b), -math.sin(b)], [0, math.sin(b), math.cos(b)]])
    E[:3, :3] = ry @ rx
    E[:3, 3] = np.asarray(t) * (1.0 + 0.05 * k)
    return E


def mvsnet_inputs(batch: int = 1, views: int = 5, height: int = 512, width: int = 640, ndepth: int = 192,
                  seed: int = 0, batch_jitter: bool = True) -> Dict[str, torch.Tensor]:
    """Inputs of `MVSNet.forward` / `UnSupLoss.forward` (jdacs/models/mvsnet.py:105, losses/unsup_loss.py:24)."""
    g = torch.Generator().manual_seed(seed)
    hf, wf = height // 4, width // 4
    imgs = torch.randn(batch, views, 3, height, width, generator=g, dtype=torch.float32)
    proj = torch.zeros(batch, views, 4, 4, dtype=torch.float64)
    cams = torch.zeros(batch, views, 2, 4, 4, dtype=torch.float64)
    K = intrinsics(width=wf, height=hf)
    for b in range(batch):
        for v in range(views):
            E = extrinsics(v).copy()
            if batch_jitter and v > 0:
                E[:3, 3] *= 1.0 + 0.03 * b  # batch items see slightly different baselines
            P = np.eye(4)
            P[:3, :4] = K @ E[:3, :4]
            proj[b, v] = torch.from_numpy(P)
            cams[b, v, 0] = torch.from_numpy(E)
            cams[b, v, 1, :3, :3] = torch.from_numpy(K)
            cams[b, v, 1, 3, 0] = DEPTH_MIN
            cams[b, v, 1, 3, 1] = DEPTH_INTERVAL
    depth_values = (DEPTH_MIN + DEPTH_INTERVAL * torch.arange(ndepth, dtype=torch.float64)).to(torch.float32)
    depth_values = depth_values.unsqueeze(0).repeat(batch, 1).contiguous()
    return {"imgs": imgs, "proj_matrices": proj.float(), "cams": cams.float(), "depth_values": depth_values}


def feature_inputs(batch: int = 1, views: int = 5, channels: int = 32, hf: int = 128, wf: int = 160,
                   ndepth: int = 192, seed: int = 0) -> Dict[str, torch.Tensor]:
    """Kernel-level inputs: feature maps instead of images (features[v] is [B,C,Hf,Wf])."""
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(views, batch, channels, hf, wf, generator=g, dtype=torch.float32)
    base = mvsnet_inputs(batch, views, hf * 4, wf * 4, ndepth, seed)
    return {"features": feats, "proj_matrices": base["proj_matrices"], "cams": base["cams"],
            "depth_values": base["depth_values"]}


def cvp_inputs(batch: int = 1, nsrc: int = 4, height: int = 512, width: int = 640, seed: int = 0,
               depth_max: float = 935.0) -> Dict[str, torch.Tensor]:
    """Inputs of `CVPMVSNet.forward` (jdacs-ms/models/network.py:84): full-resolution K, separate E.

    depth range 425..935 is one for which the reference's `torch.range` yields 48 planes (hazard H3)."""
    g = torch.Generator().manual_seed(seed)
    ref_img = torch.randn(batch, 3, height, width, generator=g)
    src_imgs = torch.randn(batch, nsrc, 3, height, width, generator=g)
    K = torch.from_numpy(intrinsics(width=width, height=height)).float()
    ref_in = K.unsqueeze(0).repeat(batch, 1, 1).contiguous()
    src_in = K.view(1, 1, 3, 3).repeat(batch, nsrc, 1, 1).contiguous()
    ref_ex = torch.from_numpy(extrinsics(0)).float().unsqueeze(0).repeat(batch, 1, 1).contiguous()
    src_ex = torch.stack([torch.from_numpy(extrinsics(v + 1)).float() for v in range(nsrc)], 0)
    src_ex = src_ex.unsqueeze(0).repeat(batch, 1, 1, 1).contiguous()
    return {"ref_img": ref_img, "src_imgs": src_imgs, "ref_in": ref_in, "src_in": src_in, "ref_ex": ref_ex,
            "src_ex": src_ex, "depth_min": torch.full((batch,), DEPTH_MIN),
            "depth_max": torch.full((batch,), float(depth_max))}


def plausible_depth(batch: int, hf: int, wf: int, seed: int = 0) -> torch.Tensor:
    """A smooth depth map inside the sweep range, for the loss-warp (a10) tests."""
    g = torch.Generator().manual_seed(seed + 17)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, hf), torch.linspace(0, 1, wf), indexing="ij")
    d = 600.0 + 120.0 * torch.sin(3.0 * xx + 0.5) * torch.cos(2.0 * yy) + 40.0 * xx
    d = d.unsqueeze(0).repeat(batch, 1, 1) + 2.0 * torch.randn(batch, hf, wf, generator=g)
    return d.contiguous()


def randomise_bn(model: torch.nn.Module, seed: int = 5) -> None:
    """Give every BatchNorm of `model` non-trivial affine parameters and running statistics (the fixtures of
    oracle/gen_golden.py use the same recipe): with the default statistics (mean 0, var 1) and default-initialised
    convolutions the regularised volume is almost flat along depth and the softmax is uniform (hazard H11)."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for m in model.modules():
            if isinstance(m, torch.nn.modules.batchnorm._BatchNorm):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=g))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=g))
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.weight.shape, generator=g))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=g))
