"""Depth-map fusion on the B200 library: what the reference does by shelling out to the vendored Gipuma `fusibile` binary
(jdacs/fusion/depthfusion.py:366-390 builds the command line; the work is fusibile.cu:138-277, run once per reference view by
main.cpp's runFusibile loop :560-640).

    camera_block(K, E)            the per-view constants the kernel consumes (P, inv(P[:, :3]), P[:, 3], centre, focal length)
    fuse_view(...)                one reference view against a subset of views -> fused points / normals / colours + mask
    fuse_scene(...)               the loop over reference views, concatenated point cloud [n, 3], normals, colours

depthfusion.py's defaults: --disp_threshold 0.25, --num_consistent 3 (:394-397); fusibile's own normal threshold is 30 degrees
(0.52 rad, algorithmparameters.h:89).  MVSNet outputs no normals: depthfusion.py writes a constant fake normal map
(fake_gipuma_normal :205-222), i.e. the normal test always passes -- `constant_normals` builds the same."""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np
import torch

from ... import ops


def camera_block(intrinsics, extrinsics) -> torch.Tensor:
    """K [3,3], E [4,4] (world -> camera) -> 32 floats: P = K E[:3] | inv(P[:, :3]) | P[:, 3] | C = -inv(P[:, :3]) P[:, 3] | f = K[0, 0]."""
    k = np.asarray(intrinsics, dtype=np.float64)
    e = np.asarray(extrinsics, dtype=np.float64)
    p = k @ e[:3]
    m_inv = np.linalg.inv(p[:, :3])
    c = -m_inv @ p[:, 3]
    out = np.zeros(32, dtype=np.float32)
    out[0:12], out[12:21], out[21:24], out[24:27], out[27] = p.reshape(-1), m_inv.reshape(-1), p[:, 3], c, k[0, 0]
    return torch.from_numpy(out)


def constant_normals(depths: torch.Tensor) -> torch.Tensor:
    """[V,H,W] depth maps -> [V,H,W,4] with the reference's fake normal (0, 0, -1)... any constant unit vector passes the test."""
    nd = torch.zeros(*depths.shape, 4, dtype=torch.float32, device=depths.device)
    nd[..., 2] = 1.0
    nd[..., 3] = depths
    return nd


def fuse_view(normals_depths: torch.Tensor, cams: torch.Tensor, ref: int, subset: Optional[Sequence[int]] = None, disp_thresh: float = 0.25,
              normal_thresh: float = 0.52, num_consistent: int = 3, images: Optional[torch.Tensor] = None):
    subset = list(range(normals_depths.shape[0])) if subset is None else list(subset)
    return ops.fusibile(normals_depths, cams, ref, subset, disp_thresh, normal_thresh, num_consistent, images)


def fuse_scene(normals_depths: torch.Tensor, cams: torch.Tensor, disp_thresh: float = 0.25, normal_thresh: float = 0.52,
               num_consistent: int = 3, images: Optional[torch.Tensor] = None):
    """Every view as the reference view in turn (main.cpp's loop): -> (xyz [n,3], normals [n,3], colours [n,3])."""
    xyz, nrm, col = [], [], []
    for ref in range(normals_depths.shape[0]):
        pts, valid = fuse_view(normals_depths, cams, ref, None, disp_thresh, normal_thresh, num_consistent, images)
        sel = pts[valid]
        xyz.append(sel[:, 0:3]); nrm.append(sel[:, 4:7]); col.append(sel[:, 8:11])
    return torch.cat(xyz), torch.cat(nrm), torch.cat(col)
