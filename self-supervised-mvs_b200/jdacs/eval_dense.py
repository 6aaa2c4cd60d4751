"""Mirror of the output side of `eval_dense.py` (jdacs/eval_dense.py:110-232; the same functions in jdacs/eval.py:110-224 and the
writer half of jdacs-ms/test.py:96-165): what is done with the network's depth / confidence maps.

    resize_outputs               F.interpolate(..., size=(1200, 1600)) of both maps (:150-153)              one kernel launch
    save_depth_outputs           save_pfm x 2 + write_depth_img per item (:160-173)                         rows flipped on the GPU
    reproject_with_depth         (:177-214)  \\  one kernel over all pixels (and all pairs of a batch): fp64 per-pixel algebra,
    check_geometric_consistency  (:217-232)  /   cv2.remap's 1/32-pixel bilinear sampling reproduced exactly

The two geometry functions keep the reference's signatures (NumPy arrays in, NumPy arrays out) and also accept CUDA tensors
(then tensors come back); `check_geometric_consistency_batch` is the form a fusion loop should call: one launch for all source
views of a reference view.  The 4x4 / 3x3 camera algebra is done on the host in float32 exactly as the reference's
np.linalg.inv / np.matmul calls do it; everything per pixel runs on the device."""
from __future__ import annotations

import os

import numpy as np
import torch

from .. import ops
from .datasets.data_io import save_pfm_flipped


def _pair_cams(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src) -> np.ndarray:
    """The six matrices the kernel needs, with the reference's float32 algebra (eval_dense.py:184-210), as 60 doubles."""
    f32 = lambda a: np.asarray(a.detach().cpu() if isinstance(a, torch.Tensor) else a, dtype=np.float32)
    kr, er, ks, es = f32(intrinsics_ref), f32(extrinsics_ref), f32(intrinsics_src), f32(extrinsics_src)
    parts = (np.linalg.inv(kr), np.matmul(es, np.linalg.inv(er))[:3], ks, np.linalg.inv(ks), np.matmul(er, np.linalg.inv(es))[:3], kr)
    return np.concatenate([p.astype(np.float64).reshape(-1) for p in parts])


def _device_of(*arrays):
    for a in arrays:
        if isinstance(a, torch.Tensor) and a.is_cuda:
            return a.device
    return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")


def _run_pairs(depth_ref, depth_srcs, cams, apply_mask):
    numpy_in = not isinstance(depth_ref, torch.Tensor)
    dev = _device_of(depth_ref, *depth_srcs)
    to_dev = lambda a: (a if isinstance(a, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))).to(dev)
    n = len(depth_srcs)
    dr = to_dev(depth_ref).unsqueeze(0).expand(n, -1, -1)
    ds = torch.stack([to_dev(d) for d in depth_srcs])
    out = ops.geo_consistency(dr, ds, torch.from_numpy(np.stack(cams)).to(dev), 1.0, 0.01, apply_mask)
    return [o.cpu().numpy() for o in out] if numpy_in else list(out)


def reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """-> depth_reprojected, x_reprojected, y_reprojected, x_src, y_src ([H,W] float32).  jdacs/eval_dense.py:177-214."""
    cams = _pair_cams(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src)
    _, drep, xs, ys, xr, yr = _run_pairs(depth_ref, [depth_src], [cams], apply_mask=False)
    return drep[0], xr[0], yr[0], xs[0], ys[0]


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    """-> mask (bool), depth_reprojected (zero outside the mask), x2d_src, y2d_src.  jdacs/eval_dense.py:217-232."""
    cams = _pair_cams(intrinsics_ref, extrinsics_ref, intrinsics_src, extrinsics_src)
    mask, drep, xs, ys, _, _ = _run_pairs(depth_ref, [depth_src], [cams], apply_mask=True)
    return mask[0], drep[0], xs[0], ys[0]


def check_geometric_consistency_batch(depth_ref, intrinsics_ref, extrinsics_ref, depth_srcs, intrinsics_srcs, extrinsics_srcs):
    """One launch for all source views of a reference view: -> mask [S,H,W], depth_reprojected [S,H,W], x2d_src, y2d_src."""
    cams = [_pair_cams(intrinsics_ref, extrinsics_ref, k, e) for k, e in zip(intrinsics_srcs, extrinsics_srcs)]
    mask, drep, xs, ys, _, _ = _run_pairs(depth_ref, list(depth_srcs), cams, apply_mask=True)
    return mask, drep, xs, ys


def resize_outputs(outputs, size=(1200, 1600), flip_rows=False):
    """outputs["depth"], outputs["photometric_confidence"] [B,h,w] -> [B,size] (nearest), both maps in one launch
    (jdacs/eval_dense.py:150-155).  flip_rows=True: rows bottom-up, ready for save_pfm_flipped."""
    d, c = outputs["depth"], outputs["photometric_confidence"]
    both = ops.upsample_nearest(torch.cat((d, c), dim=0), size, flip_rows)
    out = dict(outputs)
    out["depth"], out["photometric_confidence"] = both[:d.shape[0]], both[d.shape[0]:]
    return out


def write_depth_img(filename, depth):
    """8-bit preview ((depth - 500) / 2, clamped, truncated) as a PNG.  jdacs/eval_dense.py:110-121."""
    from PIL import Image
    os.makedirs(os.path.dirname(filename) or ".", exist_ok=True)
    t = depth if isinstance(depth, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(depth, dtype=np.float32))
    dev = _device_of(t)
    Image.fromarray(ops.depth_preview_u8(t.to(dev)).cpu().numpy(), mode="L").save(filename)
    return 1


def save_depth_outputs(outputs, filenames, outdir, size=(1200, 1600), preview=True):
    """The writer loop of save_depth (jdacs/eval_dense.py:147-173): resize on the GPU with the rows already in .pfm order, one
    D2H copy per batch, then depth_est/*.pfm, confidence/*.pfm (+ the .png preview) per item."""
    flipped = resize_outputs(outputs, size, flip_rows=True)
    depth, conf = flipped["depth"].cpu().numpy(), flipped["photometric_confidence"].cpu().numpy()
    prev = ops.depth_preview_u8(flipped["depth"]).cpu().numpy() if preview else None
    for i, filename in enumerate(filenames):
        depth_filename = os.path.join(outdir, filename.format("depth_est", ".pfm"))
        confidence_filename = os.path.join(outdir, filename.format("confidence", ".pfm"))
        os.makedirs(os.path.dirname(depth_filename) or ".", exist_ok=True)
        os.makedirs(os.path.dirname(confidence_filename) or ".", exist_ok=True)
        save_pfm_flipped(depth_filename, depth[i])
        save_pfm_flipped(confidence_filename, conf[i])
        if preview:
            from PIL import Image
            Image.fromarray(np.flipud(prev[i]), mode="L").save(depth_filename + ".png")
