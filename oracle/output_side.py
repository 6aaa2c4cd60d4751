"""ORACLE (test infrastructure, never imported by the product): NumPy restatement of the inference output side.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.  Every function cites the reference
lines it follows; it is pinned against the reference's own functions (executed from /root/reference by oracle/gen_golden_output.py)
through tests/golden/jdacs_output_side.npz, and the remap restatement against cv2.remap itself.

    remap_bilinear               cv2.remap(src, x, y, cv2.INTER_LINEAR) for a float32 map and float32 coordinates (OpenCV 4.x
                                 imgproc/src/imgwarp.cpp: RemapInvoker + remapBilinear): coordinates rounded half-to-even to
                                 1/32 pixel, float32 blend with the table weights, constant-0 border.  Third-party arithmetic
                                 (opencv-python, not pinned by the reference) restated from the call site jdacs/eval_dense.py:199.
    reproject_with_depth         jdacs/eval_dense.py:177-214
    check_geometric_consistency  jdacs/eval_dense.py:217-232
    upsample_nearest             F.interpolate(x.unsqueeze(1), size=...) default mode, jdacs/eval_dense.py:150-153 (ATen nearest)
    pfm_bytes                    save_pfm, jdacs/datasets/data_io.py:53-80
    depth_preview                write_depth_img, jdacs/eval_dense.py:110-121 (Pillow F -> L conversion)
"""
import sys

import numpy as np


def remap_bilinear(src, x, y):
    src = np.asarray(src, dtype=np.float32)
    h, w = src.shape
    with np.errstate(invalid="ignore", over="ignore"):
        fx, fy = (np.asarray(x, np.float32) * np.float32(32.0)), (np.asarray(y, np.float32) * np.float32(32.0))
        okx, oky = np.abs(fx) < 2.0e9, np.abs(fy) < 2.0e9
        sx = np.where(okx, np.rint(np.where(okx, fx, 0)), -2.0 ** 31).astype(np.int64)      # cvRound: ties to even
        sy = np.where(oky, np.rint(np.where(oky, fy, 0)), -2.0 ** 31).astype(np.int64)
    ax, ay = (sx & 31).astype(np.float32) / np.float32(32), (sy & 31).astype(np.float32) / np.float32(32)
    ix, iy = np.clip(sx >> 5, -32768, 32767), np.clip(sy >> 5, -32768, 32767)
    one = np.float32(1)
    w0, w1, w2, w3 = (one - ay) * (one - ax), (one - ay) * ax, ay * (one - ax), ay * ax

    def tap(yy, xx):
        inside = (xx >= 0) & (xx < w) & (yy >= 0) & (yy < h)
        return np.where(inside, src[np.clip(yy, 0, h - 1), np.clip(xx, 0, w - 1)], np.float32(0))
    out = tap(iy, ix) * w0
    out = out + tap(iy, ix + 1) * w1
    out = out + tap(iy + 1, ix) * w2
    out = out + tap(iy + 1, ix + 1) * w3
    return out.astype(np.float32)


def reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    x_ref, y_ref = x_ref.reshape([-1]), y_ref.reshape([-1])
    pix = np.vstack((x_ref, y_ref, np.ones_like(x_ref)))
    with np.errstate(all="ignore"):
        xyz_ref = np.matmul(np.linalg.inv(intrinsics_ref), pix * depth_ref.reshape([-1]))
        xyz_src = np.matmul(np.matmul(extrinsics_src, np.linalg.inv(extrinsics_ref)), np.vstack((xyz_ref, np.ones_like(x_ref))))[:3]
        k_xyz_src = np.matmul(intrinsics_src, xyz_src)
        xy_src = k_xyz_src[:2] / k_xyz_src[2:3]
        x_src = xy_src[0].reshape([height, width]).astype(np.float32)
        y_src = xy_src[1].reshape([height, width]).astype(np.float32)
        sampled = remap_bilinear(depth_src, x_src, y_src)
        xyz_src = np.matmul(np.linalg.inv(intrinsics_src), np.vstack((xy_src, np.ones_like(x_ref))) * sampled.reshape([-1]))
        xyz_rep = np.matmul(np.matmul(extrinsics_ref, np.linalg.inv(extrinsics_src)), np.vstack((xyz_src, np.ones_like(x_ref))))[:3]
        depth_rep = xyz_rep[2].reshape([height, width]).astype(np.float32)
        k_xyz_rep = np.matmul(intrinsics_ref, xyz_rep)
        xy_rep = k_xyz_rep[:2] / k_xyz_rep[2:3]
    return (depth_rep, xy_rep[0].reshape([height, width]).astype(np.float32), xy_rep[1].reshape([height, width]).astype(np.float32),
            x_src, y_src)


def check_geometric_consistency(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src):
    height, width = depth_ref.shape
    x_ref, y_ref = np.meshgrid(np.arange(0, width), np.arange(0, height))
    depth_rep, x_rep, y_rep, x_src, y_src = reproject_with_depth(depth_ref, intrinsics_ref, extrinsics_ref, depth_src, intrinsics_src, extrinsics_src)
    with np.errstate(all="ignore"):
        dist = np.sqrt((x_rep - x_ref) ** 2 + (y_rep - y_ref) ** 2)
        rel = np.abs(depth_rep - depth_ref) / depth_ref
        mask = np.logical_and(dist < 1, rel < 0.01)
    depth_rep = depth_rep.copy()
    depth_rep[~mask] = 0
    return mask, depth_rep, x_src, y_src


def upsample_nearest(maps, size):
    """ATen upsample_nearest2d with an explicit output size: src = min(floor(dst * (float) in / out), in - 1), float32."""
    m, h, w = maps.shape
    ho, wo = size
    ys = np.minimum(np.floor(np.arange(ho, dtype=np.float32) * (np.float32(h) / np.float32(ho))).astype(np.int64), h - 1)
    xs = np.minimum(np.floor(np.arange(wo, dtype=np.float32) * (np.float32(w) / np.float32(wo))).astype(np.int64), w - 1)
    return maps[:, ys][:, :, xs]


def pfm_bytes(image, scale=1):
    image = np.flipud(image)
    assert image.dtype.name == "float32"
    color = image.ndim == 3 and image.shape[2] == 3
    head = ("PF\n" if color else "Pf\n") + "%d %d\n" % (image.shape[1], image.shape[0])
    if image.dtype.byteorder == "<" or (image.dtype.byteorder == "=" and sys.byteorder == "little"):
        scale = -scale
    return (head + "%f\n" % scale).encode("utf-8") + image.tobytes()


def depth_preview(depth):
    v = (np.asarray(depth, np.float32) - np.float32(500)) / np.float32(2)
    return np.where(v <= 0, 0, np.where(v >= 255, 255, v)).astype(np.uint8)
