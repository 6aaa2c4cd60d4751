"""Generate tests/golden/jdacs_output_side.npz from the UNMODIFIED reference (run in the build container, where /root/reference
exists; the fixture travels, the reference does not).

jdacs/eval_dense.py cannot be imported as a module (it parses the command line and imports plyfile at import time), so the three
functions on the output side are taken from its source text as they stand -- write_depth_img, reproject_with_depth,
check_geometric_consistency (:110-121, :177-232) -- and executed with NumPy / OpenCV / Pillow; save_pfm comes from importing
jdacs/datasets/data_io.py.  The oracle restatement (oracle/output_side.py) is checked against all of them here."""
import ast
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/jdacs"
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True


def reference_functions():
    import cv2
    from PIL import Image
    import errno
    src = open(os.path.join(REF, "eval_dense.py")).read()
    tree = ast.parse(src)
    ns = {"np": np, "cv2": cv2, "Image": Image, "os": os, "errno": errno}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in ("write_depth_img", "reproject_with_depth", "check_geometric_consistency"):
            exec(compile(ast.Module([node], []), "eval_dense.py", "exec"), ns)
    sys.path.insert(0, REF)
    from datasets.data_io import read_pfm, save_pfm
    ns["save_pfm"], ns["read_pfm"] = save_pfm, read_pfm
    return ns


def cameras(h, w):
    from importlib import import_module
    synth = import_module("ssmvs_b200.synth")
    k = synth.intrinsics(w, h).astype(np.float32)
    return k, [synth.extrinsics(v).astype(np.float32) for v in range(3)]


def surface(h, w, seed):
    g = np.random.default_rng(seed)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    return (620 + 60 * np.sin(xx / 9.0) * np.cos(yy / 7.0) + 0.4 * xx + g.normal(0, 0.05, (h, w))).astype(np.float32)


def main():
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import output_side as mine
    ref = reference_functions()
    h, w = 48, 64
    k, ex = cameras(h, w)
    # a consistent scene: the source depth map is the reference surface re-rendered in the source view (approximately: its own
    # reprojection), plus a band of wrong depths and a zero hole so that both thresholds and the remap border are exercised
    d_ref = surface(h, w, 1)
    d_rep, _, _, _, _ = ref["reproject_with_depth"](d_ref, k, ex[0], d_ref, k, ex[0])
    out = {"intrinsics": k, "depth_ref": d_ref}
    for s in (1, 2):
        d_src = surface(h, w, 1 + s) * np.float32(1.0)
        d_src[10:14] *= np.float32(1.05)
        d_src[30:34, 20:40] = 0
        res = ref["check_geometric_consistency"](d_ref.copy(), k, ex[0], d_src, k, ex[s])
        rp = ref["reproject_with_depth"](d_ref.copy(), k, ex[0], d_src, k, ex[s])
        mo = mine.check_geometric_consistency(d_ref.copy(), k, ex[0], d_src, k, ex[s])
        mp = mine.reproject_with_depth(d_ref.copy(), k, ex[0], d_src, k, ex[s])
        assert np.array_equal(res[0], mo[0]), "mask differs from the reference"
        for a, b, name in zip(res[1:] + rp, mo[1:] + mp, ("depth_rep_masked", "x_src", "y_src", "depth_rep", "x_rep", "y_rep", "x_src", "y_src")):
            assert np.allclose(a, b, rtol=1e-6, atol=1e-5, equal_nan=True), name
        print("  pair %d: geo mask coverage %.3f; oracle == reference" % (s, res[0].mean()))
        out.update({"extrinsics_ref": ex[0], "extrinsics_src%d" % s: ex[s], "depth_src%d" % s: d_src, "mask%d" % s: res[0],
                    "depth_reprojected_masked%d" % s: res[1], "x2d_src%d" % s: res[2], "y2d_src%d" % s: res[3],
                    "depth_reprojected%d" % s: rp[0], "x_reprojected%d" % s: rp[1], "y_reprojected%d" % s: rp[2]})
    # remap restatement against OpenCV itself on hostile coordinates (borders, ties at 1/64 pixel, NaN, far outside)
    import cv2
    g = np.random.default_rng(7)
    img = g.normal(0, 1, (23, 31)).astype(np.float32)
    xs = g.uniform(-3, 34, (40, 50)).astype(np.float32)
    ys = g.uniform(-3, 26, (40, 50)).astype(np.float32)
    xs[0, :8] = np.array([0.015625, 1.046875, 2.5, 29.984375, 30.0, 30.5, -0.015625, -1.0], np.float32)
    ys[1, :4] = np.array([np.nan, np.inf, -np.inf, 1e9], np.float32)
    want = cv2.remap(img, xs, ys, interpolation=cv2.INTER_LINEAR)
    got = mine.remap_bilinear(img, xs, ys)
    assert np.array_equal(want, got), "remap restatement != cv2.remap (max diff %g)" % np.nanmax(np.abs(want - got))
    out.update({"remap_img": img, "remap_x": xs, "remap_y": ys, "remap_out": want})
    # nearest upsample (ATen) and the PFM bytes / preview of the reference
    import torch
    small = g.normal(600, 50, (2, 12, 16)).astype(np.float32)
    up = torch.nn.functional.interpolate(torch.from_numpy(small).unsqueeze(1), size=(45, 70)).squeeze(1).numpy()
    assert np.array_equal(up, mine.upsample_nearest(small, (45, 70)))
    tmp = "/tmp/_golden_pfm.pfm"
    ref["save_pfm"](tmp, up[0])
    pfm = np.frombuffer(open(tmp, "rb").read(), dtype=np.uint8)
    assert pfm.tobytes() == mine.pfm_bytes(up[0])
    back, scale = ref["read_pfm"](tmp)
    assert np.array_equal(back, up[0]) and scale == 1.0
    ref["write_depth_img"]("/tmp/_golden_prev/x.png", up[0])
    from PIL import Image
    prev = np.asarray(Image.open("/tmp/_golden_prev/x.png"))
    assert np.array_equal(prev, mine.depth_preview(up[0]))
    out.update({"small": small, "upsampled": up, "pfm_bytes": pfm, "preview": prev})
    path = os.path.join(ROOT, "tests", "golden", "jdacs_output_side.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
