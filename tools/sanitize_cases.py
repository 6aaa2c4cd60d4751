#!/usr/bin/env python
"""Small invocations of every kernel family of libmvs_b200.so, for compute-sanitizer (memcheck / racecheck / synccheck):

  compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
  compute-sanitizer --tool racecheck python tools/sanitize_cases.py

Shapes are tiny (the tools slow kernels down 10-100x) but cover every tap program of the tcgen05 convolution (stride-1
kd-folded / kw-folded, stride-2, transposed stride-2, 2-D 3x3 and 5x5, one and two planes per step), the TMA-staged and the
gathering plane sweep (ragged tiles, per-pixel hypotheses, sources leaving the map), the scatter backward, soft-argmin and the
loss warp.  Results are checked for finiteness only; parity is the job of tests/."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ssmvs_b200  # noqa: E402
from ssmvs_b200 import ops, synth  # noqa: E402

ssmvs_b200._lib.bind()
dev = torch.device("cuda:0")
torch.manual_seed(0)
only = set(sys.argv[1:])


def case(name):
    def deco(fn):
        if not only or name in only:
            fn()
            torch.cuda.synchronize()
            print("ran", name, flush=True)
        return fn
    return deco


@case("conv3d_tc")
def _():
    for cin, cout, stride, tr, shape in ((32, 8, 1, False, (1, 8, 20, 36)), (8, 16, 2, False, (1, 8, 20, 36)), (16, 16, 1, False, (1, 6, 12, 34)),
                                         (64, 32, 2, True, (1, 3, 6, 10)), (16, 8, 2, True, (1, 4, 10, 18)), (8, 1, 1, False, (1, 8, 20, 36)),
                                         (64, 32, 1, True, (1, 4, 8, 12))):
        for dt in (torch.float16, torch.bfloat16):
            b, d, h, w = shape
            x = ops.pack_c8(torch.randn(b, cin, d, h, w, device=dev), dt)
            wt = 0.1 * (torch.randn(cin, cout, 3, 3, 3, device=dev) if tr else torch.randn(cout, cin, 3, 3, 3, device=dev))
            g = ops.pack_conv3d_weight(wt, tr)
            skip = None
            y = ops.conv3d_raw(x, g, cout, stride, tr, torch.ones(cout, device=dev), torch.zeros(cout, device=dev), None, relu=cout > 1, algo=2)
            if cout > 1:
                skip = torch.randn_like(y)
                y = ops.conv3d_raw(x, g, cout, stride, tr, None, None, skip, relu=True, algo=2)
            assert torch.isfinite(y.float()).all()


@case("conv2d_tc")
def _():
    for cin, cout, k, stride in ((8, 8, 3, 1), (8, 16, 5, 2), (16, 16, 3, 1), (16, 32, 5, 2), (32, 32, 3, 1), (64, 64, 3, 1)):
        x = ops.pack_c8(torch.randn(3, cin, 24, 40, device=dev), torch.float16).permute(1, 0, 2, 3, 4).contiguous()
        g = ops.pack_conv2d_weight(0.1 * torch.randn(cout, cin, k, k, device=dev))
        for padded in (False, True):
            y = ops.conv2d_raw(x, g, cout, k, stride, None, torch.zeros(cout, device=dev), True, out_padded=padded)
            assert torch.isfinite(y.float()).all()


@case("warp_var")
def _():
    for c, nsrc, h, w, nd in ((32, 4, 21, 37, 20), (16, 2, 16, 32, 9)):
        inp = synth.feature_inputs(2, nsrc + 1, c, h, w, nd, seed=3)
        P = inp["proj_matrices"].clone()
        P[1, 1, 0, 3] += 5000.0       # one source of item 1 leaves the map (window does not fit: global gather inside the TMA kernel)
        rt = ops.compose_proj(P.to(dev))
        dv = inp["depth_values"].to(dev)
        for dt in (torch.float16, torch.bfloat16):
            maps = ops.pack_c8_padded(inp["features"].flatten(0, 1).to(dev), dt)
            maps = maps.view(nsrc + 1, 2, *maps.shape[1:])
            v = ops.warp_variance_maps(maps, rt, dv, dt)
            dpp = dv.view(2, nd, 1, 1) + torch.randn(2, nd, h, w, device=dev)
            v2 = ops.warp_variance_maps(maps, rt, dpp, dt, False, True)
            assert torch.isfinite(v.float()).all() and torch.isfinite(v2.float()).all()
        f = [t.to(dev).requires_grad_(True) for t in inp["features"]]
        for dt in (torch.float32, torch.bfloat16):
            out = ops.warp_variance(f[0], f[1:], rt, dv, dt)
            out.float().sum().backward()
            assert all(torch.isfinite(t.grad).all() for t in f)


@case("tail")
def _():
    cost = torch.randn(2, 24, 13, 19, device=dev, requires_grad=True)
    dv = (425.0 + 2.5 * torch.arange(24.0, device=dev)).unsqueeze(0).repeat(2, 1)
    d, i, c, p = ops.soft_argmin(cost, dv, want_prob=True)
    d.sum().backward()
    inp = synth.mvsnet_inputs(1, 3, 64, 96, 16, seed=1)
    depth = synth.plausible_depth(1, 16, 24, seed=1).to(dev).requires_grad_(True)
    wimg, m = ops.inverse_warp(torch.randn(1, 16, 24, 3, device=dev), inp["cams"][:, 0].to(dev), inp["cams"][:, 1].to(dev), depth)
    wimg.sum().backward()
    assert torch.isfinite(depth.grad).all() and torch.isfinite(cost.grad).all()


@case("train_convs")
def _():
    from ssmvs_b200.jdacs.models.mvsnet import CostRegNet
    net = CostRegNet().to(dev).train()
    x = torch.randn(1, 32, 8, 16, 24, device=dev, requires_grad=True)
    net(x).sum().backward()
    assert torch.isfinite(x.grad).all()


@case("train_16bit")
def _():
    """The tensor-core training path: 16-bit CostRegNet (tcgen05 forward / dgrad, MMA wgrad, fp64 BN), FeatureNet on the same
    kernels with per-view statistics, the split-by-source sweep backward (2, 4 and 8 lanes per pixel)."""
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    for dt in (torch.bfloat16, torch.float16):
        model = MVSNet(refine=False, train_dtype=dt).to(dev).train()
        inp = {k: v.to(dev) for k, v in synth.mvsnet_inputs(1, 3, 64, 96, 8, seed=2).items()}
        model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])["depth"].sum().backward()
        assert all(p.grad is None or torch.isfinite(p.grad).all() for p in model.parameters())
    for nsrc in (1, 3, 6):
        inp = synth.feature_inputs(1, nsrc + 1, 16, 13, 29, 12, seed=4)
        f = [t.to(dev).requires_grad_(True) for t in inp["features"]]
        rt = ops.compose_proj(inp["proj_matrices"].to(dev))
        ops.warp_variance(f[0], f[1:], rt, inp["depth_values"].to(dev), torch.float16).float().sum().backward()
        assert all(torch.isfinite(t.grad).all() for t in f)


@case("loss_output_fusion")
def _():
    """Fused UnSupLoss forward / backward, the fused FeatureNet front, the output side and the fusion kernel."""
    from ssmvs_b200.jdacs.eval_dense import _pair_cams
    from ssmvs_b200.jdacs.fusion import fusibile as fz
    from ssmvs_b200.jdacs.models.mvsnet import FeatureNet
    inp = synth.mvsnet_inputs(2, 5, 64, 80, 8, seed=3)
    depth = synth.plausible_depth(2, 16, 20, seed=3).to(dev).requires_grad_(True)
    out = ops.unsup_loss(inp["imgs"].to(dev), inp["cams"].to(dev), depth, 1.0, 0.18)
    out[0].backward()
    assert torch.isfinite(out).all() and torch.isfinite(depth.grad).all()
    net = FeatureNet().to(dev).eval()
    with torch.no_grad():
        maps = net.forward_maps(torch.randn(1, 3, 3, 36, 52, device=dev), torch.float16)       # ragged tiles of the fused front
    assert torch.isfinite(maps.float()).all()
    import numpy as np
    k, e0, e1 = synth.intrinsics(40, 32).astype(np.float32), synth.extrinsics(0).astype(np.float32), synth.extrinsics(1).astype(np.float32)
    d = torch.rand(2, 32, 40, device=dev) * 300 + 500
    res = ops.geo_consistency(d, d.flip(0), torch.from_numpy(np.stack([_pair_cams(k, e0, k, e1)] * 2)).to(dev))
    assert torch.isfinite(res[1]).all()
    up = ops.upsample_nearest(d, (75, 100), flip_rows=True)
    assert ops.depth_preview_u8(up).dtype == torch.uint8
    cams = torch.stack([fz.camera_block(k, synth.extrinsics(v)) for v in range(4)]).to(dev)
    pts, valid = fz.fuse_view(fz.constant_normals(torch.rand(4, 32, 40, device=dev) * 50 + 600), cams, 0, None, 0.25, 0.52, 1)
    assert torch.isfinite(pts).all()


print("sanitize_cases: done")
