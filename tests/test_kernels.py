"""Per-kernel parity of the C ABI against the oracle and the golden fixtures.

Every test runs twice: on the host-emulation build of the kernel sources (CPU box, `-m "not gpu"`) and on
libmvs_b200.so on the B200 (`-m gpu`).  Tolerances are for fp32 storage / fp32 accumulation: coordinates are
re-derived (fp64 projection inverse instead of the reference's fp32 LU), so samples agree to ~1e-5 relative."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

TOL = 5e-5


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def be(request):
    return request.getfixturevalue(request.param)


def _ops():
    from ssmvs_b200 import ops
    return ops


def test_c8_roundtrip(be):
    ops = _ops()
    x = torch.randn(2, 24, 3, 5, 7)
    p = ops.pack_c8(be.to(x))
    assert p.shape == (2, 3, 3, 5, 7, 8)
    assert torch.equal(p.cpu()[1, 2, 1, 4, 6], x[1, 16:24, 1, 4, 6])
    assert torch.equal(ops.unpack_c8(p).cpu(), x)


def test_compose_proj(be, oracle, golden):
    ops = _ops()
    g = golden("jdacs_warp")
    rt = ops.compose_proj(be.to(torch.stack([g["ref_proj"], g["src_proj"], g["src_proj2"]], 1))).cpu()
    for s, key in enumerate(("src_proj", "src_proj2")):
        rot, trans = oracle.relative_projection(g[key], g["ref_proj"])
        assert rel_err(rt[s, :, :9].reshape(-1, 3, 3), rot) < 1e-6
        assert rel_err(rt[s, :, 9:], trans.squeeze(-1)) < 1e-6


@pytest.mark.parametrize("align", [False, True])
def test_homo_warp_matches_reference(be, oracle, golden, align):
    ops = _ops()
    g = golden("jdacs_warp")
    rt = ops.compose_proj(be.to(torch.stack([g["ref_proj"], g["src_proj"], g["src_proj2"]], 1)))
    out = ops.homo_warp(be.to(g["src_fea"]), rt[0], be.to(g["depth_values"]), align)
    want = g["warped"] if not align else oracle.homo_warping(g["src_fea"], g["src_proj"], g["ref_proj"], g["depth_values"], True)
    assert rel_err(out, want) < TOL
    far = ops.homo_warp(be.to(g["src_fea"]), rt[1], be.to(g["depth_far"]), align)
    assert torch.allclose(far.cpu(), oracle.homo_warping(g["src_fea"], g["src_proj2"], g["ref_proj"], g["depth_far"], align), atol=1e-5)


def test_homo_warp_degenerate_coordinates(be):
    """z <= 0, NaN and inf depths must give exact zeros (every tap rejected), like grid_sample's zero padding."""
    ops = _ops()
    src = torch.randn(1, 4, 6, 9)
    P = torch.eye(4).reshape(1, 1, 4, 4).repeat(1, 2, 1, 1)
    P[0, 1, 2, 3] = -10.0  # source camera: z = depth - 10
    rt = ops.compose_proj(be.to(P))
    depth = torch.tensor([[10.0, 5.0, float("nan"), float("inf"), 1e30]])
    out = ops.homo_warp(be.to(src), rt[0], be.to(depth)).cpu()
    assert torch.isfinite(out).all()
    assert torch.count_nonzero(out[:, :, 0]) == 0 and torch.count_nonzero(out[:, :, 2]) == 0


@pytest.mark.parametrize("ref_sq,align", [(False, False), (True, False), (False, True)])
def test_warp_variance_forward_backward(be, oracle, golden, ref_sq, align):
    ops = _ops()
    g = golden("jdacs_warp")
    torch.manual_seed(3)
    ref = torch.randn(2, 8, 12, 16)
    srcs = [g["src_fea"], 0.5 * g["src_fea"].flip(0)]
    projs = [g["src_proj"], g["src_proj2"]]
    wt = torch.randn(2, 8, 6, 12, 16)
    rt = ops.compose_proj(be.to(torch.stack([g["ref_proj"]] + projs, 1)))
    a = [be.to(t).detach().clone().requires_grad_(True) for t in [ref] + srcs]
    var = ops.warp_variance(a[0], a[1:], rt, be.to(g["depth_values"]), torch.float32, align, ref_sq)
    (var * ops.pack_c8(be.to(wt))).sum().backward()
    b = [t.detach().clone().requires_grad_(True) for t in [ref] + srcs]
    want = oracle.variance_volume(b[0], b[1:], g["ref_proj"], projs, g["depth_values"], ref_sq, align)
    (want * wt).sum().backward()
    assert rel_err(ops.unpack_c8(var), want) < TOL
    for x, y in zip(a, b):
        assert rel_err(x.grad, y.grad) < TOL


def test_warp_variance_per_pixel_hypotheses(be, golden):
    """proj_cost: per-pixel hypotheses + the CVP variance (hazard H2), against the reference's own output."""
    ops = _ops()
    g = golden("ms_warp")
    rt = ops.compose_proj_ke(*[be.to(g[k]) for k in ("ref_in", "src_in", "ref_ex", "src_ex")], 1.0)
    v = ops.warp_variance(be.to(g["ref_fea"]), [be.to(g["src_fea0"]), be.to(g["src_fea1"])], rt, be.to(g["refine_hypos"]),
                          ref_sq_in_sum=True)
    assert rel_err(ops.unpack_c8(v), g["proj_cost"]) < TOL
    rt1 = ops.compose_proj_ke(*[be.to(g[k]) for k in ("ref_in", "src_in", "ref_ex", "src_ex")], 2.0)
    w = ops.homo_warp(be.to(g["src_fea"]), rt1[0], be.to(g["sweep_hypos"]))
    assert rel_err(w, g["warped"]) < TOL


def test_warp_variance_odd_sizes(be, oracle):
    """Ragged map (odd H, W not a multiple of the block) and a depth count that does not divide the chunking."""
    ops = _ops()
    from ssmvs_b200 import synth
    inp = synth.feature_inputs(1, 3, 16, 7, 13, 5, seed=4)
    f = inp["features"]
    rt = ops.compose_proj(be.to(inp["proj_matrices"]))
    v = ops.warp_variance(be.to(f[0]), [be.to(f[1]), be.to(f[2])], rt, be.to(inp["depth_values"]))
    P = inp["proj_matrices"]
    want = oracle.variance_volume(f[0], [f[1], f[2]], P[:, 0], [P[:, 1], P[:, 2]], inp["depth_values"])
    assert rel_err(ops.unpack_c8(v), want) < TOL


def test_pack_c8_padded_layouts(be):
    """Zero-bordered gather layout from the three source layouts: interior = the map, border = zeros."""
    ops = _ops()
    torch.manual_seed(5)
    x = torch.randn(3, 16, 5, 9)
    want = F.pad(x, (1, 1, 1, 2)).reshape(3, 2, 8, 8, 11).permute(0, 1, 3, 4, 2)
    a = ops.pack_c8_padded(be.to(x))                                                        # fp32 NCHW
    b = ops.pack_c8_padded(be.to(x).contiguous(memory_format=torch.channels_last))          # channels-last
    c = ops.pack_c8_padded(ops.pack_c8(be.to(x)))                                           # C8
    for got in (a, b, c):
        assert got.shape == (3, 2, 8, 11, 8) and torch.equal(got.cpu(), want)


def test_pack_images_c8_and_conv2d_weight(be):
    """Image stack for the 2-D feature layers: [B,N,3,H,W] -> [1, N*B, H, W, 8], image m = v*B + b, channels 3..7 zero; and
    the gather form of a Conv2d weight (tap = kh*k + kw, zero padded to 8 channels)."""
    ops = _ops()
    torch.manual_seed(6)
    imgs = torch.randn(2, 3, 3, 5, 7)
    st = ops.pack_images_c8(be.to(imgs), torch.float32).cpu()
    assert st.shape == (1, 6, 5, 7, 8)
    for v in range(3):
        for b in range(2):
            assert torch.equal(st[0, v * 2 + b, :, :, :3], imgs[b, v].permute(1, 2, 0))
    assert torch.count_nonzero(st[..., 3:]) == 0
    w = torch.randn(16, 3, 5, 5)
    g = ops.pack_conv2d_weight(be.to(w)).cpu()
    assert g.shape == (25, 8, 16)
    assert torch.equal(g[7, :3, :], w[:, :, 1, 2].t()) and torch.count_nonzero(g[:, 3:, :]) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [(torch.float16, 2e-3), (torch.bfloat16, 2e-2)])
@pytest.mark.parametrize("channels,nsrc,ref_sq,per_pixel", [(32, 4, False, False), (16, 2, True, True), (32, 3, False, True),
                                                           (16, 6, False, False), (8, 2, False, False), (64, 1, True, False)])
def test_warp_variance_16bit_padded(gpu, oracle, dtype, tol, channels, nsrc, ref_sq, per_pixel):
    """The production kernel (16-bit storage, zero-bordered maps; 4 / 2 channel blocks per thread; every source-count
    bucket) against the oracle on features rounded to the storage type: agreement to the storage rounding."""
    ops = _ops()
    from ssmvs_b200 import synth
    inp = synth.feature_inputs(2, nsrc + 1, channels, 21, 37, 12, seed=7)
    f = [t.to(dtype).float() for t in inp["features"]]
    P = inp["proj_matrices"]
    depth = inp["depth_values"]
    if per_pixel:
        depth = depth.view(2, -1, 1, 1) + 3.0 * torch.randn(2, depth.shape[1], 21, 37)
    rt = ops.compose_proj(gpu.to(P))
    maps = ops.pack_c8_padded(gpu.to(torch.stack(f, 0).flatten(0, 1)), dtype)
    v = ops.warp_variance_maps(maps.view(nsrc + 1, 2, *maps.shape[1:]), rt, gpu.to(depth), dtype, False, ref_sq)
    want = oracle.variance_volume(f[0], f[1:], P[:, 0], [P[:, s] for s in range(1, nsrc + 1)], depth, ref_sq, False)
    got = ops.unpack_c8(v).cpu()
    assert (got - want).abs().max().item() < tol * want.abs().max().item()
    # the autograd entry takes the same kernel
    v2 = ops.warp_variance(gpu.to(f[0]), [gpu.to(t) for t in f[1:]], rt, gpu.to(depth), dtype, False, ref_sq)
    assert torch.equal(v2, v)


@pytest.mark.gpu
@pytest.mark.parametrize("baseline", [1.0, 4.0, 25.0])
@pytest.mark.parametrize("shape", [(32, 4, 64, 96, 40), (16, 2, 37, 53, 9)])
def test_warp_variance_tma_staging_is_transparent(gpu, knob, shape, baseline):
    """TMA-staged source windows vs the same kernel arithmetic gathering from global memory: bit-identical volumes, whether
    every window fits (baseline 1), some do (4) or none does (25: the per-source fallback inside the staged kernel)."""
    ops = _ops()
    from ssmvs_b200 import synth
    c, nsrc, h, w, nd = shape
    inp = synth.feature_inputs(2, nsrc + 1, c, h, w, nd, seed=11)
    P = inp["proj_matrices"].clone()
    ref_inv = torch.linalg.inv(P[:, 0].double())
    for v in range(1, nsrc + 1):                      # widen the baseline: scale the translation of the relative pose
        rel = P[:, v].double() @ ref_inv
        rel[:, :3, 3] *= baseline
        P[:, v] = (rel @ P[:, 0].double()).float()
    rt = ops.compose_proj(gpu.to(P))
    maps = ops.pack_c8_padded(gpu.to(inp["features"].flatten(0, 1)), torch.float16)
    maps = maps.view(nsrc + 1, 2, *maps.shape[1:])
    dv = gpu.to(inp["depth_values"])
    knob("warp_tma", 1)
    a = ops.warp_variance_maps(maps, rt, dv, torch.float16)
    knob("warp_tma", 0)
    b = ops.warp_variance_maps(maps, rt, dv, torch.float16)
    torch.cuda.synchronize()
    assert torch.equal(a, b)
    assert torch.isfinite(a.float()).all() and a.float().abs().max() > 0


@pytest.mark.gpu
def test_warp_variance_16bit_padded_degenerate(gpu):
    """Planes behind / through the source camera and non-finite depths: finite output, exact zeros where every tap is
    outside the map (variance of (ref, 0) = ref^2 / 4), no out-of-bounds access at the clamped borders."""
    ops = _ops()
    torch.manual_seed(2)
    feats = torch.randn(2, 1, 16, 10, 14)
    P = torch.eye(4).reshape(1, 1, 4, 4).repeat(1, 2, 1, 1)
    P[0, 1, 2, 3] = -10.0  # source camera: z = depth - 10
    P[0, 1, 0, 3] = 3.0
    depth = torch.tensor([[10.0, 5.0, float("nan"), float("inf"), 1e30, 1e-30, 10.0001, 11.0]])
    rt = ops.compose_proj(gpu.to(P))
    maps = ops.pack_c8_padded(gpu.to(feats.flatten(0, 1)), torch.float16)
    v = ops.unpack_c8(ops.warp_variance_maps(maps.view(2, 1, *maps.shape[1:]), rt, gpu.to(depth), torch.float16)).cpu()
    assert torch.isfinite(v).all()
    r = feats[0].half().float()
    for d in (0, 2, 3):
        assert torch.allclose(v[:, :, d], r * r / 4, rtol=2e-3, atol=1e-4)


@pytest.mark.gpu
@pytest.mark.parametrize("tma", [1, 0])
def test_warp_variance_bottom_right_overshoot_reads_only_the_slack_vector(gpu, knob, tma):
    """A source that projects beyond the bottom-right corner clamps to (H+1, W+1): its zero-weight 4th tap is the vector right
    after the last plane.  The C8P contract gives the buffer exactly one zero vector of slack; everything behind it is poisoned
    with NaN here, so a read one vector further (or a missing slack) turns the volume into NaN."""
    ops = _ops()
    torch.manual_seed(4)
    c, h, w, nd = 16, 10, 14, 8
    feats = torch.randn(2, 1, c, h, w)
    P = torch.eye(4).reshape(1, 1, 4, 4).repeat(1, 2, 1, 1)
    P[0, 1, 0, 3] = 1e5   # huge +x, +y translation: every sample lands far beyond the bottom-right corner
    P[0, 1, 1, 3] = 1e5
    depth = (10.0 + torch.arange(nd, dtype=torch.float32)).unsqueeze(0)
    rt = ops.compose_proj(gpu.to(P))
    good = ops.pack_c8_padded(gpu.to(feats.flatten(0, 1)), torch.float16)
    n = good.numel()
    big = torch.full((n + 8 + 4096,), float("nan"), dtype=torch.float16, device=gpu.device)
    big[:n] = good.flatten()
    big[n:n + 8] = 0          # the one slack vector the header asks for
    maps = big[:n].view(2, 1, *good.shape[1:])
    knob("warp_tma", tma)
    v = ops.unpack_c8(ops.warp_variance_maps(maps, rt, gpu.to(depth), torch.float16)).cpu()
    assert torch.isfinite(v).all()
    r = feats[0].half().float()
    assert torch.allclose(v[:, :, 0], r * r / 4, rtol=2e-3, atol=1e-4)     # variance of (ref, 0)
    # per-pixel hypotheses always take the gather kernel
    dpp = depth.view(1, nd, 1, 1).expand(1, nd, h, w).contiguous()
    v2 = ops.unpack_c8(ops.warp_variance_maps(maps, rt, gpu.to(dpp), torch.float16)).cpu()
    assert torch.isfinite(v2).all() and torch.allclose(v2[:, :, 0], r * r / 4, rtol=2e-3, atol=1e-4)
    # and the allocation helper really provides that slack
    st = good.untyped_storage()
    assert st.nbytes() >= (good.storage_offset() + n + 8) * 2
    tail = torch.empty(0, dtype=torch.float16, device=gpu.device).set_(st, good.storage_offset() + n, (8,))
    assert torch.count_nonzero(tail) == 0


def test_soft_argmin(be, oracle, golden):
    ops = _ops()
    g = golden("jdacs_mvsnet")
    cost = be.to(g["cost_reg"]).detach().clone().requires_grad_(True)
    depth, index, conf, prob = ops.soft_argmin(cost, be.to(g["depth_values"]), True)
    assert rel_err(depth, g["depth"]) < 1e-6
    assert rel_err(conf, g["photometric_confidence"]) < 1e-5
    assert torch.equal(index.cpu(), g["index"])                      # bit-exact index on the reference's own cost_reg
    wm = torch.randn(1, 16, 24)
    (depth * be.to(wm)).sum().backward()
    c2 = g["cost_reg"].clone().requires_grad_(True)
    p2, d2 = oracle.soft_argmin(c2, g["depth_values"])
    (d2 * wm).sum().backward()
    assert rel_err(prob, p2) < 1e-5 and rel_err(cost.grad, c2.grad) < 1e-5
    dpp = g["depth_values"].view(1, 8, 1, 1) + torch.randn(1, 8, 16, 24)
    d3, *_ = ops.soft_argmin(be.to(g["cost_reg"]), be.to(dpp))
    assert rel_err(d3, oracle.soft_argmin(g["cost_reg"], dpp)[1]) < 1e-6


def test_soft_argmin_ragged(be, oracle):
    """Depth count not a multiple of 4 and a column count not a multiple of the block: the shared-memory kernel's tails."""
    ops = _ops()
    torch.manual_seed(9)
    cost = 3.0 * torch.randn(2, 13, 7, 9)
    dv = 400.0 + 2.5 * torch.arange(13, dtype=torch.float32).repeat(2, 1)
    depth, index, conf, prob = ops.soft_argmin(be.to(cost), be.to(dv), True)
    p_o, d_o = oracle.soft_argmin(cost, dv)
    i_o, c_o = oracle.photometric_confidence(p_o)
    assert rel_err(depth, d_o) < 1e-6 and rel_err(prob, p_o) < 1e-6 and rel_err(conf, c_o) < 1e-5
    assert (index.cpu() != i_o).sum().item() == 0


def test_soft_argmin_index_exact_on_peaky_columns(be, oracle):
    """One-hot-like columns: expected index lands on integers; truncation must agree with the oracle everywhere."""
    ops = _ops()
    torch.manual_seed(5)
    cost = torch.randn(2, 32, 9, 11)
    peak = torch.randint(0, 32, (2, 9, 11))
    cost.scatter_(1, peak.unsqueeze(1), 60.0)
    dv = (425.0 + 2.65 * torch.arange(32.0)).unsqueeze(0).repeat(2, 1)
    _, index, conf, _ = ops.soft_argmin(be.to(cost), be.to(dv))
    oi, oc = oracle.photometric_confidence(F.softmax(cost, 1))
    assert torch.equal(index.cpu(), oi) and rel_err(conf, oc) < 1e-5


CONVS = [(32, 8, 1, False), (8, 16, 2, False), (16, 16, 1, False), (64, 32, 2, True), (16, 8, 2, True), (8, 1, 1, False),
         (64, 32, 1, True)]


@pytest.mark.parametrize("cin,cout,stride,tr", CONVS)
def test_conv3d_forward_and_gradients(be, cin, cout, stride, tr):
    ops = _ops()
    torch.manual_seed(cin + cout)
    x = torch.randn(2, cin, 4, 6, 8)
    w = 0.1 * (torch.randn(cin, cout, 3, 3, 3) if tr else torch.randn(cout, cin, 3, 3, 3))
    bias = torch.randn(cout) if cout == 1 else None

    def ref(xx, ww, bb):
        return F.conv_transpose3d(xx, ww, bb, stride, 1, stride - 1) if tr else F.conv3d(xx, ww, bb, stride, 1)
    xa, wa = ops.pack_c8(be.to(x)).requires_grad_(True), be.to(w).requires_grad_(True)
    ba = be.to(bias).requires_grad_(True) if bias is not None else None
    ya = ops.conv3d(xa, wa, ba, stride, tr)
    xb, wb = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    bb = bias.clone().requires_grad_(True) if bias is not None else None
    yb = ref(xb, wb, bb)
    wt = torch.randn_like(yb)
    (ya * (be.to(wt.squeeze(1)) if cout == 1 else ops.pack_c8(be.to(wt)))).sum().backward()
    (yb * wt).sum().backward()
    got = ya.detach().unsqueeze(1) if cout == 1 else ops.unpack_c8(ya)
    assert rel_err(got, yb) < 1e-5
    assert rel_err(ops.unpack_c8(xa.grad), xb.grad) < 1e-5
    assert rel_err(wa.grad, wb.grad) < 1e-5
    if bias is not None:
        assert rel_err(ba.grad, bb.grad) < 1e-5


def test_conv3d_fused_epilogue(be):
    """Inference form: relu(conv * scale + shift) + skip in one kernel."""
    ops = _ops()
    torch.manual_seed(1)
    x, w = torch.randn(1, 16, 4, 4, 8), 0.1 * torch.randn(8, 16, 3, 3, 3)
    scale, shift, skip = torch.rand(8) + 0.5, torch.randn(8), torch.randn(1, 8, 4, 4, 8)
    y = ops.conv3d_raw(ops.pack_c8(be.to(x)), ops.pack_conv3d_weight(be.to(w), False), 8, 1, False, be.to(scale), be.to(shift),
                       ops.pack_c8(be.to(skip)), relu=True, algo=1)
    want = F.relu(F.conv3d(x, w, None, 1, 1) * scale.view(1, 8, 1, 1, 1) + shift.view(1, 8, 1, 1, 1)) + skip
    assert rel_err(ops.unpack_c8(y), want) < 1e-5


def test_batch_norm_training(be):
    ops = _ops()
    torch.manual_seed(2)
    bn = torch.nn.BatchNorm3d(16)
    bn.weight.data.uniform_(0.5, 1.5)
    bn.bias.data.normal_()
    bn2 = torch.nn.BatchNorm3d(16)
    bn2.load_state_dict(bn.state_dict())
    bn = bn.to(be.device)
    x, skip, wt = torch.randn(2, 16, 4, 6, 8) * 2 + 0.5, torch.randn(2, 16, 4, 6, 8), torch.randn(2, 16, 4, 6, 8)
    xa, sk = ops.pack_c8(be.to(x)).requires_grad_(True), ops.pack_c8(be.to(skip)).requires_grad_(True)
    ya = ops.bn_act_train(xa, bn, sk, True)
    (ya * ops.pack_c8(be.to(wt))).sum().backward()
    xb, sb = x.clone().requires_grad_(True), skip.clone().requires_grad_(True)
    yb = F.relu(bn2(xb)) + sb
    (yb * wt).sum().backward()
    assert rel_err(ops.unpack_c8(ya), yb) < 1e-5
    assert rel_err(ops.unpack_c8(xa.grad), xb.grad) < 1e-4 and rel_err(ops.unpack_c8(sk.grad), sb.grad) < 1e-6
    assert rel_err(bn.weight.grad, bn2.weight.grad) < 1e-4 and rel_err(bn.bias.grad, bn2.bias.grad) < 1e-4
    assert rel_err(bn.running_mean, bn2.running_mean) < 1e-5 and rel_err(bn.running_var, bn2.running_var) < 1e-5


def test_inverse_warp(be, golden):
    ops = _ops()
    g = golden("jdacs_invwarp")
    d = be.to(g["depth"]).detach().clone().requires_grad_(True)
    w, m = ops.inverse_warp(be.to(g["img"]), be.to(g["cams"][:, 0]), be.to(g["cams"][:, 2]), d)
    (w * be.to(g["weight"])).sum().backward()
    assert rel_err(w, g["warped"]) < TOL and torch.equal(m.cpu(), g["mask"])
    assert rel_err(d.grad, g["grad_depth"]) < TOL
    w, m = ops.inverse_warp(be.to(g["img"]), be.to(g["cams"][:, 0]), be.to(g["cams"][:, 1]), be.to(g["depth_far"]))
    bad = (m.cpu() != g["mask_far"]).sum().item()
    assert bad <= 2                                       # a mask bit may flip only where floor() sits on an integer
    both = (m.cpu() * g["mask_far"])
    assert rel_err(w.cpu() * both, g["warped_far"] * both) < TOL


def test_inverse_warp_image_gradient(be, oracle, golden):
    ops = _ops()
    g = golden("jdacs_invwarp")
    img = be.to(g["img"]).detach().clone().requires_grad_(True)
    w, _ = ops.inverse_warp(img, be.to(g["cams"][:, 0]), be.to(g["cams"][:, 2]), be.to(g["depth"]))
    (w * be.to(g["weight"])).sum().backward()
    i2 = g["img"].clone().requires_grad_(True)
    w2, _ = oracle.inverse_warping(i2, g["cams"][:, 0], g["cams"][:, 2], g["depth"])
    (w2 * g["weight"]).sum().backward()
    assert rel_err(img.grad, i2.grad) < TOL


def test_depth_hypotheses(be, golden):
    ops = _ops()
    g = golden("ms_warp")
    h = ops.depth_hypo_refine(*[be.to(t) for t in (g["depth_up"], g["ref_in"], g["src_in"][:, 0], g["ref_ex"], g["src_ex"][:, 0])])
    assert rel_err(h, g["refine_hypos"]) < 1e-6


def test_bn_passes_of_the_training_path(be):
    """mvs_bn_stats_t / _finalize / _act_fwd_t / _act_bwd_reduce_t / _act_bwd_apply_t (fp64 statistics, any storage type; here fp32
    on both backends) against nn.BatchNorm3d + ReLU + skip autograd, including the running-statistics update."""
    ops = _ops()
    from ssmvs_b200._lib import call, ptr, dtype_code
    torch.manual_seed(3)
    b, c, s = 2, 16, (3, 5, 7)
    z = (torch.randn(b, c, *s) * 2.0 + 5.0).requires_grad_(True)          # |mean| >> std: the cancellation case
    skip = torch.randn(b, c, *s)
    bn = torch.nn.BatchNorm3d(c)
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5); bn.bias.normal_(0, 0.5)
    gamma, beta = be.to(bn.weight.detach()), be.to(bn.bias.detach())
    rm, rv = be.to(bn.running_mean), be.to(bn.running_var)
    y_ref = torch.relu(bn(z)) + skip
    gy = torch.randn_like(y_ref)
    y_ref.backward(gy)
    z8, sk8, gy8 = ops.pack_c8(be.to(z)), ops.pack_c8(be.to(skip)), ops.pack_c8(be.to(gy))
    n = s[0] * s[1] * s[2]
    sums = torch.zeros(2, c, dtype=torch.float64, device=z8.device)
    call("mvs_bn_stats_t", z8, ptr(z8), dtype_code(torch.float32), ptr(sums), b, c, n)
    st = torch.empty(4, c, device=z8.device)
    call("mvs_bn_finalize", z8, ptr(sums), ptr(gamma), ptr(beta), 1e-5, 0.1, float(b * n), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ptr(rm), ptr(rv), c)
    y8 = torch.empty_like(z8)
    call("mvs_bn_act_fwd_t", z8, ptr(z8), ptr(st[0]), ptr(st[1]), ptr(sk8), ptr(y8), dtype_code(torch.float32), b, c, n, 1)
    assert rel_err(ops.unpack_c8(y8), y_ref) < 1e-5
    assert rel_err(rm, bn.running_mean) < 1e-6 and rel_err(rv, bn.running_var) < 1e-5
    red = torch.zeros(2, c, dtype=torch.float64, device=z8.device)
    call("mvs_bn_act_bwd_reduce_t", z8, ptr(z8), ptr(gy8), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ptr(red), dtype_code(torch.float32), b, c, n, 1)
    gz8 = torch.empty_like(z8)
    gp = torch.empty(2, c, device=z8.device)
    call("mvs_bn_act_bwd_apply_t", z8, ptr(z8), ptr(gy8), ptr(st[0]), ptr(st[1]), ptr(st[2]), ptr(st[3]), ptr(red), ptr(gz8), ptr(gp[0]), ptr(gp[1]),
         dtype_code(torch.float32), b, c, n, 1, 0)
    assert rel_err(ops.unpack_c8(gz8), z.grad) < 2e-4
    assert rel_err(gp[0], bn.weight.grad) < 1e-4 and rel_err(gp[1], bn.bias.grad) < 1e-5
    lifted = torch.empty(6, 8, device=z8.device)
    src = be.to(torch.arange(6.0))
    call("mvs_lift_c1", src, ptr(src), ptr(lifted), dtype_code(torch.float32), 6)
    assert torch.equal(lifted[:, 0].cpu(), torch.arange(6.0)) and torch.count_nonzero(lifted[:, 1:]) == 0


def test_random_sweeps_of_the_small_kernels_against_the_oracle(emu, oracle):
    """Property-style sweeps on the host build (random map sizes, batch, plane counts, cameras of every synthetic view, depth
    maps that leave the sources): the refined hypotheses (a9), the loss warp with its mask and d/d depth (a10), and the soft-argmin
    with the truncated expected index bit-exact (a7 / a8, hazard H12)."""
    from hypothesis import given, settings, strategies as st
    from ssmvs_b200 import synth
    ops = _ops()

    @settings(max_examples=30, deadline=None, derandomize=True)
    @given(batch=st.integers(1, 3), h=st.integers(6, 21), w=st.integers(6, 27), view=st.integers(0, 3), seed=st.integers(0, 99),
           scale=st.sampled_from([1.0, 0.7, 1.4]))
    def hypos(batch, h, w, view, seed, scale):
        c = synth.cvp_inputs(batch, 4, h, w, seed=seed)
        depth_up = (synth.plausible_depth(batch, h, w, seed=seed) * scale).contiguous()
        args = (depth_up, c["ref_in"], c["src_in"][:, view].contiguous(), c["ref_ex"], c["src_ex"][:, view].contiguous())
        assert rel_err(ops.depth_hypo_refine(*args), oracle.depth_hypos_refine(*args)) < 1e-6

    @settings(max_examples=30, deadline=None, derandomize=True)
    @given(batch=st.integers(1, 2), h=st.integers(5, 19), w=st.integers(5, 23), view=st.integers(1, 4), seed=st.integers(0, 99),
           scale=st.sampled_from([1.0, 0.5, 0.3, 2.0]), channels=st.sampled_from([1, 3, 4]))
    def invwarp(batch, h, w, view, seed, scale, channels):
        cams = synth.mvsnet_inputs(batch, 5, 4 * h, 4 * w, 8, seed=seed)["cams"]
        depth = (synth.plausible_depth(batch, h, w, seed=seed) * scale).contiguous()
        g = torch.Generator().manual_seed(seed)
        img, weight = torch.randn(batch, h, w, channels, generator=g), torch.randn(batch, h, w, channels, generator=g)
        d0, d1 = depth.clone().requires_grad_(True), depth.clone().requires_grad_(True)
        want, wmask = oracle.inverse_warping(img, cams[:, 0], cams[:, view], d0)
        got, mask = ops.inverse_warp(img, cams[:, 0].contiguous(), cams[:, view].contiguous(), d1)
        flips = mask != wmask                                 # a bit may flip only where floor() sits on an integer
        assert flips.sum().item() <= 2
        keep = (~flips).float()
        assert rel_err(got * keep, want * keep) < TOL
        if want.abs().sum() > 0 and flips.sum().item() == 0:
            (want * weight).sum().backward()
            (got * weight).sum().backward()
            assert rel_err(d1.grad, d0.grad) < 2e-4

    @settings(max_examples=30, deadline=None, derandomize=True)
    @given(batch=st.integers(1, 2), d=st.integers(1, 40), h=st.integers(1, 9), w=st.integers(1, 37), seed=st.integers(0, 99),
           sharp=st.sampled_from([0.0, 1.0, 8.0, 60.0]), per_pixel=st.booleans())
    def argmin(batch, d, h, w, seed, sharp, per_pixel):
        g = torch.Generator().manual_seed(seed)
        cost = torch.randn(batch, d, h, w, generator=g) * sharp
        planes = 425.0 + 2.65 * torch.arange(d, dtype=torch.float32).unsqueeze(0).repeat(batch, 1)
        if per_pixel:
            planes = (planes.view(batch, d, 1, 1) + torch.rand(batch, 1, h, w, generator=g)).contiguous()
        depth, index, conf, _ = ops.soft_argmin(cost, planes)
        prob, want_depth = oracle.soft_argmin(cost, planes)
        want_index, want_conf = oracle.photometric_confidence(prob)
        assert rel_err(depth, want_depth) < 1e-6
        # the expected index is a float sum truncated to an integer: equal wherever the oracle's own sum is not within
        # float rounding of an integer (ascending-order accumulation is part of the contract; ties are hazard H12)
        expect = (prob * torch.arange(d, dtype=torch.float32).view(1, d, 1, 1)).sum(1)
        safe = (expect - expect.round()).abs() > 1e-4
        assert torch.equal(index.long()[safe], want_index.long()[safe])
        assert (conf - want_conf).abs()[safe].max().item() < 1e-5 if safe.any() else True

    hypos()
    invwarp()
    argmin()


def test_random_sweeps_of_the_plane_sweep_against_the_oracle(emu, oracle):
    """The fused warp + variance (a1 / a3 / a4), forward and every gradient, on random shapes: 1-6 source views, 8 / 16 channels,
    ragged maps, plane counts that do not divide the chunking, shared and per-pixel hypotheses, both variance variants (hazard H2),
    both sampling geometries (hazard H1), plane ranges that push the samples out of the sources."""
    from hypothesis import given, settings, strategies as st
    from ssmvs_b200 import synth
    ops = _ops()

    @settings(max_examples=40, deadline=None, derandomize=True)
    @given(batch=st.integers(1, 2), nsrc=st.integers(1, 6), channels=st.sampled_from([8, 16]), h=st.integers(3, 14), w=st.integers(3, 19),
           d=st.integers(1, 11), seed=st.integers(0, 99), ref_sq=st.booleans(), align=st.booleans(), per_pixel=st.booleans(),
           scale=st.sampled_from([1.0, 0.4, 2.5]))
    def check(batch, nsrc, channels, h, w, d, seed, ref_sq, align, per_pixel, scale):
        inp = synth.feature_inputs(batch, nsrc + 1, channels, h, w, d, seed=seed)
        P, planes = inp["proj_matrices"], (inp["depth_values"] * scale).contiguous()
        g = torch.Generator().manual_seed(seed)
        if per_pixel:
            planes = (planes.view(batch, d, 1, 1) + 3.0 * torch.rand(batch, d, h, w, generator=g)).contiguous()
        wt = torch.randn(batch, channels, d, h, w, generator=g)
        a = [t.detach().clone().requires_grad_(True) for t in inp["features"]]
        b = [t.detach().clone().requires_grad_(True) for t in inp["features"]]
        var = ops.warp_variance(a[0], a[1:], ops.compose_proj(P), planes, torch.float32, align, ref_sq)
        want = oracle.variance_volume(b[0], b[1:], P[:, 0], [P[:, v] for v in range(1, nsrc + 1)], planes, ref_sq, align)
        assert rel_err(ops.unpack_c8(var), want) < TOL
        (var * ops.pack_c8(wt)).sum().backward()
        (want * wt).sum().backward()
        for x, y in zip(a, b):
            assert rel_err(x.grad, y.grad) < 2 * TOL

    check()


def test_random_sweeps_of_the_fp32_convolution_against_aten(emu):
    """The fp32 3-D convolution (a5 / a6 layer shapes: stride 1 / 2, transposed with output padding, a single output channel with
    bias) on random extents -- odd D / H / W for stride 1 and the transposed layers, even for stride 2 as the reference needs --
    forward, input gradient, weight gradient against ATen."""
    from hypothesis import given, settings, strategies as st
    ops = _ops()

    @settings(max_examples=25, deadline=None, derandomize=True)
    @given(batch=st.integers(1, 2), cin=st.sampled_from([8, 16, 32]), cout=st.sampled_from([1, 8, 16]), mode=st.sampled_from(["s1", "s2", "t1", "t2"]),
           d=st.integers(1, 5), h=st.integers(1, 7), w=st.integers(1, 9), seed=st.integers(0, 99))
    def check(batch, cin, cout, mode, d, h, w, seed):
        stride, tr = (2 if mode[1] == "2" else 1), mode[0] == "t"
        if stride == 2 and not tr:
            d, h, w = 2 * d, 2 * h, 2 * w
        if cout == 1 and (tr or stride == 2):
            cout = 8                                     # the single-channel layer is `prob`: stride 1, not transposed, biased
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(batch, cin, d, h, w, generator=g)
        wgt = 0.1 * (torch.randn(cin, cout, 3, 3, 3, generator=g) if tr else torch.randn(cout, cin, 3, 3, 3, generator=g))
        bias = torch.randn(cout, generator=g) if cout == 1 else None
        xa, wa = ops.pack_c8(x).requires_grad_(True), wgt.clone().requires_grad_(True)
        ya = ops.conv3d(xa, wa, bias, stride, tr)
        xb, wb = x.clone().requires_grad_(True), wgt.clone().requires_grad_(True)
        yb = F.conv_transpose3d(xb, wb, bias, stride, 1, stride - 1) if tr else F.conv3d(xb, wb, bias, stride, 1)
        wt = torch.randn(yb.shape, generator=g)
        (ya * (wt.squeeze(1) if cout == 1 else ops.pack_c8(wt))).sum().backward()
        (yb * wt).sum().backward()
        assert rel_err(ya.detach().unsqueeze(1) if cout == 1 else ops.unpack_c8(ya), yb) < 1e-5
        assert rel_err(ops.unpack_c8(xa.grad), xb.grad) < 1e-5 and rel_err(wa.grad, wb.grad) < 1e-5

    check()
