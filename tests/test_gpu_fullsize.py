"""Full-size (BASELINE.json configs[1]: N=5, 128x160 maps, C=32, D=192; configs[2]: CVP-MVSNet 3 levels) checks on the B200.

End-to-end parity runs the ORACLE itself on the GPU in fp32 (TF32 off) at the full size, on probability volumes that are
asserted to be peaky (tests/parity_util.py; hazard H11), and compares the fp32, fp16 and bf16 product paths with it: relative
depth error (north_star: <= 1e-3), expected-plane index mismatches (H12), confidence.  Stage checks use size-independent
properties and plain PyTorch fp32 GPU ops of the same arithmetic (ATen grid_sample, conv3d, softmax)."""
import json
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu


def _record(name, doc):
    """Keep the measured numbers next to the other GPU artefacts (gpurun_out/ travels back from the box)."""
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    try:
        os.makedirs(out, exist_ok=True)
        with open(os.path.join(out, "parity_%s.json" % name), "w") as f:
            json.dump(doc, f, indent=1)
    except OSError:
        pass


@pytest.fixture
def cfg2(gpu):
    from ssmvs_b200 import ops, synth
    inp = synth.feature_inputs(1, 5, 32, 128, 160, 192, seed=0)
    feats = [f.to(gpu.device) for f in inp["features"]]
    rt = ops.compose_proj(inp["proj_matrices"].to(gpu.device))
    return {"feats": feats, "rt": rt, "dv": inp["depth_values"].to(gpu.device), "proj": inp["proj_matrices"].to(gpu.device)}


def _torch_variance(feats, proj, dv):
    """The reference's op sequence on the GPU (grid_sample + elementwise), jdacs/models/mvsnet.py:120-136."""
    b, c, h, w = feats[0].shape
    d = dv.shape[1]
    ref = feats[0].unsqueeze(2).expand(b, c, d, h, w)
    s1, s2 = ref.clone(), ref ** 2
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dv.device), torch.arange(w, dtype=torch.float32, device=dv.device), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dv.device)))
    for s in range(1, len(feats)):
        rel = proj[:, s] @ torch.inverse(proj[:, 0])
        ray = rel[:, :3, :3] @ pix.unsqueeze(0)
        pts = ray.unsqueeze(2) * dv.view(b, 1, d, 1) + rel[:, :3, 3].view(b, 3, 1, 1)
        xy = pts[:, :2] / pts[:, 2:3]
        grid = torch.stack((xy[:, 0] / ((w - 1) / 2) - 1, xy[:, 1] / ((h - 1) / 2) - 1), dim=3).view(b, d * h, w, 2)
        wv = F.grid_sample(feats[s], grid, mode="bilinear", padding_mode="zeros", align_corners=False).view(b, c, d, h, w)
        s1 += wv
        s2 += wv ** 2
    n = len(feats)
    return s2 / n - (s1 / n) ** 2


def test_variance_volume_matches_aten_at_full_size(cfg2):
    from ssmvs_b200 import ops
    want = _torch_variance(cfg2["feats"], cfg2["proj"], cfg2["dv"])
    v32 = ops.unpack_c8(ops.warp_variance(cfg2["feats"][0], cfg2["feats"][1:], cfg2["rt"], cfg2["dv"], torch.float32))
    assert rel_err(v32, want) < 1e-4
    scale = want.abs().max().item()
    for dt, tol in ((torch.float16, 2e-3), (torch.bfloat16, 2e-2)):   # storage rounding of inputs and output only
        v = ops.unpack_c8(ops.warp_variance(cfg2["feats"][0], cfg2["feats"][1:], cfg2["rt"], cfg2["dv"], dt))
        assert (v - want).abs().max().item() < tol * scale, dt


def test_variance_properties(cfg2):
    from ssmvs_b200 import ops
    f, rt, dv = cfg2["feats"], cfg2["rt"], cfg2["dv"]
    v = ops.warp_variance(f[0], f[1:], rt, dv, torch.float32)
    assert torch.isfinite(v).all() and v.min().item() > -1e-4            # a variance (fp32 cancellation only)
    v2 = ops.warp_variance(2 * f[0], [2 * s for s in f[1:]], rt, dv, torch.float32)
    assert torch.equal(v2, 4 * v)                                          # exact: scaling by 2 commutes with fp32 rounding
    perm = [2, 0, 3, 1]
    vp = ops.warp_variance(f[0], [f[1 + i] for i in perm], rt[perm].contiguous(), dv, torch.float32)
    assert rel_err(vp, v) < 1e-5                                           # symmetric in the source views
    eye = ops.compose_proj(torch.eye(4, device=dv.device).expand(1, 2, 4, 4).contiguous())
    same = ops.warp_variance(f[0], [f[0]], eye, dv, torch.float32, align_corners=True)
    assert same.abs().max().item() < 1e-5    # identical views at identity pose: zero variance (align_corners=True geometry;
    shifted = ops.warp_variance(f[0], [f[0]], eye, dv, torch.float32, align_corners=False)
    assert shifted.abs().max().item() > 1e-2  # with the torch>=1.3 default the same pair is sampled half a pixel off: hazard H1)


def test_regularisation_matches_cudnn_at_full_size(gpu):
    """CostRegNet (our kernels) vs the same state_dict run through ATen/cuDNN conv3d in fp32."""
    from ssmvs_b200 import ops
    from ssmvs_b200.jdacs.models.mvsnet import CostRegNet
    torch.manual_seed(0)
    net = CostRegNet().to(gpu.device).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        x = torch.randn(1, 32, 48, 64, 80, device=gpu.device).abs()
        sd = {k: v for k, v in net.state_dict().items()}

        def cbr(t, name, stride):
            y = F.conv3d(t, sd[name + ".conv.weight"], None, stride, 1)
            return F.relu(F.batch_norm(y, sd[name + ".bn.running_mean"], sd[name + ".bn.running_var"], sd[name + ".bn.weight"], sd[name + ".bn.bias"], False, 0.1, 1e-5))

        def dbr(t, name):
            y = F.conv_transpose3d(t, sd[name + ".0.weight"], None, 2, 1, 1)
            return F.relu(F.batch_norm(y, sd[name + ".1.running_mean"], sd[name + ".1.running_var"], sd[name + ".1.weight"], sd[name + ".1.bias"], False, 0.1, 1e-5))
        prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
        torch.backends.cudnn.allow_tf32 = False
        c0 = cbr(x, "conv0", 1); c2 = cbr(cbr(c0, "conv1", 2), "conv2", 1); c4 = cbr(cbr(c2, "conv3", 2), "conv4", 1)
        y = cbr(cbr(c4, "conv5", 2), "conv6", 1)
        y = c4 + dbr(y, "conv7"); y = c2 + dbr(y, "conv9"); y = c0 + dbr(y, "conv11")
        want = F.conv3d(y, sd["prob.weight"], sd["prob.bias"], 1, 1).squeeze(1)
        torch.backends.cudnn.allow_tf32 = prev[0]
        net.algo = 1
        got = net(ops.pack_c8(x))
        assert rel_err(got, want) < 1e-4
        net.algo = 0
        for dt, tol in ((torch.float16, 1e-2), (torch.bfloat16, 5e-2)):
            net.act_dtype = dt
            got = net(ops.pack_c8(x, dt))
            assert (got - want).abs().max().item() < tol * want.abs().max().item(), dt


def test_softargmin_matches_aten_at_full_size(gpu):
    from ssmvs_b200 import ops
    torch.manual_seed(1)
    cost = 4.0 * torch.randn(1, 192, 128, 160, device=gpu.device)
    dv = (425.0 + 2.65 * torch.arange(192.0, device=gpu.device)).unsqueeze(0)
    depth, index, conf, _ = ops.soft_argmin(cost, dv)
    p = F.softmax(cost, 1)
    want = (p * dv.view(1, 192, 1, 1)).sum(1)
    widx = (p * torch.arange(192.0, device=gpu.device).view(1, 192, 1, 1)).sum(1)
    assert rel_err(depth, want) < 1e-5
    # the index is a truncated float sum (hazard H12): it may differ only where the sum sits within 1e-4 of an integer
    diff = index != widx.long()
    assert (diff & ((widx - widx.round()).abs() > 1e-4)).sum().item() == 0
    assert diff.float().mean().item() < 1e-3
    win = 4 * F.avg_pool3d(F.pad(p.unsqueeze(1), (0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1).squeeze(1)
    assert rel_err(conf[~diff], torch.gather(win, 1, index.unsqueeze(1)).squeeze(1)[~diff]) < 1e-5


# Tolerances of the end-to-end comparisons below (measured values in profiles/r02_parity.json).  north_star asks <= 1e-3
# relative on depth.  The fp32 path meets it everywhere (measured 6e-6).  The 16-bit paths sit exactly on the floor set by 16-bit
# STORAGE on this ill-conditioned input (random-weight network on noise images, multi-modal peaky softmax): the oracle's own fp32
# arithmetic with every stored tensor rounded once to fp16 / bf16 (parity_util.ideal_storage_mvsnet) shows the same errors, and
# the oracle under torch's default TF32 convolutions -- what the unmodified reference does on this GPU -- is slightly further
# from the fp32 oracle than the fp16 product path.  So the asserted bounds are (i) absolute caps a little above the measured
# numbers and (ii) "no worse than 1.25x the storage floor / the reference's own TF32 spread".
TOL = {"fp32": {"max": 1e-4}, "fp16": {"p999": 2e-3, "max": 4e-3, "within": 0.97}, "bf16": {"p999": 2e-2, "max": 4e-2}, "vs_floor": 1.25}


def test_config2_end_to_end_parity_on_a_peaky_volume(gpu):
    """BASELINE configs[1] (N=5, 512x640, D=192), whole MVSNet.forward, product path in fp32 / fp16 / bf16 storage vs the fp32
    oracle run on the GPU.  The probability volume is peaky by construction and that is asserted (hazard H11)."""
    import parity_util as pu
    model, inp, want, cond = pu.peaky_mvsnet(gpu.device, 5, 512, 640, 192, seed=0, target_peak=0.3)
    assert cond["peak"] >= 0.3 and cond["depth_std"] >= 10.0, cond         # a uniform softmax would make this test vacuous
    assert 425.0 < cond["depth_min"] < cond["depth_max"] < 935.0
    doc = {"conditions": cond, "product": {}, "ideal_16bit_storage": {}}
    r32 = pu.depth_parity(pu.product_mvsnet(model, inp, torch.float32), want)
    doc["product"]["fp32"] = r32
    for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
        got = pu.product_mvsnet(model, inp, dt)
        doc["product"][name] = pu.depth_parity(got, want)
        doc["ideal_16bit_storage"][name] = pu.depth_parity(pu.ideal_storage_mvsnet(model, inp, dt), want)
    doc["oracle_tf32_default"] = pu.depth_parity(pu.oracle_mvsnet(model, inp, tf32=True)[0], want)
    _record("config2", doc)
    assert r32["depth_rel_max"] <= TOL["fp32"]["max"], r32
    # fp32 storage: the expected-plane index is a truncated float sum (H12): a handful of pixels whose sum sits next to an integer
    # may flip by one plane with the summation order; nothing else may differ
    assert r32["index_mismatch"] <= 5 and r32["index_off_by_more_than_1"] == 0, r32
    assert r32["conf_abs_p999"] <= 1e-3, r32
    tf = doc["oracle_tf32_default"]
    for name in ("fp16", "bf16"):
        r, ideal = doc["product"][name], doc["ideal_16bit_storage"][name]
        assert r["depth_rel_p999"] <= TOL[name]["p999"] and r["depth_rel_max"] <= TOL[name]["max"], (name, r)
        assert r["depth_rel_p999"] <= TOL["vs_floor"] * ideal["depth_rel_p999"], (name, r, ideal)     # the kernels add nothing to the storage floor
    r = doc["product"]["fp16"]
    assert r["frac_within_1e-3"] >= TOL["fp16"]["within"] and r["index_off_by_more_than_1"] == 0, r
    assert r["depth_rel_p999"] <= TOL["vs_floor"] * tf["depth_rel_p999"], (r, tf)     # no further from fp32 than the reference's own TF32 default
    assert r["index_mismatch"] <= TOL["vs_floor"] * tf["index_mismatch"], (r, tf)


def test_feature_net_folded_fast_path(gpu):
    """Library-side eval fast path (BN folded, fp16 channels-last) against the plain fp32 FeatureNet, and its hand-off
    to the sweep (channels-last -> C8 repack kernel)."""
    from ssmvs_b200 import ops
    from ssmvs_b200.jdacs.models.mvsnet import FeatureNet
    torch.manual_seed(0)
    net = FeatureNet().to(gpu.device).eval()
    with torch.no_grad():
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
        x = torch.randn(3, 3, 128, 160, device=gpu.device)
        want = net(x)
        got = net.forward_folded(x, torch.float16)
        assert got.dtype == torch.float16 and got.is_contiguous(memory_format=torch.channels_last)
        assert (got.float() - want).abs().max().item() < 2e-2 * want.abs().max().item()
        assert torch.equal(ops.unpack_c8(ops.pack_c8(got, torch.float16)), got.float())


def test_cvpmvsnet_16bit_parity_small(gpu):
    """CVP-MVSNet (coarse sweep + one per-pixel refinement level) at 128x160: fp32 / fp16 product paths vs the fp32 oracle on a
    peaky coarse volume.  The CVP regulariser needs a last-layer gain of ~2000 to become peaky on random weights, which amplifies
    every rounding: the fp16 bound is therefore stated against the oracle's own spread under torch's default TF32 convolutions."""
    import parity_util as pu
    model, inp, want, cond = pu.peaky_cvp(gpu.device, 3, 2, 128, 160, seed=3, target_peak=0.3)
    assert cond["peak"] >= 0.3 and cond["depth_std"] >= 10.0, cond
    r32 = pu.cvp_parity(pu.product_cvp(model, inp, torch.float32), want)
    r16 = pu.cvp_parity(pu.product_cvp(model, inp, torch.float16), want)
    tf = pu.cvp_parity(pu.oracle_cvp(model, inp, 2, tf32=True)[0], want)
    _record("cvp_small", {"conditions": cond, "fp32": r32, "fp16": r16, "oracle_tf32_default": tf})
    for lvl in ("level0", "level1"):
        assert r32[lvl]["depth_rel_max"] <= 2e-4, r32
        assert r16[lvl]["depth_rel_p999"] <= 3e-2 and r16[lvl]["depth_rel_p999"] <= 1.5 * tf[lvl]["depth_rel_p999"], (r16, tf)
    assert r32["conf_abs_max"] <= 2e-3


def test_config3_cvp_three_stage_end_to_end_parity(gpu):
    """BASELINE configs[2]: CVP-MVSNet, 3 pyramid levels, 1 + 4 views of 512x640: fp32, fp16 and bf16 product paths against the
    fp32 oracle run on the GPU (cvp_forward), coarse probability volume asserted peaky.  Finer levels build their 8 hypotheses
    around the coarser level's own depth, so their numbers include the propagated differences.  fp32 storage meets north_star's
    1e-3 with two orders of margin; fp16 matches the oracle's own TF32-default spread; bf16 (8 mantissa bits under a last-layer
    gain of 2048) is reported and only loosely bounded."""
    import parity_util as pu
    model, inp, want, cond = pu.peaky_cvp(gpu.device, 4, 3, 512, 640, seed=5, target_peak=0.3)
    assert cond["peak"] >= 0.3 and cond["depth_std"] >= 10.0, cond
    doc = {"conditions": cond, "product": {}}
    for name, dt in (("fp32", torch.float32), ("fp16", torch.float16), ("bf16", torch.bfloat16)):
        got = pu.product_cvp(model, inp, dt)
        assert [tuple(t.shape) for t in got["depth_est_list"]] == [(1, 512, 640), (1, 256, 320), (1, 128, 160)]
        assert all(torch.isfinite(t).all() for t in got["depth_est_list"]) and got["conf"].shape == (1, 512, 640)
        doc["product"][name] = pu.cvp_parity(got, want)
    tf = doc["oracle_tf32_default"] = pu.cvp_parity(pu.oracle_cvp(model, inp, 3, tf32=True)[0], want)
    _record("config3", doc)
    for lvl in ("level0", "level1", "level2"):
        assert doc["product"]["fp32"][lvl]["depth_rel_max"] <= 1e-4, doc["product"]["fp32"]
        r16 = doc["product"]["fp16"][lvl]
        assert r16["depth_rel_p999"] <= 3e-2 and r16["depth_rel_p999"] <= 1.5 * tf[lvl]["depth_rel_p999"], (r16, tf[lvl])
        rb = doc["product"]["bf16"][lvl]
        assert rb["depth_rel_p999"] <= 0.25 and rb["depth_rel_median"] <= 2e-2, rb


def test_config4_train_step_full_size(gpu):
    """BASELINE configs[3]: one self-supervised training step (photometric loss; N = 5 views because the reference's top-3 view
    selection needs >= 3 sources, hazard H5) at 512x640, D = 192: finite loss, gradients on every trainable parameter, an
    Adam step that changes the weights.  (The co-segmentation term needs pretrained VGG weights: out of scope)"""
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    torch.manual_seed(0)
    model = MVSNet(refine=False).to(gpu.device).train()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    inp = gpu.to(synth.mvsnet_inputs(1, 5, 512, 640, 192, seed=2))
    before = model.cost_regularization.conv0.conv.weight.detach().clone()
    out = model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])
    loss = UnSupLoss()(inp["imgs"], inp["cams"], out["depth"])
    opt.zero_grad()
    loss.backward()
    assert torch.isfinite(loss)
    missing = [n for n, p in model.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing
    assert all(torch.isfinite(p.grad).all() for p in model.parameters() if p.grad is not None)
    opt.step()
    assert not torch.equal(before, model.cost_regularization.conv0.conv.weight.detach())


def test_streamed_forward_pipeline_matches_direct_calls(gpu):
    """graph.StreamedForward (two captured graphs, uploads one batch ahead, one forward in flight behind the result handed out):
    every batch's host result equals the direct module call on that batch, in order, for an odd and an even number of batches."""
    from ssmvs_b200 import synth
    from ssmvs_b200.graph import StreamedForward
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    torch.manual_seed(0)
    model = MVSNet(refine=False, volume_dtype=torch.float16).eval().to(gpu.device)
    with torch.no_grad():
        model.cost_regularization.prob.weight.mul_(64.0)
    batches = []
    for seed in range(5):
        inp = synth.mvsnet_inputs(2, 3, 64, 96, 16, seed=seed)
        batches.append([inp[k].pin_memory() for k in ("imgs", "proj_matrices", "depth_values")])
    with torch.no_grad():
        want = [model(*[t.to(gpu.device) for t in b]) for b in batches]
        pipe = StreamedForward(lambda i, pm, dv: model(i, pm, dv), [t.to(gpu.device) for t in batches[0]], ("depth", "photometric_confidence"))
        for n in (5, 2, 1):
            got = [{k: v.clone() for k, v in out.items()} for out in pipe.run(batches[:n])]
            assert len(got) == n
            for g, w in zip(got, want):
                assert torch.equal(g["depth"], w["depth"].cpu()) and torch.equal(g["photometric_confidence"], w["photometric_confidence"].cpu())


def test_fused_loss_at_full_size_against_the_oracle(gpu, oracle):
    """UnSupLoss at the headline size (N = 5 views of 512x640, 128x160 depth map) against oracle.unsup_loss on the CPU: the four
    scalars and d/d depth (the fp64 block-reduced sums vs torch's fp32 reductions: 1e-5 / 1e-4 as on the small fixtures)."""
    from ssmvs_b200 import ops, synth
    inp = synth.mvsnet_inputs(1, 5, 512, 640, 8, seed=4)
    depth = synth.plausible_depth(1, 128, 160, seed=4)
    d_ref = depth.clone().requires_grad_(True)
    want = oracle.unsup_loss(inp["imgs"], inp["cams"], d_ref, True, 0.18)
    want["total"].backward()
    d = depth.to(gpu.device).requires_grad_(True)
    out = ops.unsup_loss(inp["imgs"].to(gpu.device), inp["cams"].to(gpu.device), d, 1.0, 0.18)
    out[0].backward()
    for i, k in enumerate(("total", "reconstr", "ssim", "smooth")):
        assert rel_err(out[i], want[k]) < 1e-5, k
    assert rel_err(d.grad, d_ref.grad) < 2e-4


def test_geometric_filter_at_output_size_against_the_oracle(gpu):
    """mvs_geo_consistency on 1200 x 1600 maps (the size eval_dense.py writes) against the NumPy restatement that the golden
    fixture pins to the reference's own functions: identical masks, identical coordinates."""
    import importlib.util
    import numpy as np
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs import eval_dense as ed
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("output_side_oracle", os.path.join(root, "oracle", "output_side.py"))
    side = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(side)
    h, w = 1200, 1600
    k = synth.intrinsics(w, h).astype(np.float32)
    e0, e1 = synth.extrinsics(0).astype(np.float32), synth.extrinsics(1).astype(np.float32)
    g = np.random.default_rng(0)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    d_ref = (620 + 40 * np.sin(xx / 90.0) * np.cos(yy / 70.0)).astype(np.float32)
    d_src = (d_ref + g.normal(0, 1.5, (h, w))).astype(np.float32)
    d_src[300:340] = 0
    want = side.check_geometric_consistency(d_ref, k, e0, d_src, k, e1)
    got = ed.check_geometric_consistency(torch.from_numpy(d_ref).to(gpu.device), k, e0, torch.from_numpy(d_src).to(gpu.device), k, e1)
    mism = int((got[0].cpu().numpy() != want[0]).sum())
    assert mism <= 2, mism                               # a threshold decision may sit on the last bit of an fp64 product
    assert 0.02 < want[0].mean() < 0.98
    same = got[0].cpu().numpy() & want[0]
    assert np.allclose(got[1].cpu().numpy()[same], want[1][same], rtol=1e-6, atol=1e-4)
    assert np.allclose(got[2].cpu().numpy(), want[2], rtol=1e-6, atol=1e-3, equal_nan=True)
