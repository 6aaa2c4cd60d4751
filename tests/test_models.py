"""The reference-facing modules (same names / signatures / state-dict keys) against outputs of the reference itself."""
from types import SimpleNamespace

import pytest
import torch

from conftest import rel_err, state_dict_of


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def be(request):
    return request.getfixturevalue(request.param)


def test_mvsnet_eval_and_train(be, golden):
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    g = golden("jdacs_mvsnet")
    model = MVSNet(refine=False, volume_dtype=torch.float32, train_dtype=torch.float32)     # the reference's arithmetic
    res = model.load_state_dict(state_dict_of(g), strict=False)
    assert not res.unexpected_keys and all("num_batches_tracked" in k for k in res.missing_keys)
    model = model.to(be.device).eval()
    args = [be.to(g[k]) for k in ("imgs", "proj_matrices", "depth_values")]
    with torch.no_grad():
        out = model(*args)
    assert set(out) == {"depth", "photometric_confidence"}
    assert rel_err(out["depth"], g["depth"]) < 1e-4            # north_star: <= 1e-3 relative on depth maps
    assert rel_err(out["photometric_confidence"], g["photometric_confidence"]) < 1e-3
    model.train()
    out = model(*args)
    assert rel_err(out["depth"], g["train_depth"]) < 1e-4
    (out["depth"] * be.to(g["loss_weight"])).sum().backward()
    params = dict(model.named_parameters())
    for k, v in g.items():
        if k.startswith("grad."):
            assert rel_err(params[k[5:]].grad, v) < 2e-3, k


def test_mvsnet_stage_taps(be, golden):
    """Variance volume and cost_reg of the eval forward, stage by stage, against the reference's tensors."""
    from ssmvs_b200 import ops
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    g = golden("jdacs_mvsnet")
    model = MVSNet(refine=False)
    model.load_state_dict(state_dict_of(g), strict=False)
    model = model.to(be.device).eval()
    with torch.no_grad():
        feats = be.to(g["features"])
        rt = ops.compose_proj(be.to(g["proj_matrices"]))
        var = ops.warp_variance(feats[0], [feats[1], feats[2]], rt, be.to(g["depth_values"]))
        assert rel_err(ops.unpack_c8(var), g["variance"]) < 5e-5
        reg = model.cost_regularization(ops.pack_c8(be.to(g["variance"])))
        assert rel_err(reg, g["cost_reg"]) < 1e-4
        assert rel_err(model.cost_regularization(be.to(g["variance"])).squeeze(1), g["cost_reg"]) < 1e-4  # reference-shaped input
        _, index, _, _ = ops.soft_argmin(reg, be.to(g["depth_values"]))
        assert (index.cpu() != g["index"]).sum().item() == 0


def test_state_dict_keys_match_reference(golden):
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet
    mine = {k for k in MVSNet(refine=False).state_dict() if "num_batches_tracked" not in k}
    assert mine == set(state_dict_of(golden("jdacs_mvsnet")))
    mine = {k for k in CVPMVSNet(SimpleNamespace(nsrc=2, nscale=2, mode="test")).state_dict() if "num_batches_tracked" not in k}
    assert mine == set(state_dict_of(golden("ms_cvp")))
    assert {k for k in MVSNet(refine=True).state_dict()} >= {"refine_network.res.conv.weight"}


def test_cvpmvsnet(be, golden):
    from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet
    g = golden("ms_cvp")
    model = CVPMVSNet(SimpleNamespace(nsrc=2, nscale=2, mode="test"), volume_dtype=torch.float32, train_dtype=torch.float32)
    model.load_state_dict(state_dict_of(g), strict=False)
    model = model.to(be.device).eval()
    args = [be.to(g[k]) for k in ("ref_img", "src_imgs", "ref_in", "src_in", "ref_ex", "src_ex", "depth_min", "depth_max")]
    with torch.no_grad():
        out = model(*args)
    assert len(out["depth_est_list"]) == 2 and out["depth_est_list"][0].shape == (1, 32, 48)
    for i, d in enumerate(out["depth_est_list"]):
        assert rel_err(d, g["depth_est_list.%d" % i]) < 1e-4
    assert rel_err(out["prob_confidence"], g["prob_confidence"]) < 1e-3
    model.train()
    out = model(*args)
    sum(d.sum() for d in out["depth_est_list"]).backward()
    assert model.cost_reg_refine.conv0.conv.weight.grad.abs().sum() > 0
    assert model.featurePyramid.conv0aa[0].weight.grad.abs().sum() > 0


def test_unsup_loss(be, golden):
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs_ms.losses.unsup_loss import UnSupLoss as UnSupLossMS
    for name, cls in (("jdacs_unsup_loss", UnSupLoss), ("ms_unsup_loss", UnSupLossMS)):
        g = golden(name)
        d = be.to(g["depth"]).detach().clone().requires_grad_(True)
        crit = cls()
        total = crit(be.to(g["imgs"]), be.to(g["cams"]), d)
        total.backward()
        assert rel_err(total, g["total"]) < 1e-5
        assert rel_err(crit.reconstr_loss, g["reconstr"]) < 1e-5 and rel_err(crit.ssim_loss, g["ssim"]) < 1e-5
        assert rel_err(crit.smooth_loss, g["smooth"]) < 1e-5
        assert rel_err(d.grad, g["grad_depth"]) < 1e-4
    with pytest.raises(RuntimeError):                          # hazard H5: top-3 needs >= 3 source views
        g = golden("jdacs_unsup_loss")
        UnSupLoss()(be.to(g["imgs"][:, :3]), be.to(g["cams"][:, :3]), be.to(g["depth"]))


def _fake_replica(module):
    """What nn.parallel.replicate() hands a DataParallel worker under no_grad: a shallow copy of every module (shared __dict__
    entries such as the pack caches!) whose parameters are plain tensor copies, not nn.Parameters."""
    memo = {}
    for name, m in module.named_modules():
        r = m._replicate_for_data_parallel()
        memo[m] = r
    for m, r in memo.items():
        for k, child in m._modules.items():
            r._modules[k] = memo[child] if child is not None else None
        for k, p in m._parameters.items():
            r._parameters[k] = None if p is None else p.detach().clone()
        for k, b in m._buffers.items():
            r._buffers[k] = None if b is None else b.detach().clone()
    return memo[module]


def test_replicas_never_populate_or_hit_the_pack_caches(emu, golden):
    """DataParallel replicas share the original's cache objects and hold fresh tensors with recycled ids / addresses /
    _version 0, so an entry cached for one replica could be served to a later replica with different weights."""
    from ssmvs_b200 import regnet
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    g = golden("jdacs_mvsnet")
    model = MVSNet(refine=False)
    model.load_state_dict(state_dict_of(g), strict=False)
    model.eval()
    args = [emu.to(g[k]) for k in ("imgs", "proj_matrices", "depth_values")]
    with torch.no_grad():
        rep = _fake_replica(model)
        assert regnet.owned(model) and not regnet.owned(rep) and rep.cost_regularization._cache is model.cost_regularization._cache
        a = rep(*args)["depth"]
        assert not model.cost_regularization._cache._store and not model.cost_regularization.conv0._cache._store
        assert rel_err(a, g["depth"]) < 1e-4
        # "an epoch later": new weights, a new replica -- must see the new weights
        model.cost_regularization.prob.weight.mul_(0.5)
        b = _fake_replica(model)(*args)["depth"]
        want = model(*args)["depth"]
        assert torch.equal(b, want) and not torch.equal(a, b)
        assert model.cost_regularization._cache._store            # the owner itself does cache ...
        model.train()
        assert not model.cost_regularization._cache._store        # ... and train() drops everything derived from the weights


def test_mvsnet_degenerate_batches(be):
    """Inputs at the edge of the interface: an empty batch is an empty result (the kernels are not launched on it), a lone
    reference view is refused with a message (the reference runs on to a depth map that ignores the images), D / H / W that the
    3-D U-Net cannot halve three times are refused as the reference's own layers refuse them."""
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    inp = be.to(synth.mvsnet_inputs(1, 3, 64, 96, 16, seed=0))
    model = MVSNet(refine=False).to(be.device).eval()
    model.keep_index = True
    with torch.no_grad():
        out = model(inp["imgs"][:0], inp["proj_matrices"][:0], inp["depth_values"][:0])
        assert out["depth"].shape == (0, 16, 24) and out["photometric_confidence"].shape == (0, 16, 24)
        assert out["depth_index"].dtype == torch.int64 and out["depth"].device.type == be.device.type
        with pytest.raises(ValueError, match="source view"):
            model(inp["imgs"][:, :1], inp["proj_matrices"][:, :1], inp["depth_values"])
        with pytest.raises(ValueError, match="divisible by 8"):
            model(inp["imgs"], inp["proj_matrices"], inp["depth_values"][:, :12])
        with pytest.raises(ValueError, match="divisible by 8"):
            model(inp["imgs"][..., :60, :].contiguous(), inp["proj_matrices"], inp["depth_values"])
        with pytest.raises(AssertionError):                      # mvsnet.py:108
            model(inp["imgs"], inp["proj_matrices"][:, :2], inp["depth_values"])


def test_model_forward_random_configurations_against_the_oracle(emu, oracle):
    """Property-style sweep at module level on the host build: MVSNet.eval() over random batch / view counts / image sizes / plane
    counts and CVPMVSNet over random source counts / pyramid depths, BatchNorm statistics randomised so that the softmax is not
    flat (hazard H11), against the oracle's whole-forward restatements with the same state dict."""
    from hypothesis import given, settings, strategies as st
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet

    @settings(max_examples=14, deadline=None, derandomize=True)
    @given(batch=st.integers(1, 2), views=st.integers(2, 5), hb=st.integers(1, 2), wb=st.integers(1, 3), db=st.integers(1, 3), seed=st.integers(0, 20))
    def mvsnet(batch, views, hb, wb, db, seed):
        torch.manual_seed(seed)
        model = MVSNet(refine=False, volume_dtype=torch.float32).eval()
        synth.randomise_bn(model, seed=seed)
        inp = synth.mvsnet_inputs(batch, views, 32 * hb, 32 * wb, 8 * db, seed=seed)
        with torch.no_grad():
            got = model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])
            want = oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], model.state_dict())
        assert rel_err(got["depth"], want["depth"]) < 1e-4
        assert rel_err(got["photometric_confidence"], want["photometric_confidence"]) < 2e-3

    @settings(max_examples=10, deadline=None, derandomize=True)
    @given(nsrc=st.integers(1, 4), nscale=st.integers(1, 3), hb=st.integers(1, 2), wb=st.integers(1, 2), seed=st.integers(0, 20))
    def cvp(nsrc, nscale, hb, wb, seed):
        torch.manual_seed(seed)
        model = CVPMVSNet(SimpleNamespace(nsrc=nsrc, nscale=nscale, mode="test"), volume_dtype=torch.float32).eval()
        synth.randomise_bn(model, seed=seed)
        f = 2 ** nscale                                        # the coarsest level must still be even (network.py:44-74)
        c = synth.cvp_inputs(1, nsrc, f * 4 * hb, f * 4 * wb, seed=seed)
        keys = ("ref_img", "src_imgs", "ref_in", "src_in", "ref_ex", "src_ex", "depth_min", "depth_max")
        with torch.no_grad():
            got = model(*[c[k] for k in keys])
            want = oracle.cvp_forward(c, model.state_dict(), nscale)
        assert len(got["depth_est_list"]) == nscale
        for a, b in zip(got["depth_est_list"], want["depth_est_list"]):
            assert rel_err(a, b) < 1e-4
        assert rel_err(got["prob_confidence"], want["prob_confidence"]) < 2e-3

    mvsnet()
    cvp()
