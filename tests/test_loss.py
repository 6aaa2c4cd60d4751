"""Fused self-supervised loss (csrc/loss.cu) against the oracle's restatement of UnSupLoss.forward
(oracle/planesweep.py unsup_loss, pinned to the reference by tests/golden/*_unsup_loss.npz) on seeded inputs the golden
fixtures do not cover: more / fewer views, partial masks (a depth map that throws part of the image outside the sources),
both resize variants, each of the four outputs driven separately through the backward."""
import pytest
import torch

from conftest import rel_err


@pytest.fixture(params=["emu", pytest.param("gpu", marks=pytest.mark.gpu)])
def be(request):
    return request.getfixturevalue(request.param)


def _case(synth, batch, views, hf, wf, downscale, seed, depth_scale=1.0):
    s = 4 if downscale else 1
    li = synth.mvsnet_inputs(batch, views, hf * 4, wf * 4, 8, seed=seed)
    imgs = li["imgs"] if downscale else torch.nn.functional.avg_pool2d(li["imgs"].flatten(0, 1), 4).view(batch, views, 3, hf, wf)
    assert imgs.shape[-2:] == (hf * s, wf * s)
    depth = synth.plausible_depth(batch, hf, wf, seed=seed) * depth_scale
    return imgs.contiguous(), li["cams"], depth.contiguous()


@pytest.mark.parametrize("batch,views,hf,wf,downscale,scale", [
    (1, 4, 12, 20, True, 1.0), (2, 6, 16, 24, True, 1.0), (1, 5, 16, 20, False, 1.0),
    (2, 5, 16, 20, True, 0.35),        # a band of pixels leaves the sources: masks, clamped taps, fewer than 3 valid views
])
def test_fused_loss_matches_oracle(be, oracle, batch, views, hf, wf, downscale, scale):
    from ssmvs_b200 import ops, synth
    imgs, cams, depth = _case(synth, batch, views, hf, wf, downscale, seed=11 + views, depth_scale=scale)
    w_s = 0.18 if downscale else 0.05
    d_ref = depth.clone().requires_grad_(True)
    want = oracle.unsup_loss(imgs, cams, d_ref, downscale, w_s)
    d = be.to(depth).requires_grad_(True)
    out = ops.unsup_loss(be.to(imgs), be.to(cams), d, 1.0, w_s)
    for i, k in enumerate(("total", "reconstr", "ssim", "smooth")):
        assert rel_err(out[i], want[k]) < 1e-5, k
    # each output separately, then an arbitrary mix: the backward takes d L / d out[4]
    for coeffs in ((1, 0, 0, 0), (0, 1, 0, 0), (0, 0, 1, 0), (0, 0, 0, 1), (0.3, -2.0, 0.7, 5.0)):
        d_ref.grad = None
        d.grad = None
        sum(c * want[k] for c, k in zip(coeffs, ("total", "reconstr", "ssim", "smooth"))).backward(retain_graph=True)
        (out * torch.tensor(coeffs, dtype=torch.float32, device=out.device)).sum().backward(retain_graph=True)
        assert rel_err(d.grad, d_ref.grad) < 2e-4, coeffs


def test_fused_loss_module_attributes_and_errors(be, oracle):
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    imgs, cams, depth = _case(synth, 1, 5, 16, 20, True, seed=2)
    crit = UnSupLoss()
    total = crit(be.to(imgs), be.to(cams), be.to(depth))
    assert total is crit.unsup_loss and total.dim() == 0
    assert rel_err(12 * crit.reconstr_loss + 6 * crit.ssim_loss + 0.18 * crit.smooth_loss, total) < 1e-6
    with pytest.raises(RuntimeError):          # hazard H5: three source views at least
        crit(be.to(imgs[:, :3]), be.to(cams[:, :3]), be.to(depth))
    with pytest.raises(ValueError):            # views must be 4x the depth map (jdacs) ...
        crit(be.to(imgs[..., ::2, ::2].contiguous()), be.to(cams), be.to(depth))
    with pytest.raises(NotImplementedError):   # ... and carry no gradient
        crit(be.to(imgs).requires_grad_(True), be.to(cams), be.to(depth))


def test_loss_helper_shims_match_oracle(oracle):
    """losses/modules.py keeps the reference's helper names for the (out-of-scope) co-segmentation loss: same values as the oracle."""
    from ssmvs_b200.jdacs.losses import modules as m
    g = torch.Generator().manual_seed(0)
    x, y = torch.rand(2, 9, 11, 3, generator=g), torch.rand(2, 9, 11, 3, generator=g)
    mask = (torch.rand(2, 9, 11, 1, generator=g) > 0.3).float()
    depth = torch.rand(2, 9, 11, 1, generator=g)
    assert rel_err(m.SSIM()(x, y, mask), oracle.ssim_map(x, y, mask)) < 1e-5
    assert rel_err(m.compute_reconstr_loss(y, x, mask, simple=False), oracle.reconstr_loss(y, x, mask)) < 1e-6
    assert rel_err(m.depth_smoothness(depth, x, 1.0), oracle.depth_smoothness(depth, x, 1.0)) < 1e-6
    assert torch.equal(m.gradient_x(x), x[:, :, :-1] - x[:, :, 1:]) and torch.equal(m.gradient_y(x), x[:, :-1] - x[:, 1:])
    dx, dy = m.gradient(x)
    assert torch.equal(dx, x[:, :, 1:] - x[:, :, :-1]) and torch.equal(dy, x[:, 1:] - x[:, :-1])


def test_fused_loss_random_shapes_against_oracle(emu, oracle):
    """Property-style sweep on the host-emulation build: random batch / view counts, odd map sizes, depth scales that push parts
    of the image out of the sources -- scalars and d/d depth must match the oracle every time."""
    from hypothesis import given, settings, strategies as st
    from ssmvs_b200 import ops, synth

    @settings(max_examples=10, deadline=None, derandomize=True)
    @given(batch=st.integers(1, 2), views=st.integers(4, 7), hf=st.integers(5, 13), wf=st.integers(5, 17),
           downscale=st.booleans(), scale=st.sampled_from([1.0, 0.6, 0.35]), seed=st.integers(0, 50))
    def check(batch, views, hf, wf, downscale, scale, seed):
        li = synth.mvsnet_inputs(batch, views, hf * 4, wf * 4, 8, seed=seed)
        imgs = li["imgs"] if downscale else torch.nn.functional.avg_pool2d(li["imgs"].flatten(0, 1), 4).view(batch, views, 3, hf, wf).contiguous()
        depth = (synth.plausible_depth(batch, hf, wf, seed=seed) * scale).contiguous()
        w_s = 0.18 if downscale else 0.05
        d_ref = depth.clone().requires_grad_(True)
        want = oracle.unsup_loss(imgs, li["cams"], d_ref, downscale, w_s)
        want["total"].backward()
        d = depth.clone().requires_grad_(True)
        out = ops.unsup_loss(imgs, li["cams"], d, 1.0, w_s)
        out[0].backward()
        for i, k in enumerate(("total", "reconstr", "ssim", "smooth")):
            assert rel_err(out[i], want[k]) < 2e-5, (k, float(out[i]), float(want[k]))
        assert rel_err(d.grad, d_ref.grad) < 3e-4

    check()
