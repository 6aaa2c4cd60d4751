"""Generate tests/golden/*.npz by running the UNMODIFIED reference, imported in place.

TEST INFRASTRUCTURE.  Run only in the build container (needs /root/reference, which does not
exist on the GPU box):      python oracle/gen_golden.py

The reference holds no golden vectors for the plane-sweep path, so the
oracle and the CUDA kernels are pinned to what the reference modules themselves return on
seeded synthetic inputs.  Only harness shims are applied, none touches reference arithmetic:
  * sys.argv reset before importing jdacs/losses (config.py parses argv at import; hazard H9)
  * torch.Tensor.cuda -> identity on this CPU-only box for jdacs-ms (hard-coded .cuda(); H4)
  * the two trees both use the package name `models`/`losses`, so each runs in its own process.
While generating, every fixture is also compared with oracle/planesweep.py; a mismatch aborts.
"""
from __future__ import annotations

import argparse
import importlib.util
import os
import subprocess
import sys
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
OUT = os.path.join(ROOT, "tests", "golden")


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synth = _load("synth", os.path.join(ROOT, "self-supervised-mvs_b200", "synth.py"))
oracle = _load("planesweep_oracle", os.path.join(ROOT, "oracle", "planesweep.py"))


def _np(d):
    out = {}
    for k, v in d.items():
        if isinstance(v, torch.Tensor):
            out[k] = v.detach().cpu().numpy()
        elif isinstance(v, (list, tuple)):
            for i, t in enumerate(v):
                out["%s.%d" % (k, i)] = t.detach().cpu().numpy()
        else:
            out[k] = np.asarray(v)
    return out


def _save(name, **arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **_np(arrays))
    print("wrote %s (%.1f KB)" % (path, os.path.getsize(path) / 1024))


def _close(a, b, tol, what):
    a, b = a.detach().float(), b.detach().float()
    err = (a - b).abs().max().item()
    ref = b.abs().max().item() + 1e-12
    print("  oracle vs reference %-28s max|d|=%.3e (rel %.2e)" % (what, err, err / ref))
    assert err <= tol * max(ref, 1.0), "oracle disagrees with the reference on " + what


def _small_cams(batch, views, hf, wf, ndepth, seed=0):
    base = synth.mvsnet_inputs(batch, views, hf * 4, wf * 4, ndepth, seed)
    return base


def _state(model, prefix=""):
    return {prefix + k: v.detach().clone() for k, v in model.state_dict().items() if "num_batches_tracked" not in k}


# ---------------------------------------------------------------------------------------------
def gen_jdacs():
    sys.argv = ["gen_golden"]  # H9
    sys.path.insert(0, os.path.join(REF, "jdacs"))
    sys.dont_write_bytecode = True
    from models import module as rmod
    from models import mvsnet as rnet
    torch.set_num_threads(4)

    # --- a1: homo_warping -----------------------------------------------------------------
    g = torch.Generator().manual_seed(1)
    cam = _small_cams(2, 3, 12, 16, 6)
    fea = torch.randn(2, 8, 12, 16, generator=g)
    P = cam["proj_matrices"]
    dv = cam["depth_values"]
    ref_out = rmod.homo_warping(fea, P[:, 1], P[:, 0], dv)
    _close(oracle.homo_warping(fea, P[:, 1], P[:, 0], dv), ref_out, 1e-5, "homo_warping")
    _close(oracle.homo_warping(fea, P[:, 1], P[:, 0], dv, restated_sampler=True), ref_out, 1e-5, "homo_warping(restated)")
    # a wide-baseline pair that leaves the frustum on most planes (zero padding, behind-camera)
    far = dv.clone() * 0.02
    ref_far = rmod.homo_warping(fea, P[:, 2], P[:, 0], far)
    _close(oracle.homo_warping(fea, P[:, 2], P[:, 0], far, restated_sampler=True), ref_far, 1e-5, "homo_warping(out of frustum)")
    _save("jdacs_warp", src_fea=fea, src_proj=P[:, 1], src_proj2=P[:, 2], ref_proj=P[:, 0], depth_values=dv,
          depth_far=far, warped=ref_out, warped_far=ref_far)

    # --- a4..a8: whole MVSNet forward, eval ---------------------------------------------------
    torch.manual_seed(0)
    model = rnet.MVSNet(refine=False)
    with torch.no_grad():
        model.cost_regularization.prob.weight.mul_(64.0)  # peaky softmax (H11)
        # non-trivial running statistics so that eval-mode BN folding is exercised
        gg = torch.Generator().manual_seed(5)
        for m in model.modules():
            if isinstance(m, (torch.nn.BatchNorm3d, torch.nn.BatchNorm2d)):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=gg))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=gg))
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.weight.shape, generator=gg))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=gg))
    inp = synth.mvsnet_inputs(1, 3, 64, 96, 8, seed=2)
    inp["depth_values"] = (425.0 + 40.0 * torch.arange(8, dtype=torch.float32)).unsqueeze(0)
    cap = {}
    hook = model.cost_regularization.register_forward_hook(
        lambda m, i, o: cap.update(variance=i[0].detach().clone(), cost_reg=o.detach().clone()))
    model.eval()
    with torch.no_grad():
        feats = torch.stack([model.feature(inp["imgs"][:, v]) for v in range(3)])
        out = model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])
    sd = _state(model)
    st = {}
    with torch.no_grad():
        mine = oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], sd, False, False, st)
    _close(st["features"], feats, 1e-5, "FeatureNet")
    _close(st["variance"], cap["variance"], 1e-5, "variance volume")
    _close(st["cost_reg"], cap["cost_reg"].squeeze(1), 2e-4, "CostRegNet")
    _close(mine["depth"], out["depth"], 1e-5, "depth")
    _close(mine["photometric_confidence"], out["photometric_confidence"], 1e-4, "confidence")
    prob = torch.softmax(cap["cost_reg"].squeeze(1), 1)
    index = torch.sum(prob * torch.arange(8, dtype=torch.float32).reshape(1, 8, 1, 1), 1).long()
    assert torch.equal(index, st["index"]), "depth index mismatch"
    print("  depth range %.2f..%.2f, peak prob mean %.3f" % (out["depth"].min(), out["depth"].max(),
                                                              prob.max(1)[0].mean()))

    # --- training-mode forward + backward (batch-stat BN, autograd through warp / reg / softargmin) ----
    model.train()
    hook.remove()
    gw = torch.Generator().manual_seed(9)
    wmap = torch.randn(1, 16, 24, generator=gw)
    model.zero_grad()
    out_t = model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])
    (out_t["depth"] * wmap).sum().backward()
    grads = {"grad." + k: v.grad.detach().clone() for k, v in model.named_parameters()
             if k in ("cost_regularization.conv0.conv.weight", "cost_regularization.prob.weight",
                      "cost_regularization.conv6.bn.weight", "cost_regularization.conv7.0.weight",
                      "cost_regularization.conv1.conv.weight", "feature.feature.weight", "feature.conv0.conv.weight")}
    _save("jdacs_mvsnet", imgs=inp["imgs"], proj_matrices=inp["proj_matrices"], depth_values=inp["depth_values"],
          features=feats, variance=cap["variance"], cost_reg=cap["cost_reg"].squeeze(1), depth=out["depth"],
          photometric_confidence=out["photometric_confidence"], index=index, train_depth=out_t["depth"],
          loss_weight=wmap, **{"sd." + k: v for k, v in sd.items()}, **grads)

    # --- a10 / a11: loss warp and UnSupLoss -----------------------------------------------------
    from losses import homography as rhom
    from losses import unsup_loss as rloss
    li = synth.mvsnet_inputs(2, 5, 64, 80, 8, seed=3)
    depth = synth.plausible_depth(2, 16, 20, seed=3).requires_grad_(True)
    g2 = torch.Generator().manual_seed(4)
    img = torch.randn(2, 16, 20, 3, generator=g2)
    wimg = torch.randn(2, 16, 20, 3, generator=g2)
    warped, mask = rhom.inverse_warping(img, li["cams"][:, 0], li["cams"][:, 2], depth)
    (warped * wimg).sum().backward()
    gdepth = depth.grad.detach().clone()
    d2 = depth.detach().clone().requires_grad_(True)
    ow, om = oracle.inverse_warping(img, li["cams"][:, 0], li["cams"][:, 2], d2)
    (ow * wimg).sum().backward()
    _close(ow, warped, 1e-5, "inverse_warping")
    assert torch.equal(om, mask), "inverse_warping mask mismatch"
    _close(d2.grad, gdepth, 1e-4, "inverse_warping d/ddepth")
    # a depth map that throws a band of pixels outside the source image (mask = 0 region, clamped taps)
    dfar = (depth.detach() * 0.35).contiguous()
    warped_far, mask_far = rhom.inverse_warping(img, li["cams"][:, 0], li["cams"][:, 1], dfar)
    ow2, om2 = oracle.inverse_warping(img, li["cams"][:, 0], li["cams"][:, 1], dfar)
    _close(ow2, warped_far, 1e-5, "inverse_warping(far)")
    assert torch.equal(om2, mask_far)
    print("  mask coverage: %.2f / far %.2f" % (mask.mean(), mask_far.mean()))
    _save("jdacs_invwarp", img=img, cams=li["cams"], depth=depth, weight=wimg, warped=warped, mask=mask,
          grad_depth=gdepth, depth_far=dfar, warped_far=warped_far, mask_far=mask_far)

    crit = rloss.UnSupLoss()
    d3 = depth.detach().clone().requires_grad_(True)
    total = crit(li["imgs"], li["cams"], d3)
    total.backward()
    d4 = depth.detach().clone().requires_grad_(True)
    mine = oracle.unsup_loss(li["imgs"], li["cams"], d4, True, 0.18)
    mine["total"].backward()
    _close(mine["total"], total, 1e-5, "UnSupLoss total")
    _close(mine["reconstr"], crit.reconstr_loss, 1e-5, "UnSupLoss reconstr")
    _close(mine["ssim"], crit.ssim_loss, 1e-5, "UnSupLoss ssim")
    _close(mine["smooth"], crit.smooth_loss, 1e-5, "UnSupLoss smooth")
    _close(d4.grad, d3.grad, 1e-4, "UnSupLoss d/ddepth")
    _save("jdacs_unsup_loss", imgs=li["imgs"], cams=li["cams"], depth=depth, total=total,
          reconstr=crit.reconstr_loss, ssim=crit.ssim_loss, smooth=crit.smooth_loss, grad_depth=d3.grad)


# ---------------------------------------------------------------------------------------------
def gen_jdacs_ms():
    sys.argv = ["gen_golden"]
    sys.path.insert(0, os.path.join(REF, "jdacs-ms"))
    sys.dont_write_bytecode = True
    torch.Tensor.cuda = lambda self, *a, **k: self  # H4: CPU-only box
    from models import modules as rmod
    from models import network as rnet
    torch.set_num_threads(4)

    ci = synth.cvp_inputs(1, 2, 32, 48, seed=6)
    # --- a2: homo_warping from (K, E) -------------------------------------------------------
    g = torch.Generator().manual_seed(7)
    fea = torch.randn(1, 16, 16, 24, generator=g)
    kr = oracle.condition_intrinsics(ci["ref_in"], 1)
    ks = oracle.condition_intrinsics(ci["src_in"], 1)
    hyp = oracle.sweeping_depth_hypos(ci["depth_min"], ci["depth_max"], 1)
    ref_hyp = rmod.calSweepingDepthHypo(kr, ks[:, 0], ci["ref_ex"], ci["src_ex"], ci["depth_min"], ci["depth_max"])
    assert ref_hyp.shape == hyp.shape, "reference torch.range gave %s planes" % (ref_hyp.shape,)
    _close(hyp, ref_hyp, 1e-6, "calSweepingDepthHypo")
    ref_w = rmod.homo_warping(fea, kr, ks[:, 0], ci["ref_ex"], ci["src_ex"][:, 0], ref_hyp)
    _close(oracle.homo_warping_ms(fea, kr, ks[:, 0], ci["ref_ex"], ci["src_ex"][:, 0], ref_hyp), ref_w, 1e-5, "homo_warping(K,E)")

    # --- a9: calDepthHypo ---------------------------------------------------------------------
    depth_up = synth.plausible_depth(1, 32, 48, seed=8)
    ref_h = rmod.calDepthHypo(None, depth_up, ci["ref_in"], ci["src_in"], ci["ref_ex"], ci["src_ex"],
                              ci["depth_min"], ci["depth_max"], 0)
    my_h = oracle.depth_hypos_refine(depth_up, ci["ref_in"], ci["src_in"][:, 0], ci["ref_ex"], ci["src_ex"][:, 0])
    _close(my_h, ref_h, 1e-6, "calDepthHypo")
    print("  refine interval = %.4f" % (ref_h[0, 5, 0, 0] - ref_h[0, 4, 0, 0]))

    # --- a3: proj_cost (per-pixel hypotheses, H2 variance) --------------------------------------
    ref_f = torch.randn(1, 16, 32, 48, generator=g)
    src_f = [[torch.randn(1, 16, 32, 48, generator=g)] for _ in range(2)]
    st = SimpleNamespace(nsrc=2, mode="train")
    ref_c = rmod.proj_cost(st, ref_f.clone(), src_f, 0, ci["ref_in"], ci["src_in"], ci["ref_ex"], ci["src_ex"], ref_h)
    rp = oracle.compose_projection(ci["ref_in"], ci["ref_ex"])
    sp = [oracle.compose_projection(ci["src_in"][:, i], ci["src_ex"][:, i]) for i in range(2)]
    my_c = oracle.variance_volume(ref_f, [s[0] for s in src_f], rp, sp, ref_h, True)
    _close(my_c, ref_c, 1e-5, "proj_cost")
    _save("ms_warp", src_fea=fea, ref_in_l1=kr, src_in_l1=ks, ref_in=ci["ref_in"], src_in=ci["src_in"],
          ref_ex=ci["ref_ex"], src_ex=ci["src_ex"], depth_min=ci["depth_min"], depth_max=ci["depth_max"],
          sweep_hypos=ref_hyp, warped=ref_w, depth_up=depth_up, refine_hypos=ref_h, ref_fea=ref_f,
          src_fea0=src_f[0][0], src_fea1=src_f[1][0], proj_cost=ref_c)

    # --- whole CVP forward, eval ----------------------------------------------------------------
    torch.manual_seed(0)
    args = SimpleNamespace(nsrc=2, nscale=2, mode="test")
    model = rnet.CVPMVSNet(args)
    with torch.no_grad():
        model.cost_reg_refine.prob0.weight.mul_(64.0)
        gg = torch.Generator().manual_seed(5)
        for m in model.modules():
            if isinstance(m, torch.nn.BatchNorm3d):
                m.running_mean.copy_(0.1 * torch.randn(m.running_mean.shape, generator=gg))
                m.running_var.copy_(0.5 + torch.rand(m.running_var.shape, generator=gg))
                m.weight.copy_(0.75 + 0.5 * torch.rand(m.weight.shape, generator=gg))
                m.bias.copy_(0.1 * torch.randn(m.bias.shape, generator=gg))
    model.eval()
    with torch.no_grad():
        out = model(ci["ref_img"], ci["src_imgs"], ci["ref_in"], ci["src_in"], ci["ref_ex"], ci["src_ex"],
                    ci["depth_min"], ci["depth_max"])
        sd = _state(model)
        mine = oracle.cvp_forward(ci, sd, 2)
    for i, (a, b) in enumerate(zip(mine["depth_est_list"], out["depth_est_list"])):
        _close(a, b, 2e-5, "CVP depth level %d" % i)
    _close(mine["prob_confidence"], out["prob_confidence"], 1e-4, "CVP confidence")
    _save("ms_cvp", **{k: v for k, v in ci.items()}, depth_est_list=out["depth_est_list"],
          prob_confidence=out["prob_confidence"], **{"sd." + k: v for k, v in sd.items()})

    # --- jdacs-ms UnSupLoss (no x0.25 resize, 0.05 smoothness weight) ------------------------------
    from losses import unsup_loss as rloss
    li = synth.mvsnet_inputs(1, 4, 64, 80, 8, seed=11)
    imgs = torch.nn.functional.interpolate(li["imgs"].flatten(0, 1), scale_factor=0.25, mode="bilinear").reshape(1, 4, 3, 16, 20)
    depth = synth.plausible_depth(1, 16, 20, seed=11).requires_grad_(True)
    crit = rloss.UnSupLoss()
    total = crit(imgs, li["cams"], depth)
    total.backward()
    d2 = depth.detach().clone().requires_grad_(True)
    mine = oracle.unsup_loss(imgs, li["cams"], d2, False, 0.05)
    mine["total"].backward()
    _close(mine["total"], total, 1e-5, "ms UnSupLoss total")
    _close(d2.grad, depth.grad, 1e-4, "ms UnSupLoss d/ddepth")
    _save("ms_unsup_loss", imgs=imgs, cams=li["cams"], depth=depth, total=total, reconstr=crit.reconstr_loss,
          ssim=crit.ssim_loss, smooth=crit.smooth_loss, grad_depth=depth.grad)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tree", choices=["jdacs", "jdacs-ms"], default=None)
    a = ap.parse_args()
    if a.tree is None:
        for t in ("jdacs", "jdacs-ms"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "--tree", t])
    elif a.tree == "jdacs":
        gen_jdacs()
    else:
        gen_jdacs_ms()
