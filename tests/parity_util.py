"""End-to-end parity harness: the product path at its benchmarked precision against the fp32 oracle, on inputs
for which the comparison means something.

Hazard H11: with default-initialised weights the softmax over depth is uniform and the
regressed depth is mean(depth_values) whatever the cost volume holds, so an end-to-end comparison proves nothing.
`peaky_*` therefore (i) randomises the BatchNorm statistics (the recipe of oracle/gen_golden.py), (ii) runs the
ORACLE in fp32 on the device (TF32 off) and scales the last convolution by the smallest power of two for which the
mean peak probability reaches `target_peak`, and (iii) returns the conditions it reached so that the caller can
assert them.  Test infrastructure: imported by tests/, __graft_entry__.smoke() and bench.py's parity check only.
"""
from __future__ import annotations

import importlib.util
import os
import sys
from types import SimpleNamespace
from typing import Callable, Dict, Optional

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def load_oracle():
    name = "planesweep_oracle"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", "planesweep.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


class fp32_exact:
    """Run the oracle on `dev` (its helper tensors are created on torch's default device) with library convolutions / matmuls in
    true fp32 (torch's default lets cuDNN use TF32; tf32=True keeps that default)."""

    def __init__(self, dev, tf32: bool = False):
        self.tf32, self.dev = tf32, torch.device(dev)

    def __enter__(self):
        self.prev = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
        torch.backends.cudnn.allow_tf32 = self.tf32
        torch.backends.cuda.matmul.allow_tf32 = self.tf32
        self.dev.__enter__()

    def __exit__(self, *a):
        self.dev.__exit__(*a)
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = self.prev


def _peak(reg: torch.Tensor) -> float:
    return F.softmax(reg, 1).max(1)[0].mean().item()


def _pow2_scale(reg0: torch.Tensor, target_peak: float) -> float:
    """smallest power of two s with mean peak probability of softmax(s * reg0) >= target (s * w is exact in fp32)."""
    s = 1.0
    while _peak(reg0 * s) < target_peak and s < 2.0 ** 20:
        s *= 2.0
    return s


def depth_parity(got: Dict[str, torch.Tensor], want: Dict[str, torch.Tensor]) -> Dict[str, float]:
    """north_star: depth within 1e-3 relative, expected-plane index exact (hazard H12: the index is a truncated
    float sum, so a mismatch is 'explained' when the oracle's own sum sits within 1e-4 of an integer)."""
    gd, wd = got["depth"].float().flatten(), want["depth"].float().flatten()
    rel = (gd - wd).abs() / wd.abs()
    n = rel.numel()
    out = {"pixels": n, "depth_rel_max": rel.max().item(), "depth_rel_p999": rel.kthvalue(max(1, int(0.999 * n)))[0].item(),
           "depth_rel_median": rel.median().item(), "depth_rel_mean": rel.mean().item(),
           "frac_within_1e-3": (rel <= 1e-3).float().mean().item()}
    if "index" in got and "index" in want:
        bad = (got["index"].flatten() != want["index"].flatten())
        out["index_mismatch"] = int(bad.sum().item())
        if "index_float" in want:
            f = want["index_float"].flatten()
            out["index_mismatch_near_integer"] = int((bad & ((f - f.round()).abs() < 1e-4)).sum().item())
        out["index_off_by_more_than_1"] = int(((got["index"].flatten() - want["index"].flatten()).abs() > 1).sum().item())
    if "conf" in got and "conf" in want:
        e = (got["conf"].float() - want["conf"].float()).abs().flatten()
        out["conf_abs_max"] = e.max().item()
        out["conf_abs_p999"] = e.kthvalue(max(1, int(0.999 * e.numel())))[0].item()
        out["conf_abs_mean"] = e.mean().item()
    return out


# ------------------------------------------------------------------------------------------------ MVSNet (config 2)
def peaky_mvsnet(dev, views: int, height: int, width: int, ndepth: int, seed: int = 0, target_peak: float = 0.3,
                 batch: int = 1):
    """-> (model on dev in eval mode, inputs on dev, oracle result, conditions).  Conditions carry the peak probability
    and depth spread reached; callers assert `peak >= target_peak` and `depth_std >= 10`."""
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    oracle = load_oracle()
    torch.manual_seed(0)
    model = MVSNet(refine=False)
    synth.randomise_bn(model, 5)
    model = model.to(dev).eval()
    model.keep_index = True
    inp = {k: v.to(dev) for k, v in synth.mvsnet_inputs(batch, views, height, width, ndepth, seed=seed).items()}
    with torch.no_grad(), fp32_exact(dev):
        st = {}
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], sd, False, False, st)
        bias = sd["cost_regularization.prob.bias"].view(1, 1, 1, 1)
        scale = _pow2_scale(st["cost_reg"] - bias, target_peak)
        del st
        model.cost_regularization.prob.weight.mul_(scale)
        want, cond = oracle_mvsnet(model, inp)
    cond["prob_scale"] = scale
    return model, inp, want, cond


def oracle_mvsnet(model, inp, tf32: bool = False):
    """The oracle's fp32 forward on the inputs' device -> ({depth, conf, index, index_float}, conditions)."""
    oracle = load_oracle()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    st = {}
    with torch.no_grad(), fp32_exact(inp["imgs"].device, tf32):
        out = oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], sd, False, False, st)
        nd = st["prob"].shape[1]
        ramp = torch.arange(nd, dtype=torch.float32, device=st["prob"].device).view(1, nd, 1, 1)
        want = {"depth": out["depth"], "conf": out["photometric_confidence"], "index": st["index"],
                "index_float": torch.sum(st["prob"] * ramp, 1)}
        cond = {"peak": st["prob"].max(1)[0].mean().item(), "depth_std": out["depth"].std().item(),
                "depth_min": out["depth"].min().item(), "depth_max": out["depth"].max().item()}
    return want, cond


def product_mvsnet(model, inp, dtype) -> Dict[str, torch.Tensor]:
    model.volume_dtype = dtype
    with torch.no_grad():
        o = model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])
    return {"depth": o["depth"], "conf": o["photometric_confidence"], "index": o["depth_index"]}


def ideal_storage_mvsnet(model, inp, dtype) -> Dict[str, torch.Tensor]:
    """What 16-bit STORAGE alone costs: the oracle's arithmetic in fp32 with every tensor the product keeps in `dtype`
    (weights, FeatureNet activations, features, variance volume, U-Net activations) rounded once to `dtype`.  No kernel can
    be closer to the fp32 oracle than this while storing in `dtype`; the product path is compared with it as well."""
    oracle = load_oracle()
    q: Callable[[torch.Tensor], torch.Tensor] = (lambda x: x.to(dtype).float())
    P = {k: v.detach().clone() for k, v in model.state_dict().items()}
    fp, R = oracle._sub(P, "feature."), oracle._sub(P, "cost_regularization.")

    def affine(Pm, pre):
        s = Pm[pre + "weight"] * torch.rsqrt(Pm[pre + "running_var"] + 1e-5)
        return s, Pm[pre + "bias"] - Pm[pre + "running_mean"] * s

    def cbr2(x, name, s, p):
        sc, sh = affine(fp, name + ".bn.")
        return q(F.relu(F.conv2d(x, q(fp[name + ".conv.weight"]), None, s, p) * sc.view(1, -1, 1, 1) + sh.view(1, -1, 1, 1)))

    def c3(x, name, s):
        sc, sh = affine(R, name + ".bn.")
        return F.relu(F.conv3d(x, q(R[name + ".conv.weight"]), None, s, 1) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))

    def d3(x, name):
        sc, sh = affine(R, name + ".1.")
        return F.relu(F.conv_transpose3d(x, q(R[name + ".0.weight"]), None, 2, 1, 1) * sc.view(1, -1, 1, 1, 1) + sh.view(1, -1, 1, 1, 1))

    with torch.no_grad(), fp32_exact(inp["imgs"].device):
        imgs, proj, dv = inp["imgs"], inp["proj_matrices"], inp["depth_values"]
        feats = []
        for v in range(imgs.shape[1]):
            x = q(imgs[:, v])
            x = cbr2(cbr2(x, "conv0", 1, 1), "conv1", 1, 1)
            x = cbr2(cbr2(cbr2(x, "conv2", 2, 2), "conv3", 1, 1), "conv4", 1, 1)
            x = cbr2(cbr2(x, "conv5", 2, 2), "conv6", 1, 1)
            feats.append(q(F.conv2d(x, q(fp["feature.weight"]), fp["feature.bias"], 1, 1)))
        var = q(oracle.variance_volume(feats[0], feats[1:], proj[:, 0], [proj[:, v] for v in range(1, imgs.shape[1])], dv))
        c0 = q(c3(var, "conv0", 1)); c2 = q(c3(q(c3(c0, "conv1", 2)), "conv2", 1)); c4 = q(c3(q(c3(c2, "conv3", 2)), "conv4", 1))
        c6 = q(c3(q(c3(c4, "conv5", 2)), "conv6", 1))
        u = q(c4 + d3(c6, "conv7")); u = q(c2 + d3(u, "conv9")); u = q(c0 + d3(u, "conv11"))
        reg = F.conv3d(u, q(R["prob.weight"]), R["prob.bias"], 1, 1).squeeze(1)
        prob, depth = oracle.soft_argmin(reg, dv)
        index, conf = oracle.photometric_confidence(prob)
    return {"depth": depth, "conf": conf, "index": index}


# ------------------------------------------------------------------------------------------------ CVP-MVSNet (config 3)
_CVP_KEYS = ("ref_img", "src_imgs", "ref_in", "src_in", "ref_ex", "src_ex", "depth_min", "depth_max")


def peaky_cvp(dev, nsrc: int, nscale: int, height: int, width: int, seed: int = 0, target_peak: float = 0.3):
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet
    oracle = load_oracle()
    torch.manual_seed(0)
    model = CVPMVSNet(SimpleNamespace(nsrc=nsrc, nscale=nscale, mode="test"))
    synth.randomise_bn(model, 5)
    model = model.to(dev).eval()
    model.keep_index = True
    inp = {k: v.to(dev) for k, v in synth.cvp_inputs(1, nsrc, height, width, seed=seed).items()}
    with torch.no_grad(), fp32_exact(dev):
        sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
        st = {}
        oracle.cvp_forward(inp, sd, nscale, False, False, st)
        scale = _pow2_scale(st["cost_reg0"] - sd["cost_reg_refine.prob0.bias"].view(1, 1, 1, 1), target_peak)
        del st
        model.cost_reg_refine.prob0.weight.mul_(scale)
    want, cond = oracle_cvp(model, inp, nscale)
    cond["prob_scale"] = scale
    return model, inp, want, cond


def oracle_cvp(model, inp, nscale: int, tf32: bool = False):
    oracle = load_oracle()
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    st = {}
    with torch.no_grad(), fp32_exact(inp["ref_img"].device, tf32):
        out = oracle.cvp_forward(inp, sd, nscale, False, False, st)
        coarse = out["depth_est_list"][-1]
        cond = {"peak": F.softmax(st["cost_reg0"], 1).max(1)[0].mean().item(), "depth_std": coarse.std().item(),
                "depth_min": coarse.min().item(), "depth_max": coarse.max().item()}
    return {"depth_est_list": out["depth_est_list"], "conf": out["prob_confidence"]}, cond


def product_cvp(model, inp, dtype):
    model.volume_dtype = dtype
    with torch.no_grad():
        o = model(*[inp[k] for k in _CVP_KEYS])
    return {"depth_est_list": o["depth_est_list"], "conf": o["prob_confidence"], "index": o.get("depth_index")}


def cvp_parity(got, want) -> Dict[str, Dict[str, float]]:
    """per pyramid level (finest first); the finer levels' hypotheses are built from the coarser level's own depth, so
    their errors include the propagated hypothesis differences, exactly as a user would see them."""
    out = {}
    for lvl, (a, b) in enumerate(zip(got["depth_est_list"], want["depth_est_list"])):
        out["level%d" % lvl] = depth_parity({"depth": a}, {"depth": b})
    e = (got["conf"].float() - want["conf"].float()).abs().flatten()
    out["conf_abs_max"] = e.max().item()
    out["conf_abs_p999"] = e.kthvalue(max(1, int(0.999 * e.numel())))[0].item()
    out["conf_abs_mean"] = e.mean().item()
    return out
