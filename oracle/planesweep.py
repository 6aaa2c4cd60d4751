"""CPU ORACLE for the MVSNet / CVP-MVSNet plane-sweep hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, with plain torch-CPU tensor arithmetic, what the reference
(ToughStoneX/Self-Supervised-MVS) computes on the path named by BASELINE.json `north_star`.
It is the checker for the CUDA kernels and the timed "port" for bench.py's cpu_baseline /
`--impl reference` arm.  Nothing under `self-supervised-mvs_b200/` imports it; only `tests/`,
`__graft_entry__.smoke()` and `bench.py` (cpu_baseline / reference arm) may.

Parity pinning: the reference holds NO golden vectors for this path, so
the oracle is pinned against outputs of the reference itself, imported in the build container
by `oracle/gen_golden.py` (fixtures in `tests/golden/`, checked by `tests/test_oracle_golden.py`).

Third-party arithmetic: the reference calls PyTorch/ATen (`F.grid_sample`, `Conv3d`,
`ConvTranspose3d`, `BatchNorm3d`, `softmax`, `avg_pool3d`, `torch.inverse`); there is no pin file,
the README names torch 1.1.0, this image has torch 2.11.0.  The oracle follows ATen 2.11
semantics as exercised by the reference call sites (hazard H1: grid_sample defaults to
align_corners=False there) and restates the bilinear sampler itself (`bilinear_zeros`) so
that the gather arithmetic is written out rather than inherited.

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# a1 / a2 / a3: homography warp
# --------------------------------------------------------------------------------------------
def relative_projection(src_proj: Tensor, ref_proj: Tensor) -> Tuple[Tensor, Tensor]:
    """rot [B,3,3], trans [B,3,1] of src_proj @ inv(ref_proj).  jdacs/models/module.py:116-118."""
    rel = src_proj @ torch.inverse(ref_proj)
    return rel[:, :3, :3], rel[:, :3, 3:4]


def compose_projection(intr: Tensor, extr: Tensor) -> Tensor:
    """[[K @ E[:3]], [0,0,0,1]] as the CVP variant builds it.  jdacs-ms/models/modules.py:71-75."""
    top = intr @ extr[:, 0:3, :]
    last = torch.tensor([[[0.0, 0.0, 0.0, 1.0]]], dtype=top.dtype).repeat(top.shape[0], 1, 1)
    return torch.cat((top, last), 1)


def sweep_coords(rot: Tensor, trans: Tensor, depth: Tensor, height: int, width: int) -> Tuple[Tensor, Tensor]:
    """Source-view pixel coordinates (u, v), each [B,D,H*W], of every reference pixel on every plane.

    depth is [B,D] (fronto-parallel sweep, jdacs/models/module.py:127-131) or [B,D,H,W]
    (per-pixel hypotheses, jdacs-ms/models/modules.py:239-242)."""
    b = rot.shape[0]
    ys, xs = torch.meshgrid(torch.arange(height, dtype=torch.float32), torch.arange(width, dtype=torch.float32),
                            indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(height * width)))  # [3,HW]
    ray = rot @ pix.unsqueeze(0).expand(b, -1, -1)                                     # [B,3,HW]
    nd = depth.shape[1]
    dd = depth.reshape(b, 1, nd, 1) if depth.dim() == 2 else depth.reshape(b, 1, nd, height * width)
    pts = ray.unsqueeze(2) * dd + trans.reshape(b, 3, 1, 1)                            # [B,3,D,HW]
    uv = pts[:, :2] / pts[:, 2:3]
    return uv[:, 0], uv[:, 1]


def normalise_coords(u: Tensor, v: Tensor, height: int, width: int) -> Tuple[Tensor, Tensor]:
    """x/((W-1)/2)-1, y/((H-1)/2)-1.  jdacs/models/module.py:132-133."""
    return u / ((width - 1) / 2) - 1, v / ((height - 1) / 2) - 1


def bilinear_zeros(fea: Tensor, xn: Tensor, yn: Tensor, align_corners: bool = False) -> Tensor:
    """Restatement of ATen grid_sampler_2d(bilinear, zeros) on normalised coords.

    fea [B,C,H,W]; xn, yn [B,P] -> [B,C,P].  Un-normalisation follows ATen
    (GridSampler.h grid_sampler_unnormalize): align_corners=False -> ((x+1)*W-1)/2,
    True -> (x+1)/2*(W-1).  Each of the four taps contributes only when it lies inside the map."""
    b, c, h, w = fea.shape
    if align_corners:
        ix, iy = (xn + 1) / 2 * (w - 1), (yn + 1) / 2 * (h - 1)
    else:
        ix, iy = ((xn + 1) * w - 1) / 2, ((yn + 1) * h - 1) / 2
    x0, y0 = torch.floor(ix), torch.floor(iy)
    x1, y1 = x0 + 1, y0 + 1
    flat = fea.reshape(b, c, h * w)
    out = torch.zeros(b, c, ix.shape[1], dtype=fea.dtype)
    for xt, yt, wt in ((x0, y0, (x1 - ix) * (y1 - iy)), (x1, y0, (ix - x0) * (y1 - iy)),
                       (x0, y1, (x1 - ix) * (iy - y0)), (x1, y1, (ix - x0) * (iy - y0))):
        ok = (xt >= 0) & (xt <= w - 1) & (yt >= 0) & (yt <= h - 1)  # NaN/inf compare false -> tap dropped
        idx = (torch.where(ok, yt, torch.zeros_like(yt)) * w + torch.where(ok, xt, torch.zeros_like(xt))).long()
        tap = torch.gather(flat, 2, idx.unsqueeze(1).expand(-1, c, -1))
        out = out + tap * torch.where(ok, wt, torch.zeros_like(wt)).unsqueeze(1)
    return out


def homo_warping(src_fea: Tensor, src_proj: Tensor, ref_proj: Tensor, depth: Tensor,
                 align_corners: bool = False, restated_sampler: bool = False) -> Tensor:
    """[B,C,D,H,W] warped source features.  jdacs/models/module.py:105-140 (a1);
    with depth [B,D,H,W] it is the warp inside proj_cost, jdacs-ms/models/modules.py:220-249 (a3).

    restated_sampler=False uses F.grid_sample (what the reference executes, and what the CPU
    baseline times); True uses `bilinear_zeros` above.  tests check the two agree."""
    b, c, h, w = src_fea.shape
    nd = depth.shape[1]
    with torch.no_grad():
        rot, trans = relative_projection(src_proj, ref_proj)
        u, v = sweep_coords(rot, trans, depth, h, w)
        xn, yn = normalise_coords(u, v, h, w)
    if restated_sampler:
        out = bilinear_zeros(src_fea, xn.reshape(b, -1), yn.reshape(b, -1), align_corners)
    else:
        grid = torch.stack((xn, yn), dim=3).reshape(b, nd * h, w, 2)
        out = F.grid_sample(src_fea, grid, mode="bilinear", padding_mode="zeros", align_corners=align_corners)
    return out.reshape(b, c, nd, h, w)


def homo_warping_ms(src_fea: Tensor, ref_in: Tensor, src_in: Tensor, ref_ex: Tensor, src_ex: Tensor,
                    depth: Tensor, align_corners: bool = False) -> Tensor:
    """CVP variant taking (K, E) pairs.  jdacs-ms/models/modules.py:62-104 (a2)."""
    return homo_warping(src_fea, compose_projection(src_in, src_ex), compose_projection(ref_in, ref_ex), depth,
                        align_corners)


# --------------------------------------------------------------------------------------------
# a4: variance cost volume
# --------------------------------------------------------------------------------------------
def variance_volume(ref_fea: Tensor, src_feas: Sequence[Tensor], ref_proj: Tensor, src_projs: Sequence[Tensor],
                    depth: Tensor, ref_sq_in_sum: bool = False, align_corners: bool = False, inplace: bool = False) -> Tensor:
    """var = S2/N - (S1/N)^2 over the reference volume and the warped sources.

    jdacs/models/mvsnet.py:120-136.  ref_sq_in_sum=True reproduces the CVP aliasing
    (hazard H2, jdacs-ms/models/network.py:114-116, modules.py:216-217): `pow_` squares the
    tensor that `volume_sum` aliases, so S1 starts from ref^2 instead of ref.
    inplace=True is the reference's eval-mode branch (mvsnet.py:130-136): the same values, accumulated in place (no autograd)."""
    nd = depth.shape[1]
    n = len(src_feas) + 1
    ref_vol = ref_fea.unsqueeze(2).repeat(1, 1, nd, 1, 1)
    s2 = ref_vol ** 2
    s1 = s2.clone() if ref_sq_in_sum else ref_vol
    for fea, proj in zip(src_feas, src_projs):
        wv = homo_warping(fea, proj, ref_proj, depth, align_corners)
        if inplace:
            s1 += wv
            s2 += wv.pow_(2)
        else:
            s1 = s1 + wv
            s2 = s2 + wv ** 2
    if inplace:
        return s2.div_(n).sub_(s1.div_(n).pow_(2))
    return s2 / n - (s1 / n) ** 2


# --------------------------------------------------------------------------------------------
# a5 / a6: 3-D U-Net regularisation (functional; P is a reference state_dict slice)
# --------------------------------------------------------------------------------------------
def _bn(x: Tensor, P: Dict[str, Tensor], pre: str, training: bool) -> Tensor:
    return F.batch_norm(x, None if training else P[pre + "running_mean"], None if training else P[pre + "running_var"],
                        P[pre + "weight"], P[pre + "bias"], training, 0.1, 1e-5)


def _cbr3(x: Tensor, P: Dict[str, Tensor], name: str, stride: int, training: bool) -> Tensor:
    """ConvBnReLU3D, jdacs/models/module.py:35-42."""
    y = F.conv3d(x, P[name + ".conv.weight"], None, stride, 1)
    return F.relu(_bn(y, P, name + ".bn.", training))


def _dbr3(x: Tensor, P: Dict[str, Tensor], name: str, stride: int, outpad: int, training: bool) -> Tensor:
    """ConvTranspose3d + BN + ReLU Sequential, jdacs/models/mvsnet.py:48-61."""
    y = F.conv_transpose3d(x, P[name + ".0.weight"], None, stride, 1, outpad)
    return F.relu(_bn(y, P, name + ".1.", training))


def cost_reg_mvsnet(x: Tensor, P: Dict[str, Tensor], training: bool = False,
                    taps: Dict[str, Tensor] | None = None) -> Tensor:
    """CostRegNet of MVSNet: [B,32,D,H,W] -> [B,1,D,H,W].  jdacs/models/mvsnet.py:37-74."""
    c0 = _cbr3(x, P, "conv0", 1, training)
    c1 = _cbr3(c0, P, "conv1", 2, training)
    c2 = _cbr3(c1, P, "conv2", 1, training)
    c3 = _cbr3(c2, P, "conv3", 2, training)
    c4 = _cbr3(c3, P, "conv4", 1, training)
    c5 = _cbr3(c4, P, "conv5", 2, training)
    c6 = _cbr3(c5, P, "conv6", 1, training)
    u7 = c4 + _dbr3(c6, P, "conv7", 2, 1, training)
    u9 = c2 + _dbr3(u7, P, "conv9", 2, 1, training)
    u11 = c0 + _dbr3(u9, P, "conv11", 2, 1, training)
    out = F.conv3d(u11, P["prob.weight"], P["prob.bias"], 1, 1)
    if taps is not None:
        taps.update(conv0=c0, conv1=c1, conv2=c2, conv3=c3, conv4=c4, conv5=c5, conv6=c6, up7=u7, up9=u9, up11=u11)
    return out


def cost_reg_cvp(x: Tensor, P: Dict[str, Tensor], training: bool = False) -> Tensor:
    """CostRegNet of CVP-MVSNet: [B,16,D,H,W] -> [B,D,H,W].  jdacs-ms/models/network.py:44-74."""
    c0 = _cbr3(_cbr3(x, P, "conv0", 1, training), P, "conv0a", 1, training)
    c2 = _cbr3(_cbr3(_cbr3(c0, P, "conv1", 2, training), P, "conv2", 1, training), P, "conv2a", 1, training)
    c4 = _cbr3(_cbr3(_cbr3(c2, P, "conv3", 1, training), P, "conv4", 1, training), P, "conv4a", 1, training)
    c5 = c2 + _dbr3(c4, P, "conv5", 1, 0, training)
    c6 = c0 + _dbr3(c5, P, "conv6", 2, 1, training)
    return F.conv3d(c6, P["prob0.weight"], P["prob0.bias"], 1, 1).squeeze(1)


# --------------------------------------------------------------------------------------------
# a7 / a8: softmax + soft-argmin + photometric confidence
# --------------------------------------------------------------------------------------------
def soft_argmin(cost_reg: Tensor, depth: Tensor) -> Tuple[Tensor, Tensor]:
    """prob [B,D,H,W], depth [B,H,W].  jdacs/models/mvsnet.py:141-143, module.py:145-148;
    depth [B,D,H,W] is depth_regression_refine, jdacs-ms/models/modules.py:330-331."""
    prob = F.softmax(cost_reg, dim=1)
    dv = depth.reshape(*depth.shape, 1, 1) if depth.dim() == 2 else depth
    return prob, torch.sum(prob * dv, 1)


def photometric_confidence(prob: Tensor) -> Tuple[Tensor, Tensor]:
    """(index int64 [B,H,W], confidence [B,H,W]).  jdacs/models/mvsnet.py:145-151:
    S_d = p[d-1]+p[d]+p[d+1]+p[d+2] (zero padded), index = trunc(sum_d p_d * d), conf = S[index]."""
    nd = prob.shape[1]
    win = 4 * F.avg_pool3d(F.pad(prob.unsqueeze(1), pad=(0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1, padding=0).squeeze(1)
    ramp = torch.arange(nd, dtype=torch.float32).reshape(1, nd, 1, 1)
    index = torch.sum(prob * ramp, 1).long()
    return index, torch.gather(win, 1, index.unsqueeze(1)).squeeze(1)


# --------------------------------------------------------------------------------------------
# 2-D feature extractors (stay library code in the product; needed here for whole-forward parity)
# --------------------------------------------------------------------------------------------
def feature_net(img: Tensor, P: Dict[str, Tensor], training: bool = False) -> Tensor:
    """FeatureNet, jdacs/models/mvsnet.py:17-34 (P keys relative to `feature.`)."""
    def cbr(x, name, k, s, p):
        y = F.conv2d(x, P[name + ".conv.weight"], None, s, p)
        return F.relu(_bn(y, P, name + ".bn.", training))
    x = cbr(cbr(img, "conv0", 3, 1, 1), "conv1", 3, 1, 1)
    x = cbr(cbr(cbr(x, "conv2", 5, 2, 2), "conv3", 3, 1, 1), "conv4", 3, 1, 1)
    x = cbr(cbr(x, "conv5", 5, 2, 2), "conv6", 3, 1, 1)
    return F.conv2d(x, P["feature.weight"], P["feature.bias"], 1, 1)


_PYR = ("conv0aa", "conv0ba", "conv0bb", "conv0bc", "conv0bd", "conv0be", "conv0bf", "conv0bg", "conv0bh")


def feature_pyramid(img: Tensor, P: Dict[str, Tensor], scales: int) -> List[Tensor]:
    """FeaturePyramid, jdacs-ms/models/network.py:16-41: same 9 convs at every x0.5 image level."""
    out = []
    for s in range(scales):
        if s > 0:
            img = F.interpolate(img, scale_factor=0.5, mode="bilinear", align_corners=None).detach()
        f = img
        for name in _PYR:
            f = F.leaky_relu(F.conv2d(f, P[name + ".0.weight"], P[name + ".0.bias"], 1, 1), 0.1)
        out.append(f)
    return out


def _sub(P: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    return {k[len(prefix):]: v for k, v in P.items() if k.startswith(prefix)}


# --------------------------------------------------------------------------------------------
# whole forward passes
# --------------------------------------------------------------------------------------------
def mvsnet_forward(imgs: Tensor, proj_matrices: Tensor, depth_values: Tensor, P: Dict[str, Tensor],
                   training: bool = False, align_corners: bool = False, stages: Dict[str, Tensor] | None = None
                   ) -> Dict[str, Tensor]:
    """MVSNet.forward with refine=False.  jdacs/models/mvsnet.py:105-155."""
    views = imgs.shape[1]
    fp = _sub(P, "feature.")
    feats = [feature_net(imgs[:, v], fp, training) for v in range(views)]
    var = variance_volume(feats[0], feats[1:], proj_matrices[:, 0], [proj_matrices[:, v] for v in range(1, views)],
                          depth_values, False, align_corners, inplace=not training and not torch.is_grad_enabled())
    reg = cost_reg_mvsnet(var, _sub(P, "cost_regularization."), training).squeeze(1)
    prob, depth = soft_argmin(reg, depth_values)
    with torch.no_grad():
        index, conf = photometric_confidence(prob)
    if stages is not None:
        stages.update(features=torch.stack(feats), variance=var, cost_reg=reg, prob=prob, index=index)
    return {"depth": depth, "photometric_confidence": conf}


def sweeping_depth_hypos(depth_min: Tensor, depth_max: Tensor, batch: int, n: int = 48) -> Tensor:
    """Exactly n uniform planes d_k = dmin + k (dmax-dmin)/(n-1), first batch item's range for all.

    jdacs-ms/models/modules.py:44-59.  The reference uses the deprecated inclusive torch.range,
    which drops the last plane for some ranges under fp32 rounding (hazard H3); the intended
    n-plane sweep is what is restated here, and fixtures use a range where both agree."""
    step = (depth_max[0] - depth_min[0]) / (n - 1)
    planes = depth_min[0] + step * torch.arange(n, dtype=torch.float32)
    return planes.unsqueeze(0).repeat(batch, 1)


def condition_intrinsics(intr: Tensor, level: int) -> Tensor:
    """K[:2] / 2^level.  jdacs-ms/models/modules.py:22-37."""
    out = intr.clone()
    out[..., :2, :] = out[..., :2, :] / float(2 ** level)
    return out


def depth_hypos_refine(depth_up: Tensor, ref_in: Tensor, src_in0: Tensor, ref_ex: Tensor, src_ex0: Tensor,
                       half: int = 4) -> Tensor:
    """[B,2*half,H,W] hypotheses depth_up + k*interval, k=-half..half-1.  jdacs-ms/models/modules.py:107-206 (a9).

    interval (one scalar per batch item) = mean over pixels of |delta_d|, the depth change that
    moves the projection into source view 0 by one pixel along the epipolar line, solved per pixel
    in fp64 from the 2x2 system of modules.py:185-194."""
    b, h, w = depth_up.shape
    out = depth_up.unsqueeze(1).repeat(1, 2 * half, 1, 1).double()
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float64), torch.arange(w, dtype=torch.float64), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, dtype=torch.float64)))  # [3,P]
    for i in range(b):
        kr, ks = ref_in[i].double(), src_in0[i].double()
        er, es = ref_ex[i].double(), src_ex0[i].double()
        d1 = depth_up[i].reshape(-1).double()

        def to_src(depth):
            cam = torch.inverse(kr) @ (pix * depth)
            world = torch.inverse(er) @ torch.cat((cam, torch.ones(1, h * w, dtype=torch.float64)))
            p = ks @ (es @ world)[:3]
            return p / p[2:3], p[2]
        x1, x1z = to_src(d1)
        x2, _ = to_src(d1 + 1)
        theta = torch.atan((x2[1] - x1[1]) / (x2[0] - x1[0]))
        x3 = x1 + torch.stack((torch.cos(theta), torch.sin(theta), torch.zeros_like(theta)))
        a = kr @ er[:3, :3] @ torch.inverse(ks @ es[:3, :3])
        t1 = x1z * (a @ x1)
        t2 = a @ x3
        # rows 1,2 of [pix | t2] (a, b)^T = rows 1,2 of t1  ->  a = delta_d
        det = pix[1] * t2[2] - t2[1] * pix[2]
        delta = (t1[1] * t2[2] - t2[1] * t1[2]) / det
        interval = delta.abs().mean()
        for k in range(-half, half):
            out[i, k + half] += k * interval
    return out.float()


def cvp_forward(inp: Dict[str, Tensor], P: Dict[str, Tensor], nscale: int, training: bool = False,
                align_corners: bool = False, stages: Dict[str, Tensor] | None = None) -> Dict[str, object]:
    """CVPMVSNet.forward.  jdacs-ms/models/network.py:84-199."""
    ref_img, src_imgs = inp["ref_img"], inp["src_imgs"]
    b, nsrc = src_imgs.shape[0], src_imgs.shape[1]
    fp, rp = _sub(P, "featurePyramid."), _sub(P, "cost_reg_refine.")
    ref_pyr = feature_pyramid(ref_img, fp, nscale)
    src_pyr = [feature_pyramid(src_imgs[:, i], fp, nscale) for i in range(nsrc)]
    top = nscale - 1
    hyp = sweeping_depth_hypos(inp["depth_min"], inp["depth_max"], b)
    rproj = compose_projection(condition_intrinsics(inp["ref_in"], top), inp["ref_ex"])
    sproj = [compose_projection(condition_intrinsics(inp["src_in"][:, i], top), inp["src_ex"][:, i]) for i in range(nsrc)]
    var = variance_volume(ref_pyr[top], [p[top] for p in src_pyr], rproj, sproj, hyp, True, align_corners)
    reg = cost_reg_cvp(var, rp, training)
    prob, depth = soft_argmin(reg, hyp)
    ests = [depth]
    if stages is not None:
        stages.update(variance0=var, cost_reg0=reg, hypos0=hyp)
    for level in range(nscale - 2, -1, -1):
        up = F.interpolate(depth[None, :], size=None, scale_factor=2, mode="bilinear", align_corners=None).squeeze(0)
        kr = condition_intrinsics(inp["ref_in"], level)
        ks = [condition_intrinsics(inp["src_in"][:, i], level) for i in range(nsrc)]
        with torch.no_grad():
            hyp = depth_hypos_refine(up, kr, ks[0], inp["ref_ex"], inp["src_ex"][:, 0])
        rproj = compose_projection(kr, inp["ref_ex"])
        sproj = [compose_projection(ks[i], inp["src_ex"][:, i]) for i in range(nsrc)]
        var = variance_volume(ref_pyr[level], [p[level] for p in src_pyr], rproj, sproj, hyp, True, align_corners)
        reg = cost_reg_cvp(var, rp, training)
        prob, depth = soft_argmin(reg, hyp)
        ests.append(depth)
        if stages is not None:
            stages["hypos_l%d" % level] = hyp
            stages["variance_l%d" % level] = var
    with torch.no_grad():
        _, conf = photometric_confidence(prob)
    ests.reverse()
    return {"depth_est_list": ests, "prob_confidence": conf}


# --------------------------------------------------------------------------------------------
# a10 / a11: photometric inverse warp and the self-supervised loss
# --------------------------------------------------------------------------------------------
def inverse_warping(img: Tensor, left_cam: Tensor, right_cam: Tensor, depth: Tensor) -> Tuple[Tensor, Tensor]:
    """Warp source image `img` [B,H,W,C] into the reference view through `depth` [B,H,W].

    jdacs/losses/homography.py:186-238 with the sampler of :292-374.  Faithful to two quirks:
    the projection uses the REFERENCE intrinsics on both sides (hazard H6) and the validity mask
    tests y0 <= H-1 while the weights use clamped x1, y1 (hazard H7)."""
    b, h, w, c = img.shape
    k_ref = left_cam[:, 1, :3, :3]
    r_l, t_l = left_cam[:, 0, :3, :3], left_cam[:, 0, :3, 3:4]
    r_r, t_r = right_cam[:, 0, :3, :3], right_cam[:, 0, :3, 3:4]
    r_rel = r_r @ r_l.transpose(1, 2)
    t_rel = t_r - r_rel @ t_l
    # pixel grid exactly as _meshgrid_abs builds it (linspace then rescale; hazard H8)
    gx = (torch.linspace(-1.0, 1.0, w).unsqueeze(0).expand(h, w) + 1.0) * 0.5 * (w - 1)
    gy = (torch.linspace(-1.0, 1.0, h).unsqueeze(1).expand(h, w) + 1.0) * 0.5 * (h - 1)
    pix = torch.stack((gx.reshape(-1), gy.reshape(-1), torch.ones(h * w)))
    cam = (torch.inverse(k_ref) @ pix.unsqueeze(0)) * depth.reshape(b, 1, h * w)
    cam_h = torch.cat((cam, torch.ones(b, 1, h * w)), 1)
    k_hom = torch.zeros(b, 4, 4)
    k_hom[:, :3, :3] = k_ref
    k_hom[:, 3, 3] = 1.0
    m_rel = torch.zeros(b, 4, 4)
    m_rel[:, :3, :3], m_rel[:, :3, 3:4], m_rel[:, 3, 3] = r_rel, t_rel, 1.0
    p = (k_hom @ m_rel) @ cam_h
    u = p[:, 0] / (p[:, 2] + 1e-10)
    v = p[:, 1] / (p[:, 2] + 1e-10)
    # _spatial_transformer normalises, _bilinear_sample un-normalises (homography.py:283-286, 324-325)
    x = ((u / (w - 1) * 2.0 - 1.0) + 1.0) * (w - 1.0) / 2.0
    y = ((v / (h - 1) * 2.0 - 1.0) + 1.0) * (h - 1.0) / 2.0
    x, y = x.reshape(-1), y.reshape(-1)
    x0 = torch.floor(x).int()
    y0 = torch.floor(y).int()
    x1, y1 = x0 + 1, y0 + 1
    mask = ((x0 >= 0) & (x1 <= w - 1) & (y0 >= 0) & (y0 <= h - 1)).float()
    x0, x1 = x0.clamp(0, w - 1), x1.clamp(0, w - 1)
    y0, y1 = y0.clamp(0, h - 1), y1.clamp(0, h - 1)
    base = (torch.arange(b) * (h * w)).reshape(b, 1).expand(b, h * w).reshape(-1)
    flat = img.reshape(-1, c).float()
    pa = flat[base + y0.long() * w + x0.long()]
    pb = flat[base + y1.long() * w + x0.long()]
    pc = flat[base + y0.long() * w + x1.long()]
    pd = flat[base + y1.long() * w + x1.long()]
    fx, fy = x1.float() - x, y1.float() - y
    out = (fx * fy).unsqueeze(1) * pa + (fx * (1 - fy)).unsqueeze(1) * pb + ((1 - fx) * fy).unsqueeze(1) * pc \
        + ((1 - fx) * (1 - fy)).unsqueeze(1) * pd
    return out.reshape(b, h, w, c), mask.reshape(b, h, w, 1)


def ssim_map(x: Tensor, y: Tensor, mask: Tensor) -> Tensor:
    """3x3 SSIM dissimilarity, NHWC in/out.  jdacs/losses/modules.py:17-52."""
    x, y, m = x.permute(0, 3, 1, 2), y.permute(0, 3, 1, 2), mask.permute(0, 3, 1, 2)
    pool = lambda t: F.avg_pool2d(t, 3, 1)
    mx, my = pool(x), pool(y)
    sx, sy, sxy = pool(x * x) - mx * mx, pool(y * y) - my * my, pool(x * y) - mx * my
    num = (2 * mx * my + 0.01 ** 2) * (2 * sxy + 0.03 ** 2)
    den = (mx * mx + my * my + 0.01 ** 2) * (sx + sy + 0.03 ** 2)
    return (pool(m) * torch.clamp((1 - num / den) / 2, 0, 1)).permute(0, 2, 3, 1)


def reconstr_loss(warped: Tensor, ref: Tensor, mask: Tensor) -> Tensor:
    """0.5*smoothL1(photo) + 0.5*(smoothL1(dx)+smoothL1(dy)).  jdacs/losses/modules.py:80-90."""
    a, r = warped * mask, ref * mask
    gx = lambda t: t[:, :, 1:, :] - t[:, :, :-1, :]
    gy = lambda t: t[:, 1:, :, :] - t[:, :-1, :, :]
    return 0.5 * F.smooth_l1_loss(a, r) + 0.5 * (F.smooth_l1_loss(gx(a), gx(r)) + F.smooth_l1_loss(gy(a), gy(r)))


def depth_smoothness(depth: Tensor, img: Tensor, lam: float = 1.0) -> Tensor:
    """Edge-aware first-order smoothness on NHWC maps.  jdacs/losses/modules.py:55-77."""
    dx = lambda t: t[:, :, :-1, :] - t[:, :, 1:, :]
    dy = lambda t: t[:, :-1, :, :] - t[:, 1:, :, :]
    wx = torch.exp(-(lam * dx(img).abs().mean(3, keepdim=True)))
    wy = torch.exp(-(lam * dy(img).abs().mean(3, keepdim=True)))
    return (dx(depth) * wx).abs().mean() + (dy(depth) * wy).abs().mean()


def unsup_loss(imgs: Tensor, cams: Tensor, depth: Tensor, downscale: bool = True, w_smooth: float = 0.18,
               lam: float = 1.0) -> Dict[str, Tensor]:
    """UnSupLoss.forward.  jdacs/losses/unsup_loss.py:24-83 (downscale, 0.18);
    jdacs-ms/losses/unsup_loss.py:23-86 is downscale=False, w_smooth=0.05."""
    views = imgs.shape[1]
    prep = (lambda t: F.interpolate(t, scale_factor=0.25, mode="bilinear")) if downscale else (lambda t: t)
    ref = prep(imgs[:, 0]).permute(0, 2, 3, 1)
    ssim = 0
    per_view = []
    for v in range(1, views):
        src = prep(imgs[:, v]).permute(0, 2, 3, 1)
        warped, mask = inverse_warping(src, cams[:, 0], cams[:, v], depth)
        per_view.append(reconstr_loss(warped, ref, mask) + 1e4 * (1 - mask))
        if v < 3:
            ssim = ssim + ssim_map(ref, warped, mask).mean()
    smooth = depth_smoothness(depth.unsqueeze(-1), ref, lam)
    vol = torch.stack(per_view).permute(1, 2, 3, 4, 0)
    top, _ = torch.topk(-vol, k=3, sorted=False)
    top = -top
    top = top * (top < 1e4).float()
    rec = top.sum(-1).mean()
    return {"reconstr": rec, "ssim": ssim, "smooth": smooth, "total": 12 * rec + 6 * ssim + w_smooth * smooth}
