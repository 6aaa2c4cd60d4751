"""Mirror of the reference's jdacs-ms `models/network.py`: FeaturePyramid, CostRegNet, CVPMVSNet, sL1_loss, MSE_loss.

Same constructor / forward signatures, sub-module names (`featurePyramid`, `cost_reg_refine`) and state-dict keys.
The coarse sweep and every refinement level use the same three fused kernels as MVSNet (warp+variance with
shared or per-pixel hypotheses, the 3-D convolutions, soft-argmin); FeaturePyramid stays library code.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops, regnet
from .modules import (ALIGN_CORNERS, ConvBnReLU3D, calDepthHypo, calSweepingDepthHypo, conditionIntrinsics, conv,
                      proj_cost)


class FeaturePyramid(nn.Module):
    """jdacs-ms/models/network.py:16-41 (2-D, library code)."""

    def __init__(self):
        super().__init__()
        self.conv0aa = conv(3, 64, kernel_size=3, stride=1)
        self.conv0ba = conv(64, 64, kernel_size=3, stride=1)
        self.conv0bb = conv(64, 64, kernel_size=3, stride=1)
        self.conv0bc = conv(64, 32, kernel_size=3, stride=1)
        self.conv0bd = conv(32, 32, kernel_size=3, stride=1)
        self.conv0be = conv(32, 32, kernel_size=3, stride=1)
        self.conv0bf = conv(32, 16, kernel_size=3, stride=1)
        self.conv0bg = conv(16, 16, kernel_size=3, stride=1)
        self.conv0bh = conv(16, 16, kernel_size=3, stride=1)

    def _trunk(self, img):
        f = self.conv0aa(img)
        return self.conv0bh(self.conv0bg(self.conv0bf(self.conv0be(self.conv0bd(self.conv0bc(self.conv0bb(self.conv0ba(f))))))))

    LAYERS = ("conv0aa", "conv0ba", "conv0bb", "conv0bc", "conv0bd", "conv0be", "conv0bf", "conv0bg", "conv0bh")

    def forward_maps(self, imgs, scales, dtype):
        """Eval-mode path on the repo's tcgen05 kernel: imgs [B,N,3,H,W] fp32 (view 0 = reference) -> per pyramid level the
        zero-bordered C8P feature maps of all views [N,B,2,Hl+3,Wl+2,8] in `dtype` (finest level first), the layout the fused
        plane sweep gathers from.  Nine mvs_conv2d_fwd launches per level (bias + LeakyReLU(0.1) in the epilogue); the image
        pyramid itself is the reference's bilinear x0.5 resize (library code, 3 channels)."""
        dev = imgs.device
        sig = tuple((p.data_ptr(), p._version) for p in self.parameters())
        key = ("tc", dtype, dev)
        keep = regnet.owned(self)      # DataParallel replicas recompute (their tensors come and go): regnet.owned
        hit = self.__dict__.setdefault("_packed", {}).get(key) if keep else None
        if hit is None or hit[0] != sig:
            layers = []
            for name in self.LAYERS:
                c = getattr(self, name)[0]
                layers.append((ops.pack_conv2d_weight(c.weight), c.out_channels, c.bias.detach().float().contiguous()))
            hit = (sig, layers, {})
            if keep:
                self._packed[key] = hit
        b, n = imgs.shape[0], imgs.shape[1]
        out = []
        cur = imgs
        for level in range(scales):
            if level:
                flat = F.interpolate(cur.reshape(b * n, *cur.shape[2:]), scale_factor=0.5, mode='bilinear', align_corners=None)
                cur = flat.reshape(b, n, *flat.shape[1:])
            x = ops.pack_images_c8(cur, dtype)
            for i, (g, cout, bias) in enumerate(hit[1]):
                x = ops.conv2d_raw(x, g, cout, 3, 1, None, bias, 0.1, out_padded=(i == len(hit[1]) - 1), tile_cache=hit[2])
            out.append(x.view(n, b, *x.shape[1:]))
        return out

    def train(self, mode=True):
        self.__dict__.pop("_packed", None)
        return super().train(mode)

    def forward(self, img, scales=5):
        fp = [self._trunk(img)]
        for _ in range(scales - 1):
            img = F.interpolate(img, scale_factor=0.5, mode='bilinear', align_corners=None).detach()
            fp.append(self._trunk(img))
        return fp


class CostRegNet(nn.Module):
    """jdacs-ms/models/network.py:44-74.  C8 (or [B,16,D,H,W]) in, [B,D,H,W] fp32 out.  D, H, W must be even."""

    def __init__(self):
        super().__init__()
        self.conv0 = ConvBnReLU3D(16, 16, kernel_size=3, pad=1)
        self.conv0a = ConvBnReLU3D(16, 16, kernel_size=3, pad=1)
        self.conv1 = ConvBnReLU3D(16, 32, stride=2, kernel_size=3, pad=1)
        self.conv2 = ConvBnReLU3D(32, 32, kernel_size=3, pad=1)
        self.conv2a = ConvBnReLU3D(32, 32, kernel_size=3, pad=1)
        self.conv3 = ConvBnReLU3D(32, 64, kernel_size=3, pad=1)
        self.conv4 = ConvBnReLU3D(64, 64, kernel_size=3, pad=1)
        self.conv4a = ConvBnReLU3D(64, 64, kernel_size=3, pad=1)
        self.conv5 = nn.Sequential(
            nn.ConvTranspose3d(64, 32, kernel_size=3, padding=1, output_padding=0, stride=1, bias=False),
            nn.BatchNorm3d(32), nn.ReLU(inplace=True))
        self.conv6 = nn.Sequential(
            nn.ConvTranspose3d(32, 16, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
            nn.BatchNorm3d(16), nn.ReLU(inplace=True))
        self.prob0 = nn.Conv3d(16, 1, 3, stride=1, padding=1)
        self._cache = regnet.PackCache()
        self.algo = 0
        self.act_dtype = None

    def train(self, mode=True):
        self._cache.clear()
        return super().train(mode)

    def forward(self, x, frozen_grad=False):
        tr = self.training
        fg = frozen_grad and not tr
        x = regnet.as_c8(x, torch.float32 if (tr or fg) else (self.act_dtype or torch.float32))
        for n in (x.shape[2], x.shape[3], x.shape[4]):
            if n % 2:
                raise ValueError("CVP CostRegNet needs even D, H, W (got %s), as the reference does" % (tuple(x.shape[2:5]),))
        a, c = self.algo, self._cache
        conv0 = self.conv0a(self.conv0(x, None, a, fg), None, a, fg)
        conv2 = self.conv2a(self.conv2(self.conv1(conv0, None, a, fg), None, a, fg), None, a, fg)
        conv4 = self.conv4a(self.conv4(self.conv3(conv2, None, a, fg), None, a, fg), None, a, fg)
        conv5 = regnet.conv_bn_relu(conv4, self.conv5[0], self.conv5[1], tr, c, conv2, a, fg)
        conv6 = regnet.conv_bn_relu(conv5, self.conv6[0], self.conv6[1], tr, c, conv0, a, fg)
        return regnet.conv_bias(conv6, self.prob0, tr, c, a, fg)


class CVPMVSNet(nn.Module):
    """jdacs-ms/models/network.py:77-199.

    forward(ref_img [B,3,H,W], src_imgs [B,nsrc,3,H,W], ref_in [B,3,3], src_in [B,nsrc,3,3], ref_ex [B,4,4],
            src_ex [B,nsrc,4,4], depth_min [B], depth_max [B])
        -> {"depth_est_list": [finest ... coarsest], "prob_confidence": [B,H,W]}
    args needs .nsrc, .nscale, .mode like the reference's argparse namespace.
    volume_dtype / train_dtype: storage of maps, volumes and activations in eval / training mode; None (default) = fp16 / bf16 on a
    CUDA device (tensor-core kernels), fp32 on the host-emulation build; torch.float32 = the reference's arithmetic (see MVSNet)."""

    def __init__(self, args, volume_dtype=None, train_dtype=None):
        super().__init__()
        self.featurePyramid = FeaturePyramid()
        self.cost_reg_refine = CostRegNet()
        self.args = args
        self.volume_dtype = volume_dtype
        self.train_dtype = train_dtype
        self.feature_tc = True      # eval + 16-bit volumes: FeaturePyramid on the repo's tcgen05 convolution kernel
        self.keep_index = False     # also return "depth_index" of the finest level (network.py:187-188)

    def forward(self, ref_img, src_imgs, ref_in, src_in, ref_ex, src_ex, depth_min, depth_max):
        nsrc, nscale = self.args.nsrc, self.args.nscale
        depth_est_list = []
        ref_img, src_imgs = ref_img.float(), src_imgs.float()    # 16-bit uploads are widened here: the image pyramid is built in fp32
        auto16 = ref_img.is_cuda
        tdt = self.train_dtype if self.train_dtype is not None else (torch.bfloat16 if auto16 else torch.float32)
        # differentiable pass: training mode, or eval mode with gradients requested and a 16-bit train_dtype (frozen BatchNorm)
        diff = self.training or (torch.is_grad_enabled() and tdt != torch.float32 and any(p.requires_grad for p in self.parameters()))
        fg = diff and not self.training
        dt = tdt if diff else (self.volume_dtype if self.volume_dtype is not None else (torch.float16 if auto16 else torch.float32))
        self.cost_reg_refine.act_dtype = None if diff else dt

        fast = (not diff) and dt != torch.float32 and ref_img.is_cuda and self.feature_tc
        if fast:
            # eval, 16-bit volumes: the pyramid on the tcgen05 kernel, every level already in the sweep's gather layout
            imgs = torch.cat((ref_img.unsqueeze(1), src_imgs[:, :nsrc]), 1)
            maps = self.featurePyramid.forward_maps(imgs, nscale, dt)
            shapes = [(m.shape[3] - 3, m.shape[4] - 2) for m in maps]
        else:
            ref_pyr = self.featurePyramid(ref_img, nscale)
            src_pyrs = [self.featurePyramid(src_imgs[:, i], nscale) for i in range(nsrc)]
            shapes = [tuple(f.shape[2:]) for f in ref_pyr]
        fp_shapes = [(0, 0) + s for s in shapes]
        ref_in_ms = conditionIntrinsics(ref_in, ref_img.shape, fp_shapes)
        src_in_ms = torch.stack([conditionIntrinsics(src_in[:, i], ref_img.shape, fp_shapes)
                                 for i in range(nsrc)]).permute(1, 0, 2, 3, 4)  # [B,nsrc,nscale,3,3]

        # coarsest level: fronto-parallel sweep (network.py:110-148)
        depth_hypos = calSweepingDepthHypo(ref_in_ms[:, -1], src_in_ms[:, 0, -1], ref_ex, src_ex, depth_min, depth_max)
        rt = ops.compose_proj_ke(ref_in_ms[:, -1], src_in_ms[:, :, -1], ref_ex, src_ex[:, :nsrc], 1.0)
        if fast:
            cost_volume = ops.warp_variance_maps(maps[-1], rt, depth_hypos, dt, ALIGN_CORNERS, True)
        else:
            cost_volume = ops.warp_variance(ref_pyr[-1], [p[-1] for p in src_pyrs], rt, depth_hypos, dt, ALIGN_CORNERS, True)
        cost_reg = self.cost_reg_refine(cost_volume, frozen_grad=fg)
        depth, index, conf, _ = ops.soft_argmin(cost_reg, depth_hypos)
        depth_est_list.append(depth)

        # refinement up the pyramid (network.py:153-180)
        for level in range(nscale - 2, -1, -1):
            depth_up = F.interpolate(depth[None, :], size=None, scale_factor=2, mode='bilinear', align_corners=None).squeeze(0)
            depth_hypos = calDepthHypo(self.args, depth_up, ref_in_ms[:, level], src_in_ms[:, :, level], ref_ex, src_ex,
                                       depth_min, depth_max, level)
            if fast:
                rt = ops.compose_proj_ke(ref_in_ms[:, level], src_in_ms[:, :nsrc, level], ref_ex, src_ex[:, :nsrc], 1.0)
                cost_volume = ops.warp_variance_maps(maps[level], rt, depth_hypos, dt, ALIGN_CORNERS, True)
            else:
                cost_volume = proj_cost(self.args, ref_pyr[level], src_pyrs, level, ref_in_ms[:, level],
                                        src_in_ms[:, :, level], ref_ex, src_ex, depth_hypos, dt, as_c8=True)
            cost_reg2 = self.cost_reg_refine(cost_volume, frozen_grad=fg)
            depth, index, conf, _ = ops.soft_argmin(cost_reg2, depth_hypos)
            depth_est_list.append(depth)

        depth_est_list.reverse()  # finest first (network.py:195)
        out = {"depth_est_list": depth_est_list, "prob_confidence": conf}
        if self.keep_index:
            out["depth_index"] = index
        return out


def sL1_loss(depth_est, depth_gt, mask):
    """jdacs-ms/models/network.py:202-203."""
    return F.smooth_l1_loss(depth_est[mask], depth_gt[mask], reduction='mean')


def MSE_loss(depth_est, depth_gt, mask):
    """jdacs-ms/models/network.py:206-207."""
    return F.mse_loss(depth_est[mask], depth_gt[mask], reduction='mean')
