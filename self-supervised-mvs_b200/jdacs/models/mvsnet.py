"""Mirror of the reference's `models/mvsnet.py` (jdacs/models/mvsnet.py): FeatureNet, CostRegNet, RefineNet,
MVSNet, mvsnet_loss — same constructor / forward signatures, sub-module names and state-dict keys, so that
`from models.mvsnet import MVSNet, mvsnet_loss` (jdacs/train.py:28) and reference checkpoints keep working.

What runs where:
  FeatureNet, RefineNet      2-D convolutions, library (cuDNN) code — outside the plane-sweep path
  cost volume                ONE fused kernel: homography warp + bilinear gather + running variance
  CostRegNet                 3x3x3 (transposed) convolutions with BN/ReLU/skip fused in the epilogue
  softmax / depth / index / confidence   ONE fused kernel
"""
from __future__ import annotations

import warnings

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops, regnet
from .module import ALIGN_CORNERS, ConvBnReLU, ConvBnReLU3D


class FeatureNet(nn.Module):
    """jdacs/models/mvsnet.py:17-34."""

    def __init__(self):
        super().__init__()
        self.inplanes = 32
        self.conv0 = ConvBnReLU(3, 8, 3, 1, 1)
        self.conv1 = ConvBnReLU(8, 8, 3, 1, 1)
        self.conv2 = ConvBnReLU(8, 16, 5, 2, 2)
        self.conv3 = ConvBnReLU(16, 16, 3, 1, 1)
        self.conv4 = ConvBnReLU(16, 16, 3, 1, 1)
        self.conv5 = ConvBnReLU(16, 32, 5, 2, 2)
        self.conv6 = ConvBnReLU(32, 32, 3, 1, 1)
        self.feature = nn.Conv2d(32, 32, 3, 1, 1)
        self.fused_front = True     # forward_maps: conv0 + conv1 + conv2 as one kernel (False: one tcgen05 launch per layer)

    def forward(self, x):
        x = self.conv1(self.conv0(x))
        x = self.conv4(self.conv3(self.conv2(x)))
        return self.feature(self.conv6(self.conv5(x)))

    def train(self, mode=True):
        self.__dict__.pop("_folded", None)     # weights are about to change (or have): drop everything derived from them
        return super().train(mode)

    def forward_folded(self, x, dtype):
        """Eval-mode fast path (library code): BatchNorm folded into the convolution weights, `dtype` channels-last
        activations, so each layer is one cuDNN tensor-core convolution + ReLU.  Same arithmetic as forward() in eval
        mode up to the rounding of `dtype`."""
        key = (dtype, x.device)
        sig = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        keep = regnet.owned(self)      # DataParallel replicas recompute (their tensors come and go): regnet.owned
        hit = getattr(self, "_folded", {}).get(key) if keep else None
        if hit is None or hit[0] != sig:
            layers = []
            for blk in (self.conv0, self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6):
                scale, shift = ops.fold_bn(blk.bn)
                w = (blk.conv.weight.detach().float() * scale.view(-1, 1, 1, 1)).to(dtype).contiguous(memory_format=torch.channels_last)
                layers.append((w, shift.to(dtype), blk.conv.stride, blk.conv.padding, True))
            layers.append((self.feature.weight.detach().to(dtype).contiguous(memory_format=torch.channels_last),
                           self.feature.bias.detach().to(dtype), self.feature.stride, self.feature.padding, False))
            hit = (sig, layers)
            if keep:
                self.__dict__.setdefault("_folded", {})[key] = hit
        y = x.to(dtype).contiguous(memory_format=torch.channels_last)
        fused = getattr(torch, "cudnn_convolution_relu", None) if y.is_cuda else None
        for w, b, stride, pad, relu in hit[1]:
            if relu and fused is not None:
                y = fused(y, w, b, stride, pad, (1, 1), 1)      # cuDNN conv + bias + ReLU in one kernel
            else:
                y = F.conv2d(y, w, b, stride, pad)
                if relu:
                    y = F.relu_(y)
        return y


    def forward_maps(self, imgs, dtype):
        """Eval-mode path on the repo's own tcgen05 kernel: imgs [B,N,3,H,W] fp32 -> zero-bordered C8P feature maps of all
        views [N,B,C/8,H/4+3,W/4+2,8] in `dtype`, the layout the fused plane sweep gathers from.  Every layer is one
        mvs_conv2d_fwd launch (BatchNorm folded to the epilogue affine, ReLU fused); nothing goes through cuDNN."""
        dev = imgs.device
        sig = tuple((p.data_ptr(), p._version) for p in list(self.parameters()) + list(self.buffers()))
        key = ("tc", dtype, dev)
        keep = regnet.owned(self)
        hit = self.__dict__.setdefault("_folded", {}).get(key) if keep else None
        if hit is None or hit[0] != sig:
            layers = []
            for blk in (self.conv0, self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6):
                scale, shift = ops.fold_bn(blk.bn)
                layers.append((ops.pack_conv2d_weight(blk.conv.weight), blk.conv.out_channels, blk.conv.kernel_size[0], blk.conv.stride[0],
                               scale, shift, True))
            layers.append((ops.pack_conv2d_weight(self.feature.weight), self.feature.out_channels, 3, 1, None,
                           self.feature.bias.detach().float().contiguous(), False))
            # the three full-resolution layers as one kernel (csrc/featnet_front.cu): weight fragments + affine table
            front = ops.featnet_front_pack(self.conv0.conv.weight, self.conv1.conv.weight, self.conv2.conv.weight,
                                           [(l[4], l[5]) for l in layers[:3]], dtype)
            hit = (sig, layers, {}, front)
            if keep:
                self._folded[key] = hit
        b, n = imgs.shape[0], imgs.shape[1]
        first = 0
        if self.fused_front:
            x = ops.featnet_front(imgs, hit[3][0], hit[3][1], dtype)
            first = 3
        else:
            x = ops.pack_images_c8(imgs, dtype)
        for i, (g, cout, k, stride, scale, shift, relu) in enumerate(hit[1]):
            if i >= first:
                x = ops.conv2d_raw(x, g, cout, k, stride, scale, shift, relu, out_padded=(i == len(hit[1]) - 1), tile_cache=hit[2])
        return x.view(n, b, *x.shape[1:])


    def forward_train_tc(self, imgs, dtype, frozen=False):
        """Differentiable path on the repo's own kernels (training, or eval-mode fine-tuning with `frozen` statistics):
        imgs [B,N,3,H,W] -> one C8 feature map [B, C/8, H/4, W/4, 8] in `dtype` per view.  The B images of a view form a C8 volume
        whose depth axis is the image index and the N views are the batch entries of that volume; every layer is the 3-D training
        op of CostRegNet with zero kd = 0, 2 taps (ops.conv2d_bn_relu_tc): tcgen05 forward / input gradient, tensor-core weight
        gradient, fp64 batch statistics PER VIEW -- as BatchNorm2d sees them in the reference, which calls the extractor once
        per view (jdacs/models/mvsnet.py:115) -- all views in one launch per layer."""
        b, n, _, h, w = imgs.shape
        if n * 32 > 256:                                   # the BatchNorm kernels take <= 256 (view, channel) pairs: view by view
            return [self.forward_train_tc(imgs[:, v:v + 1], dtype, frozen)[0] for v in range(n)]
        x = ops.pack_images_c8(imgs, dtype).view(n, 1, b, h, w, 8)         # image m = v * B + b
        for blk in (self.conv0, self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6):
            x = ops.conv2d_bn_relu_tc(x, blk.conv, blk.bn, frozen)
        x = ops.conv2d_bias_tc(x, self.feature)                           # [N, C/8, B, H/4, W/4, 8]
        x = x.permute(0, 2, 1, 3, 4, 5).contiguous()
        return [x[v] for v in range(n)]


class CostRegNet(nn.Module):
    """jdacs/models/mvsnet.py:37-74.  forward takes the variance volume (C8 or [B,32,D,H,W]) and returns
    cost_reg: [B,D,H,W] fp32 for a C8 input (internal), [B,1,D,H,W] for a plain input (reference shape).
    D, H, W must be divisible by 8, as in the reference (three stride-2 levels)."""

    def __init__(self):
        super().__init__()
        self.conv0 = ConvBnReLU3D(32, 8)
        self.conv1 = ConvBnReLU3D(8, 16, stride=2)
        self.conv2 = ConvBnReLU3D(16, 16)
        self.conv3 = ConvBnReLU3D(16, 32, stride=2)
        self.conv4 = ConvBnReLU3D(32, 32)
        self.conv5 = ConvBnReLU3D(32, 64, stride=2)
        self.conv6 = ConvBnReLU3D(64, 64)
        self.conv7 = nn.Sequential(
            nn.ConvTranspose3d(64, 32, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
            nn.BatchNorm3d(32), nn.ReLU(inplace=True))
        self.conv9 = nn.Sequential(
            nn.ConvTranspose3d(32, 16, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
            nn.BatchNorm3d(16), nn.ReLU(inplace=True))
        self.conv11 = nn.Sequential(
            nn.ConvTranspose3d(16, 8, kernel_size=3, padding=1, output_padding=1, stride=2, bias=False),
            nn.BatchNorm3d(8), nn.ReLU(inplace=True))
        self.prob = nn.Conv3d(8, 1, 3, stride=1, padding=1)
        self._cache = regnet.PackCache()
        self.algo = 0           # 0 auto, 1 SIMT fp32, 2 tcgen05 (mvs_conv3d_desc.algo)
        self.act_dtype = None   # storage dtype of activations in eval mode; None = dtype of the input volume

    def train(self, mode=True):
        self._cache.clear()
        return super().train(mode)

    def forward(self, x, frozen_grad=False):
        """frozen_grad: eval-mode BatchNorm (running statistics) but differentiable -- fine-tuning under module.eval()."""
        plain = x.dim() == 5
        tr = self.training
        fg = frozen_grad and not tr
        x = regnet.as_c8(x, torch.float32 if (tr or fg) else (self.act_dtype or torch.float32))
        for n in (x.shape[2], x.shape[3], x.shape[4]):
            if n % 8:
                raise ValueError("CostRegNet needs D, H, W divisible by 8 (got %s), as the reference does" % (tuple(x.shape[2:5]),))
        if not tr and not fg and self.act_dtype is not None and x.dtype != self.act_dtype:
            raise ValueError("variance volume is %s but CostRegNet.act_dtype is %s" % (x.dtype, self.act_dtype))
        a = self.algo
        conv0 = self.conv0(x, None, a, fg)
        conv2 = self.conv2(self.conv1(conv0, None, a, fg), None, a, fg)
        conv4 = self.conv4(self.conv3(conv2, None, a, fg), None, a, fg)
        y = self.conv6(self.conv5(conv4, None, a, fg), None, a, fg)
        c = self._cache
        y = regnet.conv_bn_relu(y, self.conv7[0], self.conv7[1], tr, c, conv4, a, fg)    # conv4 + relu(bn(convT(x)))
        y = regnet.conv_bn_relu(y, self.conv9[0], self.conv9[1], tr, c, conv2, a, fg)
        y = regnet.conv_bn_relu(y, self.conv11[0], self.conv11[1], tr, c, conv0, a, fg)
        out = regnet.conv_bias(y, self.prob, tr, c, a, fg)                                # [B,D,H,W] fp32
        return out.unsqueeze(1) if plain else out


class RefineNet(nn.Module):
    """jdacs/models/mvsnet.py:77-92 (2-D, library code; off by default in train.py via --refine False)."""

    def __init__(self):
        super().__init__()
        self.conv1 = ConvBnReLU(4, 32)
        self.conv2 = ConvBnReLU(32, 32)
        self.conv3 = ConvBnReLU(32, 32)
        self.res = ConvBnReLU(32, 1)

    def forward(self, img, depth_init):
        img = F.interpolate(img, scale_factor=0.25, mode='bilinear')
        depth_init = depth_init.unsqueeze(dim=1)
        concat = torch.cat((img, depth_init), dim=1)
        depth_residual = self.res(self.conv3(self.conv2(self.conv1(concat))))
        return (depth_init + depth_residual).squeeze(dim=1)


class MVSNet(nn.Module):
    """jdacs/models/mvsnet.py:95-161.

    forward(imgs [B,N,3,H,W], proj_matrices [B,N,4,4], depth_values [B,D])
        -> {"depth": [B,H/4,W/4], "photometric_confidence": [B,H/4,W/4]}

    Extra, optional knobs:
      volume_dtype   storage of feature maps / cost volume / U-Net activations in eval mode.  None (default) = fp16 on a CUDA device
                     (every stage on the tcgen05 / TMA kernels; depth within the 16-bit storage floor of the fp32 reference, see
                     profiles/r02_parity.json), fp32 on the host-emulation build.  torch.float32 = the reference's arithmetic
                     (fp32 SIMT kernels, 1e-5 relative on depth).
      train_dtype    the same for training mode.  None (default) = bf16 activations on a CUDA device with fp32 master weights,
                     fp32 accumulation and fp64 BatchNorm statistics (tensor-core forward / dgrad / wgrad); torch.float32 = fp32.
      align_corners  True = the geometry the authors intended on torch 1.1 (hazard H1)
    Gradients under module.eval() (fine-tuning with frozen BatchNorm statistics) are supported with 16-bit train_dtype.
    """

    def __init__(self, refine=True, volume_dtype=None, align_corners=ALIGN_CORNERS, train_dtype=None):
        super().__init__()
        self.refine = refine
        self.feature = FeatureNet()
        self.cost_regularization = CostRegNet()
        if self.refine:
            self.refine_network = RefineNet()
        self.volume_dtype = volume_dtype
        self.train_dtype = train_dtype
        self.align_corners = align_corners
        self.feature_autocast = True   # eval + 16-bit volume, feature_tc off: FeatureNet as folded 16-bit library convolutions
        self.feature_tc = True         # 16-bit volumes: FeatureNet on the repo's tcgen05 convolution kernels (eval and training)
        self.keep_index = False        # also return "depth_index" (the truncated expected plane index, mvsnet.py:149-150)

    def forward(self, imgs, proj_matrices, depth_values):
        assert imgs.shape[1] == proj_matrices.shape[1], "Different number of images and projection matrices"
        b, n = imgs.shape[0], imgs.shape[1]
        if n < 2:
            # the reference runs on: one view gives an all-zero variance volume (:135) and a depth map that ignores the images
            raise ValueError("MVSNet.forward needs the reference view and at least one source view (got %d view)" % n)
        hf, wf = ((imgs.shape[-2] - 1) // 2) // 2 + 1, ((imgs.shape[-1] - 1) // 2) // 2 + 1      # two stride-2 layers (:23,:26)
        if depth_values.shape[1] % 8 or hf % 8 or wf % 8:
            # checked before any launch: the 3-D U-Net halves D, H, W three times and adds the skips back (:62-71)
            raise ValueError("CostRegNet needs D, H, W divisible by 8 (got %s), as the reference does" % ((depth_values.shape[1], hf, wf),))
        if b == 0:
            # an empty batch is an empty result, as the reference's library layers return it; the kernels are never launched on it
            empty = imgs.new_zeros((0, hf, wf), dtype=torch.float32)
            out = {"depth": empty, "photometric_confidence": empty.clone()}
            if self.keep_index:
                out["depth_index"] = empty.long()
            return out
        # step 1. feature extraction (library code).  In eval mode all views share one batched call; in training
        # each view is its own call, because BatchNorm2d statistics are per call in the reference (:115).
        auto16 = imgs.is_cuda
        tdt = self.train_dtype if self.train_dtype is not None else (torch.bfloat16 if auto16 else torch.float32)
        # differentiable pass: training mode, or eval mode with gradients requested (fine-tuning with frozen BatchNorm statistics,
        # which the reference supports; wrap inference in torch.no_grad() -- as the reference's scripts do -- to get the fused path)
        wants_grad = (not self.training) and torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())
        if wants_grad and tdt == torch.float32:
            wants_grad = False
            if not getattr(self, "_warned_eval_grad", False):
                self._warned_eval_grad = True
                warnings.warn("MVSNet.eval() with fp32 train_dtype runs the fused inference kernels: no gradient flows through the cost "
                              "volume / CostRegNet.  Use train() or a 16-bit train_dtype for eval-mode fine-tuning.")
        diff = self.training or wants_grad
        dt = tdt if diff else (self.volume_dtype if self.volume_dtype is not None else (torch.float16 if auto16 else torch.float32))
        rt = ops.compose_proj(proj_matrices)
        fast = (not diff) and dt != torch.float32 and imgs.is_cuda and self.feature_tc and imgs.shape[-1] % 4 == 0 and imgs.shape[-2] % 4 == 0
        if imgs.dtype != torch.float32 and not (fast and imgs.dtype == dt):
            imgs = imgs.float()     # 16-bit images (a host pipeline may upload them so) are only consumed as such by the 16-bit fast path
        if diff:
            # each view is its own call in training mode: BatchNorm2d statistics are per call in the reference (:115).
            # With 16-bit activations the (library) feature extractor runs under autocast on channels-last tensors -- cuDNN's
            # tensor-core convolutions and NHWC BatchNorm kernels instead of its fp32 NCHW ones (10 ms -> ~3 ms per item at 512x640)
            if dt != torch.float32 and imgs.is_cuda and self.feature_tc and imgs.shape[-1] % 4 == 0 and imgs.shape[-2] % 4 == 0:
                # the feature extractor on the repo's tensor-core training kernels, emitting C8 maps for the sweep
                features = self.feature.forward_train_tc(imgs, dt, frozen=not self.training)
            elif dt != torch.float32 and imgs.is_cuda and self.feature_autocast:
                with torch.autocast("cuda", dtype=dt):
                    features = [self.feature(imgs[:, v].contiguous(memory_format=torch.channels_last)) for v in range(n)]
            else:
                features = [self.feature(imgs[:, v]) for v in range(n)]
            # step 2. plane sweep: warp + variance, fused (:120-136)
            variance = ops.warp_variance(features[0], features[1:], rt, depth_values, dt, self.align_corners, False)
        else:
            if fast:
                # the feature extractor on the tcgen05 convolution kernel, emitting the gather layout directly
                maps = self.feature.forward_maps(imgs, dt)
            else:
                x = imgs.transpose(0, 1).reshape(n * b, *imgs.shape[2:])
                if dt != torch.float32 and x.is_cuda and self.feature_autocast:
                    # the features are stored in `dt` by the sweep anyway: let the library run its tensor-core kernels
                    f = self.feature.forward_folded(x, dt)
                else:
                    f = self.feature(x)
                # all views go to the zero-bordered gather layout in one launch
                maps = ops.pack_c8_padded(f, dt)
                maps = maps.view(n, b, *maps.shape[1:])
            # step 2. the fused plane sweep
            variance = ops.warp_variance_maps(maps, rt, depth_values, dt, self.align_corners, False)
        # step 3. regularisation (:139-141)
        self.cost_regularization.act_dtype = None if diff else dt
        cost_reg = self.cost_regularization(variance, frozen_grad=diff and not self.training)
        # softmax + regression + confidence, fused (:142-151)
        depth, index, photometric_confidence, _ = ops.soft_argmin(cost_reg, depth_values)
        if self.refine:
            depth = self.refine_network(imgs[:, 0].float(), depth)
        out = {"depth": depth, "photometric_confidence": photometric_confidence}
        if self.keep_index:
            out["depth_index"] = index
        return out


def mvsnet_loss(depth_est, depth_gt, mask):
    """jdacs/models/mvsnet.py:164-166 (supervised loss, unused by the self-supervised training)."""
    mask = mask > 0.5
    return F.smooth_l1_loss(depth_est[mask], depth_gt[mask], reduction='mean')
