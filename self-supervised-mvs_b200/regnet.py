"""Runner for the 3-D regularisation U-Nets over C8 volumes.

The parameter-holding modules keep the reference's names and shapes (checkpoint compatibility);
their arithmetic goes through libmvs_b200:
  * inference (module.eval(), the headline path): BatchNorm is folded to a per-channel affine and fused,
    with ReLU and the skip add, into the convolution epilogue -> one kernel per layer, any storage dtype;
  * training (module.train()): un-activated convolution -> batch statistics -> normalise+ReLU+skip kernels, fp32,
    differentiable (ops._Conv3d / ops._BnAct), running statistics updated like nn.BatchNorm3d.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import ops

Tensor = torch.Tensor


def owned(module: nn.Module) -> bool:
    """True when every parameter of `module` is a leaf nn.Parameter, i.e. the module is the one the user holds.  The replicas
    nn.DataParallel builds per forward (the reference's train.py:65) hold plain tensors instead -- fresh broadcast copies at
    recycled addresses with `_version` 0, and they share the original's __dict__ (replicate() copies it shallowly) -- so nothing
    derived from their weights may be cached: a later replica with NEW weights could otherwise hit a stale entry."""
    return all(isinstance(p, nn.Parameter) for p in module._parameters.values() if p is not None) and \
        all(owned(m) for m in module.children())


class PackCache:
    """Packed weights / folded BN affines of frozen (eval-mode) parameters of the module that owns them.  Entries are keyed by
    (id, data_ptr, _version, device): in-place updates (optimizer steps, load_state_dict) invalidate them; writes through `.data`
    do not bump `_version` -- call clear() (or module.train(); module.eval()) after such an update."""

    def __init__(self) -> None:
        self._store: Dict[Tuple, Tuple] = {}
        self.tiles: Dict[Tuple, Tensor] = {}   # tcgen05 weight tap tiles, keyed by the packed weight (ops.conv3d_raw)

    def clear(self) -> None:
        self._store.clear()
        self.tiles.clear()

    @staticmethod
    def _key(*tensors: Tensor) -> Tuple:
        return tuple((t.data_ptr(), t._version, str(t.device)) for t in tensors)

    def packed(self, weight: Tensor, transposed: bool) -> Tensor:
        k = ("w", id(weight), transposed)
        hit = self._store.get(k)
        ver = self._key(weight)
        if hit is None or hit[0] != ver:
            if len(self._store) > 512:  # replicas (nn.DataParallel) come and go: keep the table bounded
                self._store.clear()
            hit = (ver, ops.pack_conv3d_weight(weight, transposed))
            self._store[k] = hit
            self.tiles.clear()   # a repacked weight may land on a recycled address: never trust old tap tiles
        return hit[1]

    def folded(self, bn: nn.modules.batchnorm._BatchNorm) -> Tuple[Tensor, Tensor]:
        k = ("bn", id(bn))
        ver = self._key(bn.weight, bn.bias, bn.running_mean, bn.running_var)
        hit = self._store.get(k)
        if hit is None or hit[0] != ver:
            hit = (ver, ops.fold_bn(bn))
            self._store[k] = hit
        return hit[1]


def conv_bn_relu(x: Tensor, conv: nn.Module, bn: nn.modules.batchnorm._BatchNorm, training: bool, cache: PackCache,
                 skip: Optional[Tensor] = None, algo: int = 0, frozen_grad: bool = False) -> Tensor:
    """relu(bn(conv(x))) + skip for nn.Conv3d or nn.ConvTranspose3d holders (3x3x3, pad 1, stride 1|2).
    training: batch statistics, differentiable.  frozen_grad: eval-mode statistics but differentiable (fine-tuning under
    module.eval(), which the reference supports); 16-bit volumes only.  Otherwise: the fused inference kernel, no autograd graph."""
    transposed = isinstance(conv, nn.ConvTranspose3d)
    stride = conv.stride[0]
    cout = conv.out_channels
    if (training or frozen_grad) and x.dtype != torch.float32:
        return ops.conv_bn_act_tc(x, conv, bn, skip, frozen=not training)      # tensor-core training path (csrc/train.cu)
    if frozen_grad:
        raise RuntimeError("gradients through an eval-mode CostRegNet need 16-bit volumes (train_dtype=torch.bfloat16); "
                           "the fp32 path is differentiable in train() mode only")
    if training:
        z = ops.conv3d(x, conv.weight, None, stride, transposed)
        return ops.bn_act_train(z, bn, skip, True)
    if not (isinstance(conv.weight, nn.Parameter) and isinstance(bn.weight, nn.Parameter)):   # a DataParallel replica: see owned()
        scale, shift = ops.fold_bn(bn)
        return ops.conv3d_raw(x, ops.pack_conv3d_weight(conv.weight, transposed), cout, stride, transposed, scale, shift, skip,
                              relu=True, algo=algo)
    scale, shift = cache.folded(bn)
    return ops.conv3d_raw(x, cache.packed(conv.weight, transposed), cout, stride, transposed, scale, shift, skip,
                          relu=True, algo=algo, tile_cache=cache.tiles)


def conv_bias(x: Tensor, conv: nn.Conv3d, training: bool, cache: PackCache, algo: int = 0, frozen_grad: bool = False) -> Tensor:
    """The final single-channel `prob` convolution: plain fp32 [B,D,H,W] out."""
    if (training or frozen_grad) and x.dtype != torch.float32:
        return ops.conv_bias_tc(x, conv)
    if training:
        return ops.conv3d(x, conv.weight, conv.bias, 1, False)
    bias = conv.bias.detach().float().contiguous() if conv.bias is not None else None
    if not isinstance(conv.weight, nn.Parameter):
        return ops.conv3d_raw(x, ops.pack_conv3d_weight(conv.weight, False), conv.out_channels, 1, False, None, bias, None,
                              relu=False, algo=algo)
    return ops.conv3d_raw(x, cache.packed(conv.weight, False), conv.out_channels, 1, False, None, bias, None, relu=False,
                          algo=algo, tile_cache=cache.tiles)


def as_c8(x: Tensor, dtype: torch.dtype) -> Tensor:
    """Accept either a C8 volume or the reference's [B,C,D,H,W] fp32 tensor (API parity)."""
    if x.dim() == 6:
        return x
    if x.dim() != 5:
        raise ValueError("expected [B,C,D,H,W] or C8 [B,C/8,D,H,W,8], got %s" % (tuple(x.shape),))
    if x.requires_grad:
        return _PackC8.apply(x)
    return ops.pack_c8(x, dtype)


class _PackC8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor) -> Tensor:
        return ops.pack_c8(x, torch.float32)

    @staticmethod
    def backward(ctx, g: Tensor):
        return ops.unpack_c8(g.contiguous())


class _UnpackC8(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x: Tensor) -> Tensor:
        return ops.unpack_c8(x)

    @staticmethod
    def backward(ctx, g: Tensor):
        return ops.pack_c8(g, torch.float32)


def unpack_c8_grad(x: Tensor) -> Tensor:
    """C8 -> [B,C,*spatial] fp32, differentiable."""
    return _UnpackC8.apply(x)
