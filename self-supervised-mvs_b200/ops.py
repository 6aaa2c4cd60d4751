"""Host-side operators of the plane-sweep path: thin torch.autograd wrappers over the C ABI.

Tensors in the "C8" layout are ordinary torch tensors of shape [B, C/8, H, W, 8] (maps) or
[B, C/8, D, H, W, 8] (volumes); see include/mvs_b200.h.  PyTorch is used for device memory, streams and
autograd bookkeeping only; every arithmetic step of the path is a kernel of libmvs_b200.so.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import Conv2dDesc, Conv3dDesc, call, dtype_code, ptr

Tensor = torch.Tensor


def _f32c(t: Tensor) -> Tensor:
    return t.detach().to(torch.float32).contiguous()


# ------------------------------------------------------------------------------------------------ layout
def pack_c8(x: Tensor, dtype: torch.dtype = torch.float32) -> Tensor:
    """[B,C,*spatial] -> C8 [B,C/8,*spatial,8] in `dtype` (from fp32 NCHW, or directly from a channels-last tensor of `dtype`)."""
    if x.dim() == 4 and x.dtype == dtype and x.shape[1] % 8 == 0 and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
        x = x.detach()
        b, c, h, w = x.shape
        out = torch.empty(b, c // 8, h, w, 8, dtype=dtype, device=x.device)
        call("mvs_nhwc_to_c8", x, ptr(x), ptr(out), b, c, h * w, dtype_code(dtype))
        return out
    x = _f32c(x)
    b, c = x.shape[0], x.shape[1]
    if c % 8:
        raise ValueError("C8 layout needs a channel count divisible by 8, got %d" % c)
    sp = tuple(x.shape[2:])
    s = 1
    for v in sp:
        s *= v
    out = torch.empty((b, c // 8) + sp + (8,), dtype=dtype, device=x.device)
    call("mvs_pack_c8", x, ptr(x), ptr(out), b, c, s, dtype_code(dtype))
    return out


def alloc_c8p(m: int, cb: int, h: int, w: int, dtype: torch.dtype, device, zero: bool) -> Tensor:
    """Storage of zero-bordered C8P maps [m, cb, h+3, w+2, 8] FOLLOWED BY ONE ZERO VECTOR (include/mvs_b200.h): a sample beyond
    the bottom-right corner puts its zero-weight 4th tap at row h+2, column w+2 -- the first vector of the next plane, i.e. one
    vector past the buffer for the last plane.  The returned tensor is a view of the first m*cb*(h+3)*(w+2)*8 elements."""
    n = m * cb * (h + 3) * (w + 2) * 8
    flat = torch.zeros(n + 8, dtype=dtype, device=device) if zero else torch.empty(n + 8, dtype=dtype, device=device)
    if not zero:
        flat[n:].zero_()
    return flat[:n].view(m, cb, h + 3, w + 2, 8)


def pack_c8_padded(x: Tensor, dtype: torch.dtype = torch.float32) -> Tensor:
    """[M,C,H,W] (fp32 NCHW, or channels-last `dtype`) or C8 [M,C/8,H,W,8] `dtype` -> zero-bordered C8P [M,C/8,H+3,W+2,8] `dtype`:
    pixel (y, x) sits at row y+1, column x+1 (the layout the fused warp+variance kernel gathers from)."""
    x = x.detach()
    if x.dim() == 5:
        if x.dtype != dtype:
            raise ValueError("a C8 input must already be in the target dtype")
        x = x.contiguous()
        m, cb, h, w, _ = x.shape
        c, layout = cb * 8, 2
    elif x.dim() == 4:
        m, c, h, w = x.shape
        if c % 8:
            raise ValueError("C8 layout needs a channel count divisible by 8, got %d" % c)
        if x.dtype == dtype and not x.is_contiguous() and x.is_contiguous(memory_format=torch.channels_last):
            layout = 1
        else:
            x, layout = _f32c(x), 0
    else:
        raise ValueError("expected [M,C,H,W] or C8 [M,C/8,H,W,8], got %s" % (tuple(x.shape),))
    out = alloc_c8p(m, c // 8, h, w, dtype, x.device, zero=False)
    call("mvs_pack_c8_padded", x, ptr(x), ptr(out), m, c, h, w, layout, dtype_code(dtype))
    return out


def unpack_c8(x: Tensor) -> Tensor:
    """C8 [B,C/8,*spatial,8] -> fp32 [B,C,*spatial]."""
    x = x.detach().contiguous()
    b, cb = x.shape[0], x.shape[1]
    sp = tuple(x.shape[2:-1])
    s = 1
    for v in sp:
        s *= v
    out = torch.empty((b, cb * 8) + sp, dtype=torch.float32, device=x.device)
    call("mvs_unpack_c8", x, ptr(x), ptr(out), b, cb * 8, s, dtype_code(x.dtype))
    return out


# ------------------------------------------------------------------------------------------------ projections
def compose_proj(proj_matrices: Tensor) -> Tensor:
    """[B,N,4,4] -> rt [N-1,B,12] (rot row-major | trans) of proj[:, s+1] @ inverse(proj[:, 0])."""
    p = _f32c(proj_matrices)
    b, n = p.shape[0], p.shape[1]
    rt = torch.empty(n - 1, b, 12, dtype=torch.float32, device=p.device)
    call("mvs_compose_proj", p, ptr(p), ptr(rt), b, n)
    return rt


def compose_proj_ke(ref_in: Tensor, src_in: Tensor, ref_ex: Tensor, src_ex: Tensor, down: float = 1.0) -> Tensor:
    """(K, E) pairs -> rt [nsrc,B,12]; src_in [B,nsrc,3,3], src_ex [B,nsrc,4,4]; K[:2] is divided by `down`."""
    ref_in, src_in, ref_ex, src_ex = _f32c(ref_in), _f32c(src_in), _f32c(ref_ex), _f32c(src_ex)
    b, nsrc = src_in.shape[0], src_in.shape[1]
    rt = torch.empty(nsrc, b, 12, dtype=torch.float32, device=ref_in.device)
    call("mvs_compose_proj_ke", ref_in, ptr(ref_in), ptr(src_in), ptr(ref_ex), ptr(src_ex), float(down), ptr(rt), b, nsrc)
    return rt


# ------------------------------------------------------------------------------------------------ stand-alone warp
class _HomoWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, src_fea: Tensor, rt: Tensor, depth: Tensor, align_corners: bool) -> Tensor:
        src = _f32c(src_fea)
        depth = _f32c(depth)
        b, c, h, w = src.shape
        d = depth.shape[1]
        per_pixel = int(depth.dim() == 4)
        out = torch.empty(b, c, d, h, w, dtype=torch.float32, device=src.device)
        call("mvs_homo_warp_fwd", src, ptr(src), ptr(rt), ptr(depth), per_pixel, ptr(out), b, c, d, h, w, int(align_corners))
        ctx.save_for_backward(rt, depth)
        ctx.meta = (b, c, d, h, w, per_pixel, int(align_corners))
        return out

    @staticmethod
    def backward(ctx, grad_out: Tensor):
        rt, depth = ctx.saved_tensors
        b, c, d, h, w, per_pixel, ac = ctx.meta
        g = _f32c(grad_out)
        gsrc = torch.zeros(b, c, h, w, dtype=torch.float32, device=g.device)
        call("mvs_homo_warp_bwd", g, ptr(g), ptr(rt), ptr(depth), per_pixel, ptr(gsrc), b, c, d, h, w, ac)
        return gsrc, None, None, None


def homo_warp(src_fea: Tensor, rt: Tensor, depth: Tensor, align_corners: bool = False) -> Tensor:
    """[B,C,H,W] x rt[B,12] x depth([B,D] | [B,D,H,W]) -> [B,C,D,H,W]; gradient reaches src_fea only (hazard H13)."""
    return _HomoWarp.apply(src_fea, rt.contiguous(), depth, align_corners)


# ------------------------------------------------------------------------------------------------ fused warp + variance
def _ptr_array(tensors: Sequence[Optional[Tensor]]):
    arr = (C.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = None if t is None else t.data_ptr()
    return arr


class _WarpVariance(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rt: Tensor, depth: Tensor, dtype: torch.dtype, align_corners: bool, ref_sq_in_sum: bool,
                ref_fea: Tensor, *src_feas: Tensor) -> Tensor:
        fdt = torch.float32 if dtype == torch.float32 else dtype
        ref8 = pack_c8_padded(ref_fea, fdt)
        src8 = [pack_c8_padded(s, fdt) for s in src_feas]
        depth = _f32c(depth)
        c8_in = ref_fea.dim() == 5                     # C8 maps [B, C/8, H, W, 8] (already in `fdt`) instead of [B, C, H, W]
        if c8_in:
            b, c, h, w = ref_fea.shape[0], ref_fea.shape[1] * 8, ref_fea.shape[2], ref_fea.shape[3]
        else:
            b, c, h, w = ref_fea.shape
        d = depth.shape[1]
        per_pixel = int(depth.dim() == 4)
        var = torch.empty(b, c // 8, d, h, w, 8, dtype=dtype, device=ref_fea.device)
        call("mvs_warp_var_fwd", ref8, ptr(ref8), _ptr_array(src8), len(src8), ptr(rt), ptr(depth), per_pixel, ptr(var),
             b, c, d, h, w, dtype_code(fdt), dtype_code(dtype), int(align_corners), int(ref_sq_in_sum), 1)
        ctx.save_for_backward(rt, depth, ref8, *src8)
        ctx.meta = (b, c, d, h, w, per_pixel, fdt, dtype, int(align_corners), int(ref_sq_in_sum), c8_in)
        return var

    @staticmethod
    def backward(ctx, grad_var: Tensor):
        rt, depth, ref8, *src8 = ctx.saved_tensors
        b, c, d, h, w, per_pixel, fdt, dtype, ac, rsq, c8_in = ctx.meta
        g = grad_var.detach().to(dtype).contiguous()
        need = ctx.needs_input_grad
        gref = torch.zeros(b, c // 8, h, w, 8, dtype=torch.float32, device=g.device) if need[5] else None
        gsrc = [torch.zeros(b, c // 8, h, w, 8, dtype=torch.float32, device=g.device) if need[6 + i] else None
                for i in range(len(src8))]
        call("mvs_warp_var_bwd", g, ptr(g), ptr(ref8), _ptr_array(src8), len(src8), ptr(rt), ptr(depth), per_pixel,
             ptr(gref), _ptr_array(gsrc), b, c, d, h, w, dtype_code(fdt), dtype_code(dtype), ac, rsq, 1)
        outs = [None if t is None else (t if c8_in else unpack_c8(t)) for t in [gref] + gsrc]
        return (None, None, None, None, None, *outs)


def warp_variance(ref_fea: Tensor, src_feas: Sequence[Tensor], rt: Tensor, depth: Tensor,
                  dtype: torch.dtype = torch.float32, align_corners: bool = False, ref_sq_in_sum: bool = False) -> Tensor:
    """Variance cost volume (C8, `dtype`) of the reference map and nsrc warped source maps ([B,C,H,W] each; fp32, or
    channels-last `dtype` as the library feature extractor emits them; or C8 maps [B,C/8,H,W,8] in `dtype`, whose gradients
    come back as fp32 C8 maps)."""
    if len(src_feas) < 1 or len(src_feas) > _lib.MAX_SRC:
        raise ValueError("need 1..%d source views, got %d" % (_lib.MAX_SRC, len(src_feas)))
    return _WarpVariance.apply(rt.contiguous(), depth, dtype, align_corners, ref_sq_in_sum, ref_fea, *src_feas)


def warp_variance_maps(maps: Tensor, rt: Tensor, depth: Tensor, dtype: torch.dtype, align_corners: bool = False,
                       ref_sq_in_sum: bool = False) -> Tensor:
    """Inference form (no autograd): `maps` = zero-bordered C8P maps of ALL views, [N,B,C/8,H+3,W+2,8] in `dtype`
    (pack_c8_padded of the stacked features), view 0 = reference.  Returns the C8 variance volume [B,C/8,D,H,W,8]."""
    n, b, cb, hp, wp, _ = maps.shape
    if n < 2 or n - 1 > _lib.MAX_SRC:
        raise ValueError("need 2..%d views, got %d" % (_lib.MAX_SRC + 1, n))
    if maps.dtype != dtype or not maps.is_contiguous():
        raise ValueError("maps must be contiguous and stored in the volume dtype")
    h, w = hp - 3, wp - 2
    depth = _f32c(depth)
    d = depth.shape[1]
    var = torch.empty(b, cb, d, h, w, 8, dtype=dtype, device=maps.device)
    call("mvs_warp_var_fwd", maps, ptr(maps[0]), _ptr_array([maps[v] for v in range(1, n)]), n - 1, ptr(rt.contiguous()), ptr(depth),
         int(depth.dim() == 4), ptr(var), b, cb * 8, d, h, w, dtype_code(dtype), dtype_code(dtype), int(align_corners),
         int(ref_sq_in_sum), 1)
    return var


# ------------------------------------------------------------------------------------------------ 3-D convolution
def _out_extent(n: int, stride: int, transposed: bool) -> int:
    if stride == 1:
        return n
    if transposed:
        return 2 * n
    if n % 2:
        raise ValueError("stride-2 Conv3d needs even extents, got %d (the reference U-Net has the same constraint)" % n)
    return n // 2


def pack_conv3d_weight(weight: Tensor, transposed: bool) -> Tensor:
    """torch Conv3d [Cout,Cin,3,3,3] / ConvTranspose3d [Cin,Cout,3,3,3] weight -> gather form G[27,Cin,CoutPad] fp32."""
    w = _f32c(weight)
    if tuple(w.shape[2:]) != (3, 3, 3):
        raise ValueError("only 3x3x3 kernels are on the path")
    cin, cout = (w.shape[0], w.shape[1]) if transposed else (w.shape[1], w.shape[0])
    g = torch.empty(27, cin, (cout + 7) // 8 * 8, dtype=torch.float32, device=w.device)
    call("mvs_pack_conv3d_weight", w, ptr(w), ptr(g), cin, cout, int(transposed))
    return g


def _desc(x: Tensor, cout: int, stride: int, transposed: bool, out_dtype: torch.dtype, relu: bool, algo: int) -> Conv3dDesc:
    b, cb, d, h, w, _ = x.shape
    return Conv3dDesc(b, cb * 8, cout, d, h, w, _out_extent(d, stride, transposed), _out_extent(h, stride, transposed),
                      _out_extent(w, stride, transposed), stride, int(transposed), dtype_code(x.dtype),
                      dtype_code(out_dtype), int(relu), algo)


def conv3d_raw(x: Tensor, g: Tensor, cout: int, stride: int = 1, transposed: bool = False,
               scale: Optional[Tensor] = None, shift: Optional[Tensor] = None, skip: Optional[Tensor] = None,
               relu: bool = False, out_dtype: Optional[torch.dtype] = None, algo: int = 0,
               tile_cache: Optional[dict] = None) -> Tensor:
    """y = [relu](conv(x) * scale + shift) + skip on C8 volumes (no autograd).  cout == 1 gives plain fp32 [B,D,H,W]."""
    x = x.contiguous()
    out_dtype = torch.float32 if cout == 1 else (out_dtype or x.dtype)
    d = _desc(x, cout, stride, transposed, out_dtype, relu, algo)
    if cout == 1:
        y = torch.empty(d.B, d.Dout, d.Hout, d.Wout, dtype=torch.float32, device=x.device)
    else:
        y = torch.empty(d.B, cout // 8, d.Dout, d.Hout, d.Wout, 8, dtype=out_dtype, device=x.device)
    if skip is not None:
        skip = skip.contiguous()
        if skip.shape != y.shape or skip.dtype != y.dtype:
            raise ValueError("skip tensor must match the output (%s %s vs %s %s)" % (tuple(skip.shape), skip.dtype, tuple(y.shape), y.dtype))
    nws = _lib.lib().mvs_conv3d_workspace_bytes(C.byref(d))
    ws = None
    if nws > 0:
        # tcgen05 path: weight tap tiles live in a workspace; with frozen weights (tile_cache) they are packed once
        key = (g.data_ptr(), g._version, d.dtype_in, stride, int(transposed), str(x.device))
        ws = tile_cache.get(key) if tile_cache is not None else None
        if ws is not None:
            d.algo = 3
        else:
            ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
            if tile_cache is not None:
                if len(tile_cache) > 256:
                    tile_cache.clear()
                tile_cache[key] = ws
    call("mvs_conv3d_fwd", x, C.byref(d), ptr(x), ptr(g), ptr(scale), ptr(shift), ptr(skip), ptr(y), ptr(ws))
    return y


# ------------------------------------------------------------------------------------------------ 2-D feature layers (tcgen05)
def pack_images_c8(imgs: Tensor, dtype: torch.dtype) -> Tensor:
    """images [B,N,3,H,W] (fp32 / fp16 / bf16) -> C8 image stack [1, N*B, H, W, 8] (`dtype`; image m = v*B + b; channels 3..7 zero)."""
    x = imgs.detach().contiguous() if imgs.dtype in (torch.float16, torch.bfloat16) else _f32c(imgs)
    b, n, c, h, w = x.shape
    if c != 3:
        raise ValueError("expected 3-channel images, got %d channels" % c)
    out = torch.empty(1, n * b, h, w, 8, dtype=dtype, device=x.device)
    call("mvs_pack_images_c8", x, ptr(x), dtype_code(x.dtype), ptr(out), b, n, h, w, dtype_code(dtype))
    return out


def pack_conv2d_weight(weight: Tensor) -> Tensor:
    """torch Conv2d weight [Cout,Cin,k,k] -> gather form G[k*k, CinPad, CoutPad] fp32 (tap = kh*k + kw; pads are zero)."""
    w = _f32c(weight)
    cout, cin, k, _ = w.shape
    cinp, coutp = (cin + 7) // 8 * 8, (cout + 7) // 8 * 8
    g = torch.zeros(k * k, cinp, coutp, dtype=torch.float32, device=w.device)
    g[:, :cin, :cout] = w.permute(2, 3, 1, 0).reshape(k * k, cin, cout)
    return g


def conv2d_raw(x: Tensor, g: Tensor, cout: int, ksize: int, stride: int, scale: Optional[Tensor], shift: Optional[Tensor],
               relu, out_padded: bool = False, tile_cache: Optional[dict] = None) -> Tensor:
    """y = act(conv2d(x) * scale + shift) over a C8 image stack [Cin/8, M, H, W, 8] (16-bit storage, tcgen05 kernel);
    relu: False / True, or a float slope for LeakyReLU.
    Returns the stack [Cout/8, M, Ho, Wo, 8], or with out_padded the zero-bordered image-major maps [M, Cout/8, Ho+3, Wo+2, 8]."""
    x = x.contiguous()
    cib, m, h, w, _ = x.shape
    ho, wo = (h, w) if stride == 1 else (h // 2, w // 2)
    leaky = isinstance(relu, float)
    d = Conv2dDesc(m, cib * 8, cout, h, w, ho, wo, ksize, stride, dtype_code(x.dtype), 2 if leaky else int(bool(relu)), int(out_padded), 0,
                   float(relu) if leaky else 0.0)
    if out_padded:
        y = alloc_c8p(m, cout // 8, ho, wo, x.dtype, x.device, zero=True)
    else:
        y = torch.empty(cout // 8, m, ho, wo, 8, dtype=x.dtype, device=x.device)
    nws = _lib.lib().mvs_conv2d_workspace_bytes(C.byref(d))
    if nws <= 0:
        raise RuntimeError("conv2d_raw: unsupported layer (Cin=%d Cout=%d k=%d stride=%d %s)" % (cib * 8, cout, ksize, stride, x.dtype))
    key = (g.data_ptr(), g._version, d.dtype, ksize, stride, str(x.device))
    ws = tile_cache.get(key) if tile_cache is not None else None
    if ws is not None:
        d.ws_packed = 1
    else:
        ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
        if tile_cache is not None:
            tile_cache[key] = ws
    call("mvs_conv2d_fwd", x, C.byref(d), ptr(x), ptr(g), ptr(scale), ptr(shift), ptr(y), ptr(ws))
    return y


def featnet_front_pack(w0: Tensor, w1: Tensor, w2: Tensor, affines: Sequence[Tuple[Tensor, Tensor]], dtype: torch.dtype):
    """Weights of FeatureNet.conv0/1/2 and their folded-BN (scale, shift) pairs -> (fragment buffer, affine table [3,32]) for
    featnet_front."""
    w0, w1, w2 = _f32c(w0), _f32c(w1), _f32c(w2)
    if tuple(w0.shape) != (8, 3, 3, 3) or tuple(w1.shape) != (8, 8, 3, 3) or tuple(w2.shape) != (16, 8, 5, 5):
        raise ValueError("featnet_front: expected FeatureNet's conv0 / conv1 / conv2 weights")
    frag = torch.empty(_lib.lib().mvs_featnet_front_workspace_bytes(), dtype=torch.uint8, device=w0.device)
    call("mvs_featnet_front_pack", w0, ptr(w0), ptr(w1), ptr(w2), ptr(frag), dtype_code(dtype))
    aff = torch.zeros(3, 32, dtype=torch.float32, device=w0.device)
    for i, (scale, shift) in enumerate(affines):
        aff[i, :scale.numel()] = scale
        aff[i, 16:16 + shift.numel()] = shift
    return frag, aff


def featnet_front(imgs: Tensor, frag: Tensor, aff: Tensor, dtype: torch.dtype) -> Tensor:
    """images [B,N,3,H,W] (fp32 or `dtype`) -> relu(bn(conv2(relu(bn(conv1(relu(bn(conv0(x))))))))) as the C8 stack
    [2, N*B, H/2, W/2, 8] in `dtype`: one launch, the 8-channel full-resolution maps never reach HBM."""
    x = imgs.detach().contiguous() if imgs.dtype == dtype else _f32c(imgs)
    b, n, c, h, w = x.shape
    if c != 3 or h % 2 or w % 2:
        raise ValueError("featnet_front: expected 3-channel images with even extents, got %s" % (tuple(x.shape),))
    out = torch.empty(2, n * b, h // 2, w // 2, 8, dtype=dtype, device=x.device)
    call("mvs_featnet_front", x, ptr(x), dtype_code(x.dtype), ptr(frag), ptr(aff), ptr(out), b, n, h, w, dtype_code(dtype))
    return out


def _pad_single_channel(t: Tensor) -> Tensor:
    """plain [B,D,H,W] -> C8 [B,1,D,H,W,8] with the value in channel 0."""
    out = torch.zeros(t.shape[0], 1, t.shape[1], t.shape[2], t.shape[3], 8, dtype=torch.float32, device=t.device)
    out[..., 0] = t.unsqueeze(1)
    return out


class _Conv3d(torch.autograd.Function):
    """Un-activated 3x3x3 (transposed) convolution with optional bias, fp32 C8, differentiable (training path)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], stride: int, transposed: bool) -> Tensor:
        cout = weight.shape[1] if transposed else weight.shape[0]
        g = pack_conv3d_weight(weight, transposed)
        y = conv3d_raw(x, g, cout, stride, transposed, shift=None if bias is None else _f32c(bias), algo=1)
        ctx.save_for_backward(x, weight)
        ctx.meta = (stride, transposed, cout, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        x, weight = ctx.saved_tensors
        stride, transposed, cout, has_bias = ctx.meta
        gy = _f32c(gy)
        gbias = None
        w_eff = _f32c(weight)
        if cout == 1:  # lift the single-channel gradient (and weight) to one C8 block
            if has_bias:
                gbias = gy.sum().reshape(1)
            gy = _pad_single_channel(gy)
            pad_shape = list(w_eff.shape)
            pad_shape[1 if transposed else 0] = 8
            w8 = torch.zeros(pad_shape, dtype=torch.float32, device=w_eff.device)
            if transposed:
                w8[:, :1] = w_eff
            else:
                w8[:1] = w_eff
            w_eff, cout_eff = w8, 8
        else:
            cout_eff = cout
            if has_bias:
                gbias = gy.sum(dim=(0, 2, 3, 4)).reshape(-1)  # [B,Cb,D,H,W,8] -> [Cb*8]
        gx = gw = None
        if ctx.needs_input_grad[0]:
            # adjoint: the same torch weight packed under the opposite flag (ATen's definition of conv_transpose)
            g_adj = pack_conv3d_weight(w_eff, not transposed)
            cin = x.shape[1] * 8
            if stride == 2 and not transposed:
                gx = conv3d_raw(gy, g_adj, cin, 2, True, algo=1)
            elif stride == 2 and transposed:
                gx = conv3d_raw(gy, g_adj, cin, 2, False, algo=1)
            else:
                gx = conv3d_raw(gy, g_adj, cin, 1, not transposed, algo=1)
        if ctx.needs_input_grad[1]:
            d = _desc(x, cout_eff, stride, transposed, torch.float32, False, 1)
            gw8 = torch.zeros_like(w_eff)
            xc = x.contiguous()
            call("mvs_conv3d_bwd_weight", xc, C.byref(d), ptr(xc), ptr(gy), ptr(gw8))
            if cout == 1:
                gw = gw8[:, :1].contiguous() if transposed else gw8[:1].contiguous()
            else:
                gw = gw8
        return gx, gw, gbias, None, None


def conv3d(x: Tensor, weight: Tensor, bias: Optional[Tensor] = None, stride: int = 1, transposed: bool = False) -> Tensor:
    return _Conv3d.apply(x, weight, bias, stride, transposed)


class _BnAct(torch.autograd.Function):
    """Training-mode BatchNorm3d (+ReLU) (+skip) on an fp32 C8 volume; returns (y, batch_mean, batch_var_biased)."""

    @staticmethod
    def forward(ctx, z: Tensor, gamma: Tensor, beta: Tensor, skip: Optional[Tensor], relu: bool, eps: float):
        z = z.contiguous()
        b, cb = z.shape[0], z.shape[1]
        c = cb * 8
        s = z[0, 0].numel() // 8
        sums = torch.zeros(2, c, dtype=torch.float32, device=z.device)
        call("mvs_bn_stats", z, ptr(z), ptr(sums), b, c, s)
        m = float(b * s)
        mean = sums[0] / m
        var = (sums[1] / m - mean * mean).clamp_min_(0.0)
        invstd = torch.rsqrt(var + eps)
        gamma_c, beta_c = _f32c(gamma), _f32c(beta)
        y = torch.empty_like(z)
        skip_c = None if skip is None else skip.contiguous()
        call("mvs_bn_act_fwd", z, ptr(z), ptr(mean), ptr(invstd), ptr(gamma_c), ptr(beta_c), ptr(skip_c), ptr(y), b, c, s, int(relu))
        ctx.save_for_backward(z, mean, invstd, gamma_c, beta_c)
        ctx.meta = (b, c, s, int(relu), skip is not None)
        ctx.mark_non_differentiable(mean, var)
        return y, mean, var

    @staticmethod
    def backward(ctx, gy: Tensor, _gm, _gv):
        z, mean, invstd, gamma, beta = ctx.saved_tensors
        b, c, s, relu, has_skip = ctx.meta
        gy = _f32c(gy)
        red = torch.zeros(2, c, dtype=torch.float32, device=z.device)
        call("mvs_bn_act_bwd_reduce", z, ptr(z), ptr(gy), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(red), b, c, s, relu)
        gz = torch.empty_like(z)
        call("mvs_bn_act_bwd_apply", z, ptr(z), ptr(gy), ptr(mean), ptr(invstd), ptr(gamma), ptr(beta), ptr(red), ptr(gz), b, c, s, relu)
        return gz, red[1].clone(), red[0].clone(), (gy if has_skip else None), None, None


def bn_act_train(z: Tensor, bn: torch.nn.modules.batchnorm._BatchNorm, skip: Optional[Tensor] = None, relu: bool = True) -> Tensor:
    """Batch-statistics BN + ReLU + skip; updates bn.running_* like nn.BatchNorm3d.train() (momentum, unbiased var)."""
    y, mean, var = _BnAct.apply(z, bn.weight, bn.bias, skip, relu, bn.eps)
    if bn.track_running_stats and bn.running_mean is not None:
        with torch.no_grad():
            n = z.shape[0] * (z[0, 0].numel() // 8)
            mom = bn.momentum if bn.momentum is not None else 0.1
            bn.running_mean.mul_(1 - mom).add_(mean.to(bn.running_mean.dtype), alpha=mom)
            bn.running_var.mul_(1 - mom).add_((var * (n / max(n - 1, 1))).to(bn.running_var.dtype), alpha=mom)
            if bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
    return y


# ------------------------------------------------------------------------------------------------ training on tensor cores (16-bit activations)
def _adjoint_conv(g: Tensor, weight: Tensor, cin: int, stride: int, transposed: bool) -> Tensor:
    """gradient w.r.t. the input of a (transposed) convolution = the convolution of `g` with the same torch weight packed under the
    opposite flag (ATen's definition of conv_transpose), on the tcgen05 kernel: stride-2 Conv3d <-> stride-2 ConvTranspose3d."""
    g_adj = pack_conv3d_weight(weight, not transposed)
    return conv3d_raw(g, g_adj, cin, stride, not transposed, algo=0)


def _wgrad_mma(x: Tensor, gz: Tensor, weight: Tensor, cout: int, stride: int, transposed: bool, cout_real: int, planar: bool = False) -> Tensor:
    """planar: a 2-D layer run as a zero-kd 3-D layer over an image volume -- only the kd = 1 taps are computed."""
    d = _desc(x, cout, stride, transposed, x.dtype, False, 0)
    gw = torch.zeros_like(weight, dtype=torch.float32)
    call("mvs_conv2d_wgrad_mma" if planar else "mvs_conv3d_wgrad_mma", x, C.byref(d), ptr(x), ptr(gz), ptr(gw), cout_real)
    return gw


class _ConvBnActTC(torch.autograd.Function):
    """y = relu(bn(conv(x))) + skip for C8 volumes stored in fp16 / bf16: convolution and input gradient on the tcgen05 kernel,
    weight gradient on warp-level tensor-core MMAs, BatchNorm statistics in fp64 (csrc/train.cu).  `frozen` uses the running
    statistics (fine-tuning under module.eval()), otherwise batch statistics are taken and the running ones updated."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, gamma: Tensor, beta: Tensor, skip: Optional[Tensor], running_mean: Optional[Tensor],
                running_var: Optional[Tensor], stride: int, transposed: bool, eps: float, momentum: float, frozen: bool,
                planar: bool = False, per_item_stats: bool = False) -> Tensor:
        """per_item_stats: every batch entry of x normalises with its OWN batch statistics (the feature extractor in training:
        one entry = the images of one view, which the reference runs as separate calls).  The layout [V][C/8][S][8] of V entries
        is that of ONE entry with V C channels, so the BatchNorm kernels simply see B = 1, C' = V C; the running statistics take
        the V momentum updates the separate calls would have made, in order."""
        x = x.contiguous()
        dt = x.dtype
        cout = weight.shape[1] if transposed else weight.shape[0]
        z = conv3d_raw(x, pack_conv3d_weight(weight, transposed), cout, stride, transposed, algo=0)
        b, cb = z.shape[0], z.shape[1]
        s = z[0, 0].numel() // 8
        dev = z.device
        gamma_c, beta_c = _f32c(gamma), _f32c(beta)
        groups = b if (per_item_stats and b > 1) else 1
        if groups > 1:
            b, ceff = 1, groups * cout
            gamma_c, beta_c = gamma_c.repeat(groups), beta_c.repeat(groups)
        else:
            ceff = cout
        stats = torch.empty(4, ceff, dtype=torch.float32, device=dev)       # a, b, mean, invstd
        if frozen:
            inv = torch.rsqrt(running_var.detach().float() + eps).repeat(groups)
            stats[0] = gamma_c * inv
            stats[1] = beta_c - running_mean.detach().float().repeat(groups) * stats[0]
            stats[2] = running_mean.detach().float().repeat(groups)
            stats[3] = inv
        else:
            sums = torch.zeros(2, ceff, dtype=torch.float64, device=dev)
            call("mvs_bn_stats_t", z, ptr(z), dtype_code(dt), ptr(sums), b, ceff, s)
            direct = groups == 1
            call("mvs_bn_finalize", z, ptr(sums), ptr(gamma_c), ptr(beta_c), float(eps), float(momentum), float(b * s), ptr(stats[0]), ptr(stats[1]),
                 ptr(stats[2]), ptr(stats[3]), ptr(running_mean) if direct else None, ptr(running_var) if direct else None, ceff)
            if not direct and running_mean is not None:
                # V sequential updates r <- (1 - m) r + m x_v  ==  (1 - m)^V r + sum_v m (1 - m)^(V - 1 - v) x_v
                # (built on the device: no host-to-device copy, the step may be under CUDA-graph capture)
                wts = (momentum * torch.pow(1.0 - momentum, torch.arange(groups - 1, -1, -1, dtype=torch.float32, device=dev))).view(groups, 1)
                n = float(s)
                var_unbiased = (stats[3].pow(-2) - eps).clamp_min_(0.0) * (n / max(n - 1.0, 1.0))
                running_mean.mul_((1.0 - momentum) ** groups).add_((wts * stats[2].view(groups, cout)).sum(0))
                running_var.mul_((1.0 - momentum) ** groups).add_((wts * var_unbiased.view(groups, cout)).sum(0))
        y = torch.empty_like(z)
        skip_c = None if skip is None else skip.contiguous()
        call("mvs_bn_act_fwd_t", z, ptr(z), ptr(stats[0]), ptr(stats[1]), ptr(skip_c), ptr(y), dtype_code(dt), b, ceff, s, 1)
        ctx.save_for_backward(x, weight, z, stats)
        ctx.meta = (stride, transposed, cout, skip is not None, frozen, planar, groups)
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        x, weight, z, stats = ctx.saved_tensors
        stride, transposed, cout, has_skip, frozen, planar, groups = ctx.meta
        dt = z.dtype
        gy = gy.detach().to(dt).contiguous()
        b = z.shape[0] if groups == 1 else 1
        ceff = cout * groups
        s = z[0, 0].numel() // 8
        red = torch.zeros(2, ceff, dtype=torch.float64, device=z.device)
        call("mvs_bn_act_bwd_reduce_t", z, ptr(z), ptr(gy), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]), ptr(stats[3]), ptr(red), dtype_code(dt),
             b, ceff, s, 1)
        gz = torch.empty_like(z)
        gpar = torch.empty(2, ceff, dtype=torch.float32, device=z.device)
        call("mvs_bn_act_bwd_apply_t", z, ptr(z), ptr(gy), ptr(stats[0]), ptr(stats[1]), ptr(stats[2]), ptr(stats[3]), ptr(red), ptr(gz),
             ptr(gpar[0]), ptr(gpar[1]), dtype_code(dt), b, ceff, s, 1, int(frozen))
        if groups > 1:
            gpar = gpar.view(2, groups, cout).sum(1)
        w32 = _f32c(weight)
        gx = _adjoint_conv(gz, w32, x.shape[1] * 8, stride, transposed) if ctx.needs_input_grad[0] else None
        gw = _wgrad_mma(x, gz, w32, cout, stride, transposed, cout, planar) if ctx.needs_input_grad[1] else None
        return (gx, gw, gpar[0] if ctx.needs_input_grad[2] else None, gpar[1] if ctx.needs_input_grad[3] else None,
                gy if has_skip else None, None, None, None, None, None, None, None, None, None)


def conv_bn_act_tc(x: Tensor, conv: torch.nn.Module, bn: torch.nn.modules.batchnorm._BatchNorm, skip: Optional[Tensor], frozen: bool) -> Tensor:
    transposed = isinstance(conv, torch.nn.ConvTranspose3d)
    mom = bn.momentum if bn.momentum is not None else 0.1
    track = bn.track_running_stats and bn.running_mean is not None
    y = _ConvBnActTC.apply(x, conv.weight, bn.weight, bn.bias, skip, bn.running_mean if track else None, bn.running_var if track else None,
                           conv.stride[0], transposed, bn.eps, mom, frozen)
    if track and not frozen and bn.num_batches_tracked is not None:
        with torch.no_grad():
            bn.num_batches_tracked += 1
    return y


class _ConvBiasTC(torch.autograd.Function):
    """The single-channel `prob` convolution (stride 1, bias): 16-bit C8 volume -> plain fp32 [B,D,H,W], differentiable."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tensor:
        x = x.contiguous()
        y = conv3d_raw(x, pack_conv3d_weight(weight, False), 1, 1, False, shift=None if bias is None else _f32c(bias), algo=0)
        ctx.save_for_backward(x, weight)
        ctx.has_bias = bias is not None
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        x, weight = ctx.saved_tensors
        gy = _f32c(gy)
        b, d, h, w = gy.shape
        g8 = torch.empty(b, 1, d, h, w, 8, dtype=x.dtype, device=x.device)     # the gradient lifted to one C8 block (channel 0)
        call("mvs_lift_c1", gy, ptr(gy), ptr(g8), dtype_code(x.dtype), gy.numel())
        w8 = torch.zeros(8, weight.shape[1], 3, 3, 3, dtype=torch.float32, device=weight.device)
        w8[:1] = weight.detach().float()
        gx = _adjoint_conv(g8, w8, x.shape[1] * 8, 1, False) if ctx.needs_input_grad[0] else None
        gw = None
        if ctx.needs_input_grad[1]:
            d8 = _desc(x, 8, 1, False, x.dtype, False, 0)
            gw = torch.zeros_like(weight, dtype=torch.float32)
            call("mvs_conv3d_wgrad_mma", x, C.byref(d8), ptr(x), ptr(g8), ptr(gw), 1)
        gb = gy.sum().reshape(1) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb


def conv_bias_tc(x: Tensor, conv: torch.nn.Conv3d) -> Tensor:
    return _ConvBiasTC.apply(x, conv.weight, conv.bias)


class _ConvTC(torch.autograd.Function):
    """z = conv(x) [+ bias] on a 16-bit C8 volume, stride 1, 3x3x3, differentiable: forward and input gradient on the tcgen05
    kernel, weight gradient on the warp-level MMA kernel (the un-normalised last layer of the 2-D feature extractor)."""

    @staticmethod
    def forward(ctx, x: Tensor, weight: Tensor, bias: Optional[Tensor], planar: bool = False) -> Tensor:
        x = x.contiguous()
        cout = weight.shape[0]
        y = conv3d_raw(x, pack_conv3d_weight(weight, False), cout, 1, False, shift=None if bias is None else _f32c(bias), algo=0)
        ctx.save_for_backward(x, weight)
        ctx.has_bias, ctx.planar = bias is not None, planar
        return y

    @staticmethod
    def backward(ctx, gy: Tensor):
        x, weight = ctx.saved_tensors
        cout = weight.shape[0]
        g16 = gy.detach().to(x.dtype).contiguous()
        w32 = _f32c(weight)
        gx = _adjoint_conv(g16, w32, x.shape[1] * 8, 1, False) if ctx.needs_input_grad[0] else None
        gw = _wgrad_mma(x, g16, w32, cout, 1, False, cout, ctx.planar) if ctx.needs_input_grad[1] else None
        gb = gy.detach().float().sum(dim=(0, 2, 3, 4)).reshape(-1) if (ctx.has_bias and ctx.needs_input_grad[2]) else None
        return gx, gw, gb, None


# 2-D layers of the feature extractor in TRAINING: a batch of images is a C8 volume whose depth axis is the image index
# ([1, C/8, M, H, W, 8]); a k x k Conv2d is the 3x3x3 convolution whose kd = 0, 2 taps are zero, so that every kernel of the
# 3-D training path (tcgen05 forward / input gradient, MMA weight gradient, fp64 BatchNorm statistics) serves unchanged and the
# images never mix.  5x5 stride-2 layers run as 3x3 stride-1 layers over the space-to-depth (2x2 parity) form of their input.
def space_to_depth_c8(x: Tensor) -> Tensor:
    """[V, Cb, M, H, W, 8] -> [V, 4 Cb, M, H/2, W/2, 8]; new channel block = (row parity * 2 + column parity) * Cb + block."""
    v, cb, m, h, w, _ = x.shape
    if h % 2 or w % 2:
        raise ValueError("stride-2 feature layers need even extents, got %dx%d" % (h, w))
    return x.view(v, cb, m, h // 2, 2, w // 2, 2, 8).permute(0, 4, 6, 1, 2, 3, 5, 7).reshape(v, 4 * cb, m, h // 2, w // 2, 8)


def embed_conv2d_weight(weight: Tensor, stride: int) -> Tensor:
    """torch Conv2d weight -> the [Cout, Cin', 3, 3, 3] weight of the equivalent stride-1 3-D layer (differentiable):
    3x3 stride 1 pad 1: Cin' = Cin padded to 8, taps at kd = 1;  5x5 stride 2 pad 2: Cin' = 4 Cin (parity-major, as
    space_to_depth_c8 orders them), parity (a, b) holds taps W[.., a::2, b::2] (3 or 2 per axis, zero-padded to 3)."""
    cout, cin, k, _ = weight.shape
    F = torch.nn.functional
    if k == 3 and stride == 1:
        w2 = F.pad(weight, (0, 0, 0, 0, 0, (-cin) % 8))
    elif k == 5 and stride == 2:
        if cin % 8:
            raise ValueError("5x5 stride-2 layers need Cin divisible by 8, got %d" % cin)
        parts = []
        for a in (0, 1):
            for b in (0, 1):
                sub = weight[:, :, a::2, b::2]
                parts.append(F.pad(sub, (0, 3 - sub.shape[3], 0, 3 - sub.shape[2])))
        w2 = torch.cat(parts, dim=1)
    else:
        raise ValueError("feature layers on the path are 3x3 stride 1 or 5x5 stride 2 (got %dx%d stride %d)" % (k, k, stride))
    return F.pad(w2.unsqueeze(2), (0, 0, 0, 0, 1, 1))


def conv2d_bn_relu_tc(x: Tensor, conv: torch.nn.Conv2d, bn: torch.nn.modules.batchnorm._BatchNorm, frozen: bool) -> Tensor:
    """relu(bn(conv2d(x))) over image volumes [V, Cin/8, M, H, W, 8] (16-bit), differentiable; unless `frozen`, each of the V
    entries (one view's M images) normalises with its own batch statistics, as V separate calls of the layer would."""
    stride = conv.stride[0]
    if stride == 2:
        x = space_to_depth_c8(x)
    w3 = embed_conv2d_weight(conv.weight, stride)
    mom = bn.momentum if bn.momentum is not None else 0.1
    track = bn.track_running_stats and bn.running_mean is not None
    y = _ConvBnActTC.apply(x, w3, bn.weight, bn.bias, None, bn.running_mean if track else None, bn.running_var if track else None,
                           1, False, bn.eps, mom, frozen, True, True)
    if track and not frozen and bn.num_batches_tracked is not None:
        with torch.no_grad():
            bn.num_batches_tracked += x.shape[0]
    return y


def conv2d_bias_tc(x: Tensor, conv: torch.nn.Conv2d) -> Tensor:
    """conv2d(x) + bias over an image volume (3x3, stride 1), differentiable."""
    return _ConvTC.apply(x, embed_conv2d_weight(conv.weight, conv.stride[0]), conv.bias, True)


def fold_bn(bn: torch.nn.modules.batchnorm._BatchNorm) -> Tuple[Tensor, Tensor]:
    """Eval-mode BN as a per-channel affine: scale = gamma / sqrt(running_var + eps), shift = beta - mean * scale."""
    scale = (bn.weight.detach().float() * torch.rsqrt(bn.running_var.detach().float() + bn.eps)).contiguous()
    shift = (bn.bias.detach().float() - bn.running_mean.detach().float() * scale).contiguous()
    return scale, shift


# ------------------------------------------------------------------------------------------------ soft-argmin
class _SoftArgmin(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cost: Tensor, depth: Tensor, want_prob: bool):
        cost, depth = _f32c(cost), _f32c(depth)
        b, d, h, w = cost.shape
        per_pixel = int(depth.dim() == 4)
        dev = cost.device
        out = torch.empty(b, h, w, dtype=torch.float32, device=dev)
        index = torch.empty(b, h, w, dtype=torch.int64, device=dev)
        conf = torch.empty(b, h, w, dtype=torch.float32, device=dev)
        prob = torch.empty(b, d, h, w, dtype=torch.float32, device=dev) if want_prob else None
        call("mvs_softargmin_fwd", cost, ptr(cost), ptr(depth), per_pixel, ptr(out), ptr(index), ptr(conf), ptr(prob), b, d, h, w)
        ctx.save_for_backward(cost, depth)
        ctx.meta = (b, d, h, w, per_pixel)
        ctx.mark_non_differentiable(index, conf)
        if prob is None:
            prob = torch.empty(0, device=dev)
        ctx.mark_non_differentiable(prob)
        return out, index, conf, prob

    @staticmethod
    def backward(ctx, g_depth: Tensor, _gi, _gc, _gp):
        cost, depth = ctx.saved_tensors
        b, d, h, w, per_pixel = ctx.meta
        g = _f32c(g_depth)
        gcost = torch.empty_like(cost)
        call("mvs_softargmin_bwd", cost, ptr(cost), ptr(depth), per_pixel, ptr(g), ptr(gcost), b, d, h, w)
        return gcost, None, None


def soft_argmin(cost_reg: Tensor, depth: Tensor, want_prob: bool = False):
    """cost_reg [B,D,H,W] -> (depth [B,H,W], index int64, confidence, prob or None); gradient reaches cost_reg only."""
    out, index, conf, prob = _SoftArgmin.apply(cost_reg, depth, want_prob)
    return out, index, conf, (prob if want_prob else None)


# ------------------------------------------------------------------------------------------------ CVP hypotheses
def depth_hypo_refine(depth_up: Tensor, ref_in: Tensor, src_in0: Tensor, ref_ex: Tensor, src_ex0: Tensor, half: int = 4) -> Tensor:
    """[B,H,W] -> [B,2*half,H,W] = depth_up + k * mean|delta_d| (no gradient; the reference builds it under no_grad)."""
    depth_up = _f32c(depth_up)
    b, h, w = depth_up.shape
    ref_in, src_in0, ref_ex, src_ex0 = _f32c(ref_in), _f32c(src_in0), _f32c(ref_ex), _f32c(src_ex0)
    hyp = torch.empty(b, 2 * half, h, w, dtype=torch.float32, device=depth_up.device)
    ws = torch.zeros(b, dtype=torch.float64, device=depth_up.device)
    call("mvs_depth_hypo_refine", depth_up, ptr(depth_up), ptr(ref_in), ptr(src_in0), ptr(ref_ex), ptr(src_ex0), ptr(hyp), ptr(ws), b, h, w, half)
    return hyp


# ------------------------------------------------------------------------------------------------ loss-side inverse warp
class _InvWarp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, img: Tensor, left_cam: Tensor, right_cam: Tensor, depth: Tensor):
        img_c, depth_c = _f32c(img), _f32c(depth)
        left, right = _f32c(left_cam), _f32c(right_cam)
        b, h, w, c = img_c.shape
        warped = torch.empty_like(img_c)
        mask = torch.empty(b, h, w, 1, dtype=torch.float32, device=img_c.device)
        ws = torch.empty(b, 24, dtype=torch.float32, device=img_c.device)
        call("mvs_invwarp_fwd", img_c, ptr(img_c), ptr(left), ptr(right), ptr(depth_c), ptr(warped), ptr(mask), ptr(ws), b, h, w, c)
        ctx.save_for_backward(img_c, left, right, depth_c)
        ctx.mark_non_differentiable(mask)
        return warped, mask

    @staticmethod
    def backward(ctx, g_warped: Tensor, _gmask):
        img, left, right, depth = ctx.saved_tensors
        b, h, w, c = img.shape
        g = _f32c(g_warped)
        gdepth = torch.empty_like(depth)
        gimg = torch.zeros_like(img) if ctx.needs_input_grad[0] else None
        ws = torch.empty(b, 24, dtype=torch.float32, device=img.device)
        call("mvs_invwarp_bwd", img, ptr(img), ptr(left), ptr(right), ptr(depth), ptr(g), ptr(gdepth), ptr(gimg), ptr(ws), b, h, w, c)
        return gimg, None, None, gdepth


def inverse_warp(img: Tensor, left_cam: Tensor, right_cam: Tensor, depth: Tensor) -> Tuple[Tensor, Tensor]:
    """img [B,H,W,C], cams [B,2,4,4], depth [B,H,W] -> (warped [B,H,W,C], mask [B,H,W,1])."""
    return _InvWarp.apply(img, left_cam, right_cam, depth)


# ------------------------------------------------------------------------------------------------ fused self-supervised loss
class _UnsupLoss(torch.autograd.Function):
    """mvs_unsup_loss_fwd / _bwd: out [4] = (total, reconstr, ssim, smooth); the gradient reaches depth only."""

    @staticmethod
    def forward(ctx, depth: Tensor, imgs: Tensor, cams: Tensor, smooth_lambda: float, smooth_weight: float):
        imgs_c, cams_c, depth_c = _f32c(imgs), _f32c(cams), _f32c(depth)
        b, n, ch, hi, wi = imgs_c.shape
        if ch != 3:
            raise ValueError("UnSupLoss expects 3-channel views, got %d channels" % ch)
        h, w = depth_c.shape[1], depth_c.shape[2]
        dev, f32 = depth_c.device, torch.float32
        small = torch.empty(n, b, h, w, 3, dtype=f32, device=dev)
        warped = torch.empty(n - 1, b, h, w, 3, dtype=f32, device=dev)
        mask = torch.empty(n - 1, b, h, w, dtype=f32, device=dev)
        coef = torch.empty(2, b, h, w, 9, dtype=f32, device=dev)
        cam_ws = torch.empty(max(n - 1, 1) * b, 24, dtype=f32, device=dev)
        acc = torch.empty(48, dtype=torch.float64, device=dev)
        out = torch.empty(4, dtype=f32, device=dev)
        call("mvs_unsup_loss_fwd", depth_c, ptr(imgs_c), ptr(cams_c), ptr(depth_c), b, n, hi, wi, h, w, float(smooth_lambda),
             float(smooth_weight), ptr(small), ptr(warped), ptr(mask), ptr(coef), ptr(cam_ws), ptr(acc), ptr(out))
        ctx.save_for_backward(small, depth_c, warped, mask, coef, cam_ws, acc)
        ctx.cfg = (b, n, h, w, float(smooth_lambda), float(smooth_weight))
        return out

    @staticmethod
    def backward(ctx, g_out: Tensor):
        small, depth, warped, mask, coef, cam_ws, acc = ctx.saved_tensors
        b, n, h, w, lam, wsm = ctx.cfg
        g = _f32c(g_out)
        gdepth = torch.empty_like(depth)
        call("mvs_unsup_loss_bwd", depth, ptr(g), ptr(small), ptr(depth), ptr(warped), ptr(mask), ptr(coef), ptr(cam_ws), ptr(acc),
             ptr(gdepth), b, n, h, w, lam, wsm)
        return gdepth, None, None, None, None


def unsup_loss(imgs: Tensor, cams: Tensor, depth: Tensor, smooth_lambda: float = 1.0, smooth_weight: float = 0.18) -> Tensor:
    """imgs [B,N,3,Hi,Wi] (at depth size or 4x it), cams [B,N,2,4,4], depth [B,H,W] -> [4] = (12 rec + 6 ssim + w smooth, rec, ssim,
    smooth).  One C-ABI call forward, one backward; differentiable w.r.t. depth (the views are data)."""
    if imgs.requires_grad:
        raise NotImplementedError("the fused UnSupLoss sends no gradient to the images (they are data in the reference's training)")
    return _UnsupLoss.apply(depth, imgs, cams, smooth_lambda, smooth_weight)


# ------------------------------------------------------------------------------------------------ inference output side
def upsample_nearest(maps: Tensor, size: Tuple[int, int], flip_rows: bool = False) -> Tensor:
    """[M,H,W] fp32 -> [M,Ho,Wo]: F.interpolate(maps.unsqueeze(1), size=size) in the default nearest mode; flip_rows stores the
    rows bottom-up (the body of a .pfm file)."""
    x = _f32c(maps)
    m, h, w = x.shape
    ho, wo = int(size[0]), int(size[1])
    out = torch.empty(m, ho, wo, dtype=torch.float32, device=x.device)
    call("mvs_upsample_nearest", x, ptr(x), ptr(out), m, h, w, ho, wo, int(flip_rows))
    return out


def depth_preview_u8(depth: Tensor, offset: float = 500.0, scale: float = 2.0) -> Tensor:
    """write_depth_img's 8-bit preview: clamp((depth - offset) / scale, 0, 255), truncated."""
    x = _f32c(depth)
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device)
    call("mvs_depth_preview_u8", x, ptr(x), ptr(out), x.numel(), float(offset), float(scale))
    return out


def geo_consistency(depth_ref: Tensor, depth_src: Tensor, cams: Tensor, dist_thresh: float = 1.0, rel_thresh: float = 0.01,
                    apply_mask: bool = True):
    """depth_ref, depth_src [B,H,W] fp32; cams [B,60] float64 (see include/mvs_b200.h) ->
    (mask bool [B,H,W], depth_reprojected, x_src, y_src, x_reprojected, y_reprojected) as check_geometric_consistency /
    reproject_with_depth return them."""
    dr, ds = _f32c(depth_ref), _f32c(depth_src)
    cams = cams.detach().to(torch.float64).contiguous()
    b, h, w = dr.shape
    if ds.shape != dr.shape or cams.shape != (b, 60):
        raise ValueError("geo_consistency: depth maps must share [B,H,W] and cams be [B,60] (got %s, %s, %s)" % (tuple(dr.shape), tuple(ds.shape), tuple(cams.shape)))
    mask = torch.empty(b, h, w, dtype=torch.uint8, device=dr.device)
    outs = [torch.empty(b, h, w, dtype=torch.float32, device=dr.device) for _ in range(5)]
    call("mvs_geo_consistency", dr, ptr(dr), ptr(ds), ptr(cams), ptr(mask), *[ptr(o) for o in outs], b, h, w, float(dist_thresh),
         float(rel_thresh), int(apply_mask))
    return (mask.bool(), *outs)


# ------------------------------------------------------------------------------------------------ depth-map fusion
def fusibile(normals_depths: Tensor, cams: Tensor, ref: int, subset: Sequence[int], depth_thresh: float = 0.25,
             normal_thresh: float = 0.52, num_consistent: int = 3, images: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
    """normals_depths [V,H,W,4] (nx, ny, nz, depth), cams [V,32] (include/mvs_b200.h), optional images [V,H,W,4] ->
    (points [H,W,12] = coordinate | normal | colour, valid bool [H,W]) for reference view `ref` against the views in `subset`."""
    nd = _f32c(normals_depths)
    cams = _f32c(cams)
    v, h, w, c = nd.shape
    if c != 4 or cams.shape != (v, 32):
        raise ValueError("fusibile: normals_depths must be [V,H,W,4] and cams [V,32]")
    img = None if images is None else _f32c(images)
    sub = torch.tensor(list(subset), dtype=torch.int32, device=nd.device)
    points = torch.empty(h, w, 12, dtype=torch.float32, device=nd.device)
    valid = torch.empty(h, w, dtype=torch.uint8, device=nd.device)
    call("mvs_fusibile", nd, ptr(nd), ptr(img), ptr(cams), ptr(sub), sub.numel(), v, h, w, int(ref), float(depth_thresh), float(normal_thresh),
         int(num_consistent), ptr(points), ptr(valid))
    return points, valid.bool()
