"""ctypes binding of libmvs_b200.so (include/mvs_b200.h).

There is NO fallback: if the shared library is missing, or a tensor is not on a CUDA device, every op raises.
The only other library that can be bound is the host-emulation build of the same kernel sources, and only by
an explicit `bind(path)` call from the CPU unit tests (tests/emu); it reports mvs_is_emulation() == 1 and is
refused for CUDA tensors.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
DEFAULT_PATH = os.path.join(_HERE, "libmvs_b200.so")

F32, F16, BF16 = 0, 1, 2
_DT = {torch.float32: F32, torch.float16: F16, torch.bfloat16: BF16}
MAX_SRC = 8


class Conv3dDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("B", "Cin", "Cout", "Din", "Hin", "Win", "Dout", "Hout", "Wout", "stride",
                                       "transposed", "dtype_in", "dtype_out", "relu", "algo")]


class Conv2dDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("M", "Cin", "Cout", "Hin", "Win", "Hout", "Wout", "ksize", "stride", "dtype", "relu",
                                       "out_padded", "ws_packed")] + [("leaky_slope", C.c_float)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
_SIGS = {
    "mvs_version": ([], _I),
    "mvs_last_error": ([], C.c_char_p),
    "mvs_is_emulation": ([], _I),
    "mvs_set_knob": ([C.c_char_p, _I], _I),
    "mvs_pack_c8": ([_P, _P, _I, _I, _L, _I, _P], _I),
    "mvs_unpack_c8": ([_P, _P, _I, _I, _L, _I, _P], _I),
    "mvs_nhwc_to_c8": ([_P, _P, _I, _I, _L, _I, _P], _I),
    "mvs_compose_proj": ([_P, _P, _I, _I, _P], _I),
    "mvs_compose_proj_ke": ([_P, _P, _P, _P, _F, _P, _I, _I, _P], _I),
    "mvs_homo_warp_fwd": ([_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "mvs_homo_warp_bwd": ([_P, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "mvs_pack_c8_padded": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "mvs_warp_var_fwd": ([_P, C.POINTER(_P), _I, _P, _P, _I, _P, _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "mvs_warp_var_bwd": ([_P, _P, C.POINTER(_P), _I, _P, _P, _I, _P, C.POINTER(_P), _I, _I, _I, _I, _I, _I, _I, _I, _I, _I, _P], _I),
    "mvs_pack_conv3d_weight": ([_P, _P, _I, _I, _I, _P], _I),
    "mvs_conv3d_fwd": ([C.POINTER(Conv3dDesc), _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "mvs_conv3d_workspace_bytes": ([C.POINTER(Conv3dDesc)], _L),
    "mvs_conv3d_bwd_weight": ([C.POINTER(Conv3dDesc), _P, _P, _P, _P], _I),
    "mvs_conv2d_workspace_bytes": ([C.POINTER(Conv2dDesc)], _L),
    "mvs_conv2d_fwd": ([C.POINTER(Conv2dDesc), _P, _P, _P, _P, _P, _P, _P], _I),
    "mvs_pack_images_c8": ([_P, _I, _P, _I, _I, _I, _I, _I, _P], _I),
    "mvs_bn_stats": ([_P, _P, _I, _I, _L, _P], _I),
    "mvs_bn_act_fwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _P], _I),
    "mvs_bn_act_bwd_reduce": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _P], _I),
    "mvs_bn_act_bwd_apply": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _L, _I, _P], _I),
    "mvs_bn_stats_t": ([_P, _I, _P, _I, _I, _L, _P], _I),
    "mvs_bn_finalize": ([_P, _P, _P, _F, _F, C.c_double, _P, _P, _P, _P, _P, _P, _I, _P], _I),
    "mvs_bn_act_fwd_t": ([_P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _P], _I),
    "mvs_bn_act_bwd_reduce_t": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _P], _I),
    "mvs_bn_act_bwd_apply_t": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _L, _I, _I, _P], _I),
    "mvs_lift_c1": ([_P, _P, _I, _L, _P], _I),
    "mvs_conv3d_wgrad_mma": ([C.POINTER(Conv3dDesc), _P, _P, _P, _I, _P], _I),
    "mvs_conv2d_wgrad_mma": ([C.POINTER(Conv3dDesc), _P, _P, _P, _I, _P], _I),
    "mvs_softargmin_fwd": ([_P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "mvs_softargmin_bwd": ([_P, _P, _I, _P, _P, _I, _I, _I, _I, _P], _I),
    "mvs_depth_hypo_refine": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "mvs_invwarp_fwd": ([_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "mvs_invwarp_bwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _P], _I),
    "mvs_unsup_loss_fwd": ([_P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _F, _P, _P, _P, _P, _P, _P, _P, _P], _I),
    "mvs_featnet_front_workspace_bytes": ([], _L),
    "mvs_featnet_front_pack": ([_P, _P, _P, _P, _I, _P], _I),
    "mvs_featnet_front": ([_P, _I, _P, _P, _P, _I, _I, _I, _I, _I, _P], _I),
    "mvs_fusibile": ([_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _F, _I, _P, _P, _P], _I),
    "mvs_upsample_nearest": ([_P, _P, _I, _I, _I, _I, _I, _I, _P], _I),
    "mvs_depth_preview_u8": ([_P, _P, _L, _F, _F, _P], _I),
    "mvs_geo_consistency": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _F, _I, _P], _I),
    "mvs_unsup_loss_bwd": ([_P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _F, _F, _P], _I),
}
EXPORTS = tuple(_SIGS)

_lib: Optional[C.CDLL] = None
_emulation = False
launches = 0  # kernel-launching C-ABI calls made through this module (bench.py reports it)


def bind(path: Optional[str] = None) -> C.CDLL:
    """Load the shared library at `path` (default, and the only thing the package itself ever loads: the in-tree build)
    and type every export.  Raises if anything is missing."""
    global _lib, _emulation
    if path is None:
        path = DEFAULT_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            "libmvs_b200.so not found at %s: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU or PyTorch fallback for the plane-sweep path." % path)
    lib = C.CDLL(path)
    for name, (args, res) in _SIGS.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.argtypes, fn.restype = args, res
    _lib, _emulation = lib, bool(lib.mvs_is_emulation())
    return lib


def set_knob(name: str, value: int = -1) -> None:
    """Test / tuning knob of the bound library (mvs_set_knob); value < 0 restores the default."""
    check(lib().mvs_set_knob(name.encode(), int(value)))


def lib() -> C.CDLL:
    return _lib if _lib is not None else bind()


def is_emulation() -> bool:
    lib()
    return _emulation


def dtype_code(dt: torch.dtype) -> int:
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError("unsupported storage dtype %s (fp32 / fp16 / bf16)" % dt) from None


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def stream_of(t: torch.Tensor) -> Optional[int]:
    """cudaStream_t of torch's current stream on t's device; tensors must be CUDA unless the emulation build is bound."""
    if t.is_cuda:
        if is_emulation():
            raise RuntimeError("the host-emulation library cannot run on CUDA tensors")
        return torch.cuda.current_stream(t.device).cuda_stream
    if not is_emulation():
        raise RuntimeError("mvs_b200 ops need CUDA tensors: the plane-sweep path has no CPU implementation "
                           "(got a %s tensor)" % t.device)
    return None


def check(rc: int) -> None:
    if rc != 0:
        raise RuntimeError("libmvs_b200: %s (code %d)" % (lib().mvs_last_error().decode(), rc))


def call(name: str, anchor: torch.Tensor, *args) -> None:
    """Invoke export `name` on anchor's device/stream (stream is appended as the last argument)."""
    global launches
    fn = getattr(lib(), name)
    st = stream_of(anchor)
    launches += 1
    if anchor.is_cuda:
        with torch.cuda.device(anchor.device):
            check(fn(*args, st))
    else:
        check(fn(*args, st))
