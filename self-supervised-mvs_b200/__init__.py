"""B200-native plane-sweep hot path of MVSNet / CVP-MVSNet (ToughStoneX/Self-Supervised-MVS drop-in).

Layout of the package (only what the path needs):
  csrc/            CUDA kernels (sm_100a) + the C ABI declared in include/mvs_b200.h -> libmvs_b200.so
  _lib.py          ctypes binding (no fallback: ops raise when the library or a CUDA device is missing)
  ops.py           torch.autograd wrappers over the C ABI
  regnet.py        3-D U-Net runner (eval: folded BN fused in the conv epilogue; train: batch-stat BN kernels)
  jdacs/           mirror of the reference's jdacs/{models,losses} import surface (MVSNet, UnSupLoss, ...)
  jdacs_ms/        mirror of jdacs-ms/{models,losses} (CVPMVSNet, proj_cost, calDepthHypo, ...)
  synth.py         deterministic DTU-shaped synthetic inputs (tests, bench)
The directory name carries a hyphen, so it is imported through the top-level shim module `ssmvs_b200`.
"""
from . import _lib, ops, synth  # noqa: F401

__all__ = ["_lib", "ops", "synth"]
