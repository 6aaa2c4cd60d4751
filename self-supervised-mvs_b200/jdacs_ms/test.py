"""Mirror of the writer half of jdacs-ms `test.py` (:96-165): the finest depth map and the probability confidence of a batch are
written as `depth_est/*.pfm`, `confidence/*.pfm` and the 8-bit `*.pfm.png` preview -- no resize (CVP-MVSNet's finest level is at
image resolution), the rows flipped into .pfm order on the GPU, one D2H copy per batch."""
from __future__ import annotations

from ..jdacs.eval_dense import save_depth_outputs, write_depth_img  # noqa: F401
from .dataset.data_io import read_pfm, save_pfm  # noqa: F401


def save_outputs(outputs, filenames, outdir, preview=True):
    """outputs = CVPMVSNet.forward's dict ("depth_est_list" finest first, "prob_confidence"); filenames as the reference's loader
    yields them ("scan1/{}/00000000{}")."""
    depth = outputs["depth_est_list"][0]
    save_depth_outputs({"depth": depth, "photometric_confidence": outputs["prob_confidence"]}, filenames, outdir,
                       size=tuple(depth.shape[-2:]), preview=preview)
