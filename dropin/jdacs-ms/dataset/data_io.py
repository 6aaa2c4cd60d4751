"""Drop-in replacement for the reference's `jdacs-ms/dataset/data_io.py`: re-exports the B200 tree's PFM reader / writer.
Overlay this FILE only: the directory deliberately has no __init__.py, so the reference's own package file stays in place."""
import os
import sys

_root = os.environ.get("SSMVS_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)
import ssmvs_b200  # noqa: E402,F401  (loads ./self-supervised-mvs_b200)
from ssmvs_b200.jdacs_ms.dataset.data_io import read_pfm, save_pfm, save_pfm_flipped  # noqa: E402,F401
