"""Drop-in replacement for the reference's `jdacs-ms/losses/unsup_loss.py`: re-exports the B200 implementation."""
import os
import sys

_root = os.environ.get("SSMVS_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)
import ssmvs_b200  # noqa: E402,F401  (loads ./self-supervised-mvs_b200)
from ssmvs_b200.jdacs_ms.losses.unsup_loss import *  # noqa: E402,F401,F403
from ssmvs_b200.jdacs_ms.losses.unsup_loss import UnSupLoss  # noqa: E402,F401
