"""Drop-in for the two functions train.py imports from the reference's `jdacs/models/augmentations.py` (random_image_mask, aug_loss).
NOTE: the reference file also holds the PIL colour / blur transforms its data loader uses; keep those by appending this import to
the reference file instead of overwriting it if the loader's augmentation is on."""
import os
import sys

_root = os.environ.get("SSMVS_B200_ROOT") or os.path.dirname(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
if _root not in sys.path:
    sys.path.insert(0, _root)
import ssmvs_b200  # noqa: E402,F401
from ssmvs_b200.jdacs.models.augmentations import random_image_mask, aug_loss  # noqa: E402,F401
