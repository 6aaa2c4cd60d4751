"""Mirror of jdacs-ms `losses/homography.py` (:186-243): same inverse_warping as the jdacs tree."""
from ...jdacs.losses.homography import inverse_warping  # noqa: F401
