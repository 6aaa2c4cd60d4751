"""Mirror of jdacs-ms `losses/modules.py` (identical to the jdacs tree up to imports)."""
from ...jdacs.losses.modules import *  # noqa: F401,F403
from ...jdacs.losses.modules import SSIM, compute_reconstr_loss, depth_smoothness, gradient, gradient_x, gradient_y  # noqa: F401
