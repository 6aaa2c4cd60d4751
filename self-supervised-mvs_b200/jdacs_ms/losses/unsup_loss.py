"""Mirror of jdacs-ms `losses/unsup_loss.py` (:18-86): no x0.25 resize (images arrive at depth-map size) and a 0.05
smoothness weight (:82)."""
from ...jdacs.losses.unsup_loss import UnSupLoss as _Base


class UnSupLoss(_Base):
    def __init__(self):
        super().__init__(downscale=False, smooth_weight=0.05, smooth_lambda=1.0)
