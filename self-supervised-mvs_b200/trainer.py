"""One JDACS training batch on the plane-sweep path, as the reference's train.py drives it (jdacs/train.py:189-291):

    train_sample      forward -> UnSupLoss(imgs, cams, depth) -> backward -> optimizer.step()
    train_sample_aug  mask a random third of the reference image (models/augmentations.py:107-124) -> forward on the augmented
                      views -> smooth-L1 against the detached depth of the first pass (aug_loss) -> backward -> step()

i.e. two forward / backward / optimiser steps per batch.  The co-segmentation term (UnSupSegLoss: pretrained VGG19 + NMF) is
outside the plane-sweep path and is not part of this step (SURVEY.md hazard H10).  Multi-GPU: one process per GPU, the batch
sharded, ONE flat-bucket NCCL all-reduce of the gradients per optimiser step, issued asynchronously on NCCL's stream right after
backward; BatchNorm statistics stay per rank like nn.DataParallel's replicas (jdacs/train.py:65)."""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.distributed as dist
import torch.nn.functional as F


class FlatGrads:
    """Every parameter's .grad is a view into ONE flat fp32 buffer, so the gradient all-reduce is a single collective on memory
    autograd has already written: no flatten / unflatten copies around it."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, dtype=torch.float32, device=self.params[0].device)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self) -> None:
        self.flat.zero_()

    def allreduce(self, average: bool = True):
        """-> an async work handle (None on a single rank); wait() before the optimiser reads the gradients."""
        if not dist.is_initialized() or dist.get_world_size() == 1:
            return None
        if average:
            self.flat.div_(dist.get_world_size())
        return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)


def mask_reference_view(imgs: torch.Tensor, generator: Optional[torch.Generator] = None):
    """random_image_mask on view 0 (jdacs/models/augmentations.py:107-124): zero a random H/3 x W/3 box of the reference image.
    -> (imgs with the masked reference, filter_mask [B,3,H,W] of ones with the box zeroed)."""
    b, n, c, h, w = imgs.shape
    fh, fw = h // 3, w // 3
    x = int(torch.randint(0, w - fw, (1,), generator=generator))
    y = int(torch.randint(0, h - fh, (1,), generator=generator))
    mask = torch.ones(b, c, h, w, dtype=imgs.dtype, device=imgs.device)
    mask[:, :, y:y + fh, x:x + fw] = 0
    out = imgs.clone()
    out[:, 0] = imgs[:, 0] * mask
    return out, mask


class TrainStep:
    def __init__(self, model: torch.nn.Module, criterion: torch.nn.Module, lr: float = 1e-3, w_aug: float = 0.01):
        self.model, self.criterion, self.w_aug = model, criterion, w_aug
        self.grads = FlatGrads(model.parameters())
        self.opt = torch.optim.Adam(self.grads.params, lr=lr, betas=(0.9, 0.999), weight_decay=0.0, foreach=True)
        self.gen = torch.Generator().manual_seed(0)

    def _backward_and_step(self, loss: torch.Tensor) -> None:
        loss.backward()
        work = self.grads.allreduce()
        if work is not None:
            work.wait()
        self.opt.step()

    def __call__(self, imgs: torch.Tensor, imgs_aug: torch.Tensor, cams: torch.Tensor, proj_matrices: torch.Tensor,
                 depth_values: torch.Tensor) -> Dict[str, torch.Tensor]:
        m = self.model
        m.train()
        # ---- train_sample (jdacs/train.py:189-240), photometric term
        self.grads.zero()
        depth = m(imgs, proj_matrices, depth_values)["depth"]
        loss = self.criterion(imgs.float(), cams, depth)
        self._backward_and_step(loss)
        depth_est = depth.detach()
        # ---- train_sample_aug (jdacs/train.py:244-291)
        self.grads.zero()
        aug, fmask = mask_reference_view(imgs_aug, self.gen)
        depth_aug = m(aug, proj_matrices, depth_values)["depth"]
        fm = F.interpolate(fmask.float(), scale_factor=0.25)[:, 0] > 0.5
        aug_loss = F.smooth_l1_loss(depth_aug[fm], depth_est[fm]) * self.w_aug
        self._backward_and_step(aug_loss)
        return {"loss": loss.detach(), "augment_loss": aug_loss.detach()}
