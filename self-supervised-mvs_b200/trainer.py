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

from .jdacs.models import augmentations


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


def draw_mask_box(h: int, w: int, generator: Optional[torch.Generator] = None) -> torch.Tensor:
    """Top-left corner (x, y) of the random H/3 x W/3 box random_image_mask zeroes (jdacs/models/augmentations.py:107-124),
    drawn on the host like the reference does; int64 [2]."""
    fh, fw = h // 3, w // 3
    x = torch.randint(0, w - fw, (1,), generator=generator)
    y = torch.randint(0, h - fh, (1,), generator=generator)
    return torch.cat((x, y))


def mask_reference_view(imgs: torch.Tensor, box: torch.Tensor):
    """random_image_mask on view 0: zero the H/3 x W/3 box whose corner `box` = (x, y) is a DEVICE tensor, so the step holds no
    host-side shape or index (it can be captured in a CUDA graph and replayed with a new box).
    -> (imgs with the masked reference, filter_mask [B,1,H,W] of ones with the box zeroed)."""
    b, n, c, h, w = imgs.shape
    fh, fw = h // 3, w // 3
    ys = torch.arange(h, device=imgs.device).view(h, 1) - box[1]
    xs = torch.arange(w, device=imgs.device).view(1, w) - box[0]
    inside = (ys >= 0) & (ys < fh) & (xs >= 0) & (xs < fw)
    mask = (~inside).to(imgs.dtype).view(1, 1, h, w)
    out = imgs.clone()
    out[:, 0] = imgs[:, 0] * mask
    return out, mask.expand(b, 1, h, w)


class TrainStep:
    def __init__(self, model: torch.nn.Module, criterion: torch.nn.Module, lr: float = 1e-3, w_aug: float = 0.01):
        self.model, self.criterion, self.w_aug = model, criterion, w_aug
        self.grads = FlatGrads(model.parameters())
        dev = self.grads.flat.device
        # capturable: the step counters live on the device, so optimizer.step() has no host read (CUDA-graph safe)
        self.opt = torch.optim.Adam(self.grads.params, lr=lr, betas=(0.9, 0.999), weight_decay=0.0, foreach=True,
                                    capturable=dev.type == "cuda")
        self.gen = torch.Generator().manual_seed(0)

    def _backward_and_step(self, loss: torch.Tensor) -> None:
        loss.backward()
        work = self.grads.allreduce()
        if work is not None:
            work.wait()
        self.opt.step()

    def run(self, imgs: torch.Tensor, imgs_aug: torch.Tensor, cams: torch.Tensor, proj_matrices: torch.Tensor,
            depth_values: torch.Tensor, box: torch.Tensor) -> Dict[str, torch.Tensor]:
        """The two optimiser steps of one batch; everything on the device, no host read (graph-capturable)."""
        m = self.model
        # ---- train_sample (jdacs/train.py:189-240), photometric term
        self.grads.zero()
        depth = m(imgs, proj_matrices, depth_values)["depth"]
        loss = self.criterion(imgs.float(), cams, depth)
        self._backward_and_step(loss)
        depth_est = depth.detach()
        # ---- train_sample_aug (jdacs/train.py:244-291): smooth-L1 against the first pass inside the un-masked region
        self.grads.zero()
        aug, fmask = mask_reference_view(imgs_aug, box)
        depth_aug = m(aug, proj_matrices, depth_values)["depth"]
        # models/augmentations.aug_loss: mean over the un-masked pixels, written without a boolean gather (no data-dependent shape)
        aug_loss = augmentations.aug_loss(depth_aug, depth_est, F.interpolate(fmask.float(), scale_factor=0.25)[:, 0]) * self.w_aug
        self._backward_and_step(aug_loss)
        return {"loss": loss.detach(), "augment_loss": aug_loss.detach()}

    def __call__(self, imgs: torch.Tensor, imgs_aug: torch.Tensor, cams: torch.Tensor, proj_matrices: torch.Tensor,
                 depth_values: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.model.train()
        box = draw_mask_box(imgs.shape[-2], imgs.shape[-1], self.gen).to(imgs.device, non_blocking=True)
        return self.run(imgs, imgs_aug, cams, proj_matrices, depth_values, box)


class GraphedTrainStep:
    """TrainStep.run captured ONCE into a CUDA graph (both forward / loss / backward / all-reduce / Adam passes, ~2000 kernel
    launches and as many Python-side calls) and replayed per batch: the inputs are copied into static buffers, the random mask
    box is drawn on the host and uploaded as two integers.  The eager step spends a third of its time on launch overhead
    (19.7 ms per item eager against 13.3 ms of kernels, profiles/r02_train_*).  A few eager steps on the first batch warm the
    library autotuners up before the capture -- they are ordinary optimiser steps."""

    def __init__(self, step: TrainStep, example: Dict[str, torch.Tensor], warmup: int = 3):
        self.step = step
        self.keys = ("imgs", "imgs_aug", "cams", "proj_matrices", "depth_values")
        dev = step.grads.flat.device
        self.static = {k: example[k].to(dev).clone() for k in self.keys}
        self.box = draw_mask_box(example["imgs"].shape[-2], example["imgs"].shape[-1], step.gen).to(dev)
        step.model.train()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(warmup):
                step.run(*[self.static[k] for k in self.keys], self.box)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.out = step.run(*[self.static[k] for k in self.keys], self.box)

    def load(self, batch: Dict[str, torch.Tensor]) -> None:
        """Stage a batch (device or pinned-host tensors) into the graph's static inputs and draw the next mask box."""
        for k in self.keys:
            self.static[k].copy_(batch[k], non_blocking=True)
        h, w = self.static["imgs"].shape[-2:]
        # two integers from pageable memory: the driver stages them before the call returns, so no host buffer to keep alive
        self.box.copy_(draw_mask_box(h, w, self.step.gen), non_blocking=True)

    def replay(self) -> Dict[str, torch.Tensor]:
        self.graph.replay()
        return self.out

    def __call__(self, imgs, imgs_aug, cams, proj_matrices, depth_values) -> Dict[str, torch.Tensor]:
        self.load(dict(zip(self.keys, (imgs, imgs_aug, cams, proj_matrices, depth_values))))
        return self.replay()
