"""Multi-GPU plumbing: one process per GPU (torchrun), batch items sharded across ranks.

Every batch item is an independent MVS problem, so inference has NO data-path collective: ranks only
meet at the timing barrier.  Training adds exactly one all-reduce of a flat gradient buffer per optimiser step
(1.35 MB for MVSNet, 2.21 MB for CVP-MVSNet: latency-bound over NVLink 5 / NVSwitch, so one bucket, not many).
BatchNorm statistics stay per rank, matching nn.DataParallel's per-replica statistics in the reference
(jdacs/train.py:65)."""
from __future__ import annotations

import os
from typing import Iterable, List, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str | None = None) -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from torchrun's environment; initialises the default process group if world > 1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def bind_to_gpu_numa(local_rank: int) -> bool:
    """Pin this process to the CPU cores closest to its GPU (NVML's affinity mask) so that pinned host buffers allocated
    afterwards are first-touched on that NUMA node: with 8 ranks uploading ~30 GB/s each, remote-node pinned memory halves
    the host->device rate.  Best effort: returns False (and changes nothing) when NVML or the affinity call is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            uuid = str(torch.cuda.get_device_properties(local_rank).uuid)
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return False
        os.sched_setaffinity(0, cpus)
        return True
    except Exception:
        return False


def shard_items(n_items: int, rank: int, world: int) -> range:
    """Contiguous, balanced slice of `n_items` batch items owned by `rank` (first ranks take the remainder)."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


def allreduce_gradients(params: Iterable[torch.nn.Parameter], average: bool = True) -> int:
    """One flat-bucket all-reduce over every gradient; returns the number of elements reduced."""
    grads: List[torch.Tensor] = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size() == 1:
        return sum(g.numel() for g in grads)
    flat = torch.cat([g.reshape(-1).float() for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    if average:
        flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off


def max_over_ranks(value: float, device: torch.device | str = "cpu") -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def barrier() -> None:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()
