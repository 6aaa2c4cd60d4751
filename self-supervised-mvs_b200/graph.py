"""CUDA-graph wrapper for the inference forward.

The plane-sweep forward is ~25 short kernels; at B200 speeds the Python/driver launch path costs more than the
kernels.  `GraphedForward` captures one forward (library feature extractor + every libmvs_b200 launch, all issued on
torch's capture stream) into a CUDA graph with static input/output buffers; a step is then `copy inputs -> replay`.
Shapes, dtypes and the module's parameters must stay fixed between replays (weights are re-read from their tensors on
every replay, so in-place weight updates ARE seen: the tap tiles / folded-BN tensors are recomputed inside the graph
only if they were produced during capture; call `recapture()` after changing eval-mode weights).
"""
from __future__ import annotations

from typing import Callable, Dict, Sequence

import torch

from . import _lib


class GraphedForward:
    """`lanes` > 1 splits the batch (dim 0 of every input) into that many groups of items whose forwards are captured on
    separate streams: items are independent MVS problems, and every kernel of the path is either a persistent one-CTA-per-SM
    kernel with an uneven tail or (the plane sweep) a small-footprint kernel that fits beside one, so a second lane fills SMs
    the first leaves idle.  Outputs are concatenated back in item order."""

    def __init__(self, fn: Callable[..., Dict[str, torch.Tensor]], example_inputs: Sequence[torch.Tensor], warmup: int = 2,
                 lanes: int = 1):
        self.fn = fn
        self.static_in = [t.clone() for t in example_inputs]
        self.warmup = warmup
        self.lanes = max(1, min(int(lanes), self.static_in[0].shape[0]))
        self.recapture()

    def _run(self) -> Dict[str, torch.Tensor]:
        if self.lanes == 1:
            return self.fn(*self.static_in)
        cur = torch.cuda.current_stream()
        chunks = [torch.tensor_split(t, self.lanes, dim=0) for t in self.static_in]
        fork = torch.cuda.Event()
        fork.record(cur)
        outs = []
        for i in range(self.lanes):
            st = self._lane_streams[i]
            st.wait_event(fork)
            with torch.cuda.stream(st):
                outs.append(self.fn(*[c[i] for c in chunks]))
                done = torch.cuda.Event()
                done.record(st)
            cur.wait_event(done)
        return {k: torch.cat([o[k] for o in outs], 0) for k in outs[0]}

    def recapture(self) -> None:
        self._lane_streams = [torch.cuda.Stream() for _ in range(self.lanes)] if self.lanes > 1 else []
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side), torch.no_grad():
            self.fn(*self.static_in)          # single-lane pass first: fills the weight-tile / folded-BN caches the lanes share
            side.synchronize()
            for _ in range(self.warmup):
                self._run()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        before = _lib.launches
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.static_out = self._run()
        self.launches_per_replay = _lib.launches - before

    def __call__(self, *inputs: torch.Tensor) -> Dict[str, torch.Tensor]:
        for dst, src in zip(self.static_in, inputs):
            if dst.data_ptr() != src.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        _lib.launches += self.launches_per_replay
        return self.static_out


class StreamedForward:
    """Host-to-host inference pipeline: pinned host batch -> device -> forward -> pinned host result, double-buffered so that
    the upload of batch i+1 (copy stream) overlaps the kernels of batch i (compute stream).

        pipe = StreamedForward(fn, example_inputs, ("depth", "photometric_confidence"))
        for out in pipe.run(host_batches):      # out: dict of pinned host tensors, valid until the next iteration
            ...

    Two CUDA graphs of the same forward are captured, one per input buffer set, so a batch is uploaded straight into the
    static inputs of the graph that will consume it (no device-to-device staging copy).  Every batch is copied host->device
    and every result device->host; nothing is cached between batches."""

    def __init__(self, fn: Callable[..., Dict[str, torch.Tensor]], example_inputs: Sequence[torch.Tensor], out_keys: Sequence[str],
                 warmup: int = 2):
        self.graphs = [GraphedForward(fn, example_inputs, warmup=warmup) for _ in range(2)]
        self.keys = tuple(out_keys)
        self.copy = torch.cuda.Stream()
        self.d2h = torch.cuda.Stream()                  # results leave on their own stream: the next replay does not queue behind them
        self.uploaded = [torch.cuda.Event() for _ in range(2)]
        self.consumed = [torch.cuda.Event() for _ in range(2)]
        g = self.graphs[0]
        self.out_host = [{k: torch.empty(g.static_out[k].shape, dtype=g.static_out[k].dtype).pin_memory() for k in self.keys}
                         for _ in range(2)]
        self.out_done = [torch.cuda.Event() for _ in range(2)]
        self.h2d_bytes = sum(t.numel() * t.element_size() for t in g.static_in)
        self.d2h_bytes = sum(t.numel() * t.element_size() for t in self.out_host[0].values())
        self.launches_per_step = g.launches_per_replay

    def upload(self, i: int, host_inputs: Sequence[torch.Tensor]) -> None:
        s = i & 1
        self.copy.wait_event(self.consumed[s])          # the forward of batch i-2 has finished with this input set
        with torch.cuda.stream(self.copy):
            for dst, src in zip(self.graphs[s].static_in, host_inputs):
                dst.copy_(src, non_blocking=True)
            self.uploaded[s].record(self.copy)

    def forward(self, i: int) -> Dict[str, torch.Tensor]:
        s = i & 1
        g = self.graphs[s]
        cur = torch.cuda.current_stream()
        cur.wait_event(self.uploaded[s])
        g.graph.replay()
        _lib.launches += g.launches_per_replay
        self.consumed[s].record(cur)                    # inputs consumed, outputs produced
        # device -> host on the d2h stream.  Graph s is replayed again only at step i + 2, after run() has synchronised
        # out_done[s] to hand batch i out, so its static outputs are not overwritten under the copy.
        self.d2h.wait_event(self.consumed[s])
        with torch.cuda.stream(self.d2h):
            for k in self.keys:
                self.out_host[s][k].copy_(g.static_out[k], non_blocking=True)
            self.out_done[s].record(self.d2h)
        return self.out_host[s]

    def run(self, host_batches):
        """Yield the host result of every batch, in order.  One forward stays in flight behind the one being handed out and
        one upload ahead of it, so the GPU never waits for the host; a yielded dict is valid until the next iteration."""
        it = iter(host_batches)
        nxt = next(it, None)
        if nxt is None:
            return
        self.upload(0, nxt)
        i, pending = 0, None
        while nxt is not None:
            nxt = next(it, None)
            if nxt is not None:
                self.upload(i + 1, nxt)
            out = self.forward(i)
            if pending is not None:
                self.out_done[(i - 1) & 1].synchronize()
                yield pending
            pending = out
            i += 1
        self.out_done[(i - 1) & 1].synchronize()
        yield pending
