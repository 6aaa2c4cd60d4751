#!/usr/bin/env python
"""bench.py — depth-samples/sec of the MVSNet plane-sweep path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--dtype fp16|bf16|fp32]
  torchrun --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...      (one rank per GPU)

A step = one MVSNet.forward (eval) over one synthetic DTU-shaped item per GPU at BASELINE.json configs[1]:
N=5 views, 512x640 images -> 128x160 maps, C=32, D=192  (3,932,160 depth-samples per item; weak scaling, no
data-path collective: items are independent).  `value` times it with the inputs resident in HBM, `e2e` through the
same reference-facing call with pinned HOST inputs and a device->host read of depth + confidence every step.
`roofline` is measured live with CUDA events for the dominant stage, `cpu_baseline` is the oracle port (the
reference's own torch-CPU op sequence) on this box's host cores.  `--impl reference` times that CPU path alone.
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

VIEWS, HEIGHT, WIDTH, NDEPTH, CHANNELS = 5, 512, 640, 192, 32
HF, WF = HEIGHT // 4, WIDTH // 4
SAMPLES_PER_ITEM = NDEPTH * HF * WF
WORKLOAD = "MVSNet forward N=5 512x640 D=192 (BASELINE.json configs[1])"
METRIC = "depth-samples/sec (N=5,D=192,512x640)"
# algorithmic work per item: warp+variance bytes = s*(C*D*Hf*Wf + N*C*Hf*Wf) + 4*D ; CostRegNet flops
REG_FLOPS = 20304 * SAMPLES_PER_ITEM
CONV0_FLOPS = 2 * 27 * 32 * 8 * SAMPLES_PER_ITEM   # conv0 of CostRegNet: 6912 MAC per depth-sample
SOFTARGMIN_BYTES = 4 * SAMPLES_PER_ITEM + 12 * HF * WF


def CONV0_BYTES(elem: int) -> int:
    """conv0 of CostRegNet: read the 32-channel variance volume once, write the 8-channel activation once."""
    return elem * (CHANNELS + 8) * SAMPLES_PER_ITEM


def warp_var_bytes(elem: int) -> int:
    return elem * (CHANNELS * SAMPLES_PER_ITEM + VIEWS * CHANNELS * HF * WF) + 4 * NDEPTH


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p["bf16_tflops_sustained"],
                "source": "measured"}
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def load_oracle():
    spec = importlib.util.spec_from_file_location("planesweep_oracle", os.path.join(ROOT, "oracle", "planesweep.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_model(dtype):
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    from ssmvs_b200 import synth
    torch.manual_seed(0)
    model = MVSNet(refine=False, volume_dtype=dtype)
    synth.randomise_bn(model, 5)                          # non-trivial BN statistics and a scaled last layer: a non-uniform
    with torch.no_grad():                                 # softmax (hazard H11); the work is the same either way, and the
        model.cost_regularization.prob.weight.mul_(64.0)  # parity leg (parity_check) calibrates and asserts the peakiness itself
    return model.eval()


class ClockSampler:
    """SM clock / throttle reasons sampled WHILE the timed region runs: NVML polled from a thread every few ms (the timed
    region is tens of ms, too short for `nvidia-smi -lms`), with an nvidia-smi loop as the fallback."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, index: int, period_s: float = 0.004):
        self.index, self.period = index, period_s
        self.sm, self.mx, self.reasons, self.how = [], 0.0, set(), None
        self.stop = threading.Event()
        self.t = None
        self.proc = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        try:   # CUDA_VISIBLE_DEVICES may renumber devices: find ours by UUID
            uuid = str(torch.cuda.get_device_properties(self.index).uuid)
            return pynvml, pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
        except Exception:
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(self.index)

    def _poll_nvml(self, nv, h):
        bits = ((nv.nvmlClocksEventReasonHwSlowdown, "hw_slowdown"), (nv.nvmlClocksEventReasonHwThermalSlowdown, "hw_thermal_slowdown"),
                (nv.nvmlClocksEventReasonSwThermalSlowdown, "sw_thermal_slowdown"), (nv.nvmlClocksEventReasonSwPowerCap, "sw_power_cap"))
        while not self.stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                for b, n in bits:
                    if r & b:
                        self.reasons.add(n)
            except Exception:
                pass
            time.sleep(self.period)

    def _poll_smi(self):
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                self.sm.append(float(f[0]))
                self.mx = max(self.mx, float(f[1]))
            except ValueError:
                continue
            for n, v in zip(self.NAMES, f[2:6]):
                if v.lower().startswith("active"):
                    self.reasons.add(n)

    def __enter__(self):
        try:
            nv, h = self._nvml_handle()
            self.mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
            self.how = "nvml"
            self.t = threading.Thread(target=self._poll_nvml, args=(nv, h), daemon=True)
        except Exception:
            try:
                self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                              "--format=csv,noheader,nounits", "-lms", "20"], stdout=subprocess.PIPE,
                                             stderr=subprocess.DEVNULL, text=True)
                self.how = "nvidia-smi"
                self.t = threading.Thread(target=self._poll_smi, daemon=True)
            except Exception:
                self.t = None
        if self.t is not None:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc is not None:
            self.proc.terminate()
        if self.t is not None:
            self.t.join(timeout=2)

    def summary(self):
        return {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.mx or None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": self.how}


def cpu_forward_timer(steps: int, warmup: int):
    """The reference's CPU path (oracle port: same torch-CPU op sequence as jdacs/models/mvsnet.py:105-155)."""
    from ssmvs_b200 import synth
    oracle = load_oracle()
    cores = os.cpu_count() or 1
    model = make_model(torch.float32)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    threads = int(os.environ.get("MVS_CPU_THREADS", "0"))
    if not threads:
        # torch's CPU kernels stop scaling (and regress) on many-core hosts for this op mix: give the reference its best
        # thread count, probed on a quarter-depth problem (48 planes), instead of blindly using every core
        probe = synth.mvsnet_inputs(1, VIEWS, HEIGHT, WIDTH, 48, seed=0)
        best = None
        for n in sorted({cores, min(cores, 64), min(cores, 32), min(cores, 16)}, reverse=True):
            torch.set_num_threads(n)
            with torch.no_grad():
                oracle.mvsnet_forward(probe["imgs"], probe["proj_matrices"], probe["depth_values"], sd)
                t0 = time.perf_counter()
                oracle.mvsnet_forward(probe["imgs"], probe["proj_matrices"], probe["depth_values"], sd)
                dt = time.perf_counter() - t0
            if best is None or dt < best[0]:
                best = (dt, n)
        threads = best[1]
    torch.set_num_threads(threads)
    inp = synth.mvsnet_inputs(1, VIEWS, HEIGHT, WIDTH, NDEPTH, seed=0)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], sd)
            if i >= warmup:
                times.append(time.perf_counter() - t0)
    return times, threads


def gpu_library_timer(dev, steps: int = 3):
    """The reference's op sequence (oracle port: ATen grid_sample + elementwise variance + cuDNN conv3d + softmax, fp32, torch
    defaults) executed by PyTorch ON THE GPU, one item per forward: the library baseline the hand-written kernels replace."""
    from ssmvs_b200 import synth
    oracle = load_oracle()
    model = make_model(torch.float32)
    sd = {k: v.detach().clone().to(dev) for k, v in model.state_dict().items()}
    inp = {k: v.to(dev) for k, v in synth.mvsnet_inputs(1, VIEWS, HEIGHT, WIDTH, NDEPTH, seed=0).items()}
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    with torch.no_grad(), torch.device(dev):
        oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], sd)
        torch.cuda.synchronize(dev)
        for a, b in ev:
            a.record()
            oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], sd)
            b.record()
        torch.cuda.synchronize(dev)
    torch.cuda.empty_cache()
    return [a.elapsed_time(b) * 1e-3 for a, b in ev]


def parity_check(dev, dtype, name):
    """The product path at the benchmarked precision against the fp32 oracle run on this GPU (checker only, outside every timed
    region), at the workload's full size, on a probability volume asserted to be peaky (tests/parity_util.py, hazard H11)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import parity_util as pu
    try:
        model, inp, want, cond = pu.peaky_mvsnet(dev, VIEWS, HEIGHT, WIDTH, NDEPTH, seed=0, target_peak=0.3)
        assert cond["peak"] >= 0.3 and cond["depth_std"] >= 10.0, cond
        r = pu.depth_parity(pu.product_mvsnet(model, inp, dtype), want)
        out = {"against": "oracle/planesweep.py mvsnet_forward in fp32 on this GPU (TF32 off), 1 item of the workload", "dtype": name,
               "conditions": cond, "depth_rel_max": r["depth_rel_max"], "depth_rel_p999": r["depth_rel_p999"],
               "depth_rel_median": r["depth_rel_median"], "frac_pixels_within_1e-3": r["frac_within_1e-3"],
               "index_mismatch": r["index_mismatch"], "index_mismatch_near_integer": r["index_mismatch_near_integer"],
               "index_off_by_more_than_1": r["index_off_by_more_than_1"], "pixels": r["pixels"],
               "conf_abs_max": r["conf_abs_max"], "conf_abs_p999": r["conf_abs_p999"]}
        if dtype != torch.float32:
            ideal = pu.depth_parity(pu.ideal_storage_mvsnet(model, inp, dtype), want)
            out["storage_floor"] = {"what": "the oracle's fp32 arithmetic with every stored tensor rounded once to %s" % name,
                                    "depth_rel_max": ideal["depth_rel_max"], "depth_rel_p999": ideal["depth_rel_p999"],
                                    "index_mismatch": ideal["index_mismatch"]}
        tf = pu.depth_parity(pu.oracle_mvsnet(model, inp, tf32=True)[0], want)
        out["oracle_with_torch_default_tf32"] = {"depth_rel_max": tf["depth_rel_max"], "depth_rel_p999": tf["depth_rel_p999"],
                                                 "index_mismatch": tf["index_mismatch"]}
        return out
    except Exception as exc:
        return {"error": "%s: %s" % (type(exc).__name__, exc)}
    finally:
        torch.cuda.empty_cache()


def run_reference(args, rank):
    if rank != 0:
        return
    times, threads = cpu_forward_timer(args.steps, min(args.warmup, 1))
    total = sum(times)
    val = SAMPLES_PER_ITEM * len(times) / total
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "depth-samples/s", "n_gpus": args.gpus,
            "steps": len(times), "warmup": min(args.warmup, 1), "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "global_batch": 1, "sample": "1 item per step on host cores"},
            "cpu_baseline": {"value": val, "unit": "depth-samples/s", "cores": threads, "kind": "port",
                             "sample": "%d full forward passes of 1 item (oracle port of the reference's torch-CPU path, fp32; best of the probed thread counts, host has %d cores)" % (len(times), os.cpu_count() or 1)},
            "e2e": {"value": val, "unit": "depth-samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ================================================================================================ secondary workloads
def _timed_steps(fn, steps, warmup, dev, flush, parallel):
    """W warm-up steps, then K steps bracketed per step by CUDA events on the launching stream (L2 flushed between steps, outside
    the event pairs), barrier + synchronize on both sides; -> (sum of step ms, max over ranks)."""
    stream = torch.cuda.current_stream(dev)
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize(dev)
    parallel.barrier()
    torch.cuda.synchronize(dev)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        flush.zero_()
        a.record(stream)
        fn()
        b.record(stream)
    torch.cuda.synchronize(dev)
    parallel.barrier()
    return parallel.max_over_ranks(sum(a.elapsed_time(b) for a, b in ev), dev)


def _event_ms(fn, dev, flush, reps=3):
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        torch.cuda.synchronize(dev)
        ts.append(a.elapsed_time(b))
    return statistics.median(ts)


def cpu_train_timer(ndepth=48):
    """The reference's CPU path for one train_sample-equivalent (forward in train mode + UnSupLoss + backward + Adam) through the
    oracle port, on a bounded sample: 1 item, full image size, `ndepth` planes."""
    from ssmvs_b200 import synth
    oracle = load_oracle()
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    torch.manual_seed(0)
    model = MVSNet(refine=False)
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and "running" not in k) for k, v in model.state_dict().items()}
    leaves = [v for v in params.values() if v.requires_grad]
    opt = torch.optim.Adam(leaves, lr=1e-3)
    inp = synth.mvsnet_inputs(1, VIEWS, HEIGHT, WIDTH, ndepth, seed=0)
    threads = min(os.cpu_count() or 1, 32)
    torch.set_num_threads(threads)
    t0 = time.perf_counter()
    out = oracle.mvsnet_forward(inp["imgs"], inp["proj_matrices"], inp["depth_values"], params, True)
    loss = oracle.unsup_loss(inp["imgs"], inp["cams"], out["depth"])
    loss = loss[0] if isinstance(loss, (tuple, list)) else (loss["total"] if isinstance(loss, dict) else loss)
    opt.zero_grad()
    loss.backward()
    opt.step()
    return time.perf_counter() - t0, threads, ndepth


def run_train(args):
    """BASELINE configs[3]: one JDACS training batch = train_sample + train_sample_aug (two forward / backward / Adam steps,
    jdacs/train.py:189-291) with the photometric loss, N = 5 views (the reference's top-3 view selection needs >= 3 sources,
    SURVEY hazard H5), 512x640, D = 192; the batch is sharded over the ranks and the gradients take one NCCL all-reduce per
    optimiser step."""
    import ssmvs_b200
    from ssmvs_b200 import ops, parallel, synth
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    from ssmvs_b200.trainer import GraphedTrainStep, TrainStep
    rank, world, local = parallel.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ssmvs_b200._lib.bind()
    for kv in args.knob:
        ssmvs_b200._lib.set_knob(kv.split("=")[0], int(kv.split("=")[1]))
    tdt = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.train_dtype]
    PB = args.batch if args.batch_given else 4      # items per GPU per batch (measured 2 / 4 / 8: 10.7 / 9.5 / 8.9 ms per item; the reference fits 1 on 11 GB)
    torch.manual_seed(0)
    model = MVSNet(refine=False, train_dtype=tdt).to(dev)
    step = TrainStep(model, UnSupLoss())
    host = synth.mvsnet_inputs(PB, VIEWS, HEIGHT, WIDTH, NDEPTH, seed=rank)
    host["imgs_aug"] = host["imgs"] + 0.05 * torch.randn(host["imgs"].shape, generator=torch.Generator().manual_seed(100 + rank))
    keys = ("imgs", "imgs_aug", "cams", "proj_matrices", "depth_values")
    pinned = {k: host[k].pin_memory() for k in keys}
    res = {k: v.to(dev) for k, v in pinned.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    loss_host = torch.empty(2).pin_memory()

    graphed = None
    if not args.no_graph:
        graphed = GraphedTrainStep(step, res, warmup=3)      # the whole batch (2 x forward / loss / backward / all-reduce / Adam) as one CUDA graph
        graphed.load(res)

    def step_resident():
        if graphed is not None:
            return graphed.replay()
        return step(res["imgs"], res["imgs_aug"], res["cams"], res["proj_matrices"], res["depth_values"])

    def step_e2e():
        if graphed is not None:
            graphed.load(pinned)                             # H2D from pinned memory straight into the graph's static inputs
            o = graphed.replay()
        else:
            d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
            o = step(d["imgs"], d["imgs_aug"], d["cams"], d["proj_matrices"], d["depth_values"])
        loss_host.copy_(torch.stack((o["loss"], o["augment_loss"])), non_blocking=True)

    l0 = ssmvs_b200._lib.launches
    if graphed is None:
        step_resident()
    else:                                                    # count the C-ABI launches of one batch eagerly (a replay makes the same ones)
        step(res["imgs"], res["imgs_aug"], res["cams"], res["proj_matrices"], res["depth_values"])
    per_step_launches = ssmvs_b200._lib.launches - l0
    with ClockSampler(local) as clk:
        ms_total = _timed_steps(step_resident, args.steps, args.warmup, dev, flush, parallel)
        launches = per_step_launches * args.steps
        ms_e2e = _timed_steps(step_e2e, args.steps, 1, dev, flush, parallel)
    clocks = clk.summary()
    losses = step_resident()
    assert torch.isfinite(losses["loss"]) and torch.isfinite(losses["augment_loss"]), losses

    # ---- where the time goes: phases of the first pass, and the two dominant kernels alone
    mk = lambda: torch.cuda.Event(enable_timing=True)
    e = [mk() for _ in range(5)]
    step.grads.zero()
    flush.zero_(); e[0].record()
    depth = model(res["imgs"], res["proj_matrices"], res["depth_values"])["depth"]
    e[1].record()
    loss = step.criterion(res["imgs"], res["cams"], depth)
    e[2].record()
    loss.backward()
    e[3].record()
    step.opt.step()
    e[4].record()
    torch.cuda.synchronize(dev)
    stages = {k: e[i].elapsed_time(e[i + 1]) for i, k in enumerate(("forward", "loss", "backward", "allreduce+adam"))}
    peaks = load_peaks()
    rooflines = {}
    if tdt != torch.float32:
        elem = 2
        var = torch.randn(PB, CHANNELS // 8, NDEPTH, HF, WF, 8, device=dev).to(tdt)
        gz = torch.randn(PB, 1, NDEPTH, HF, WF, 8, device=dev).to(tdt)
        w0 = torch.zeros(8, CHANNELS, 3, 3, 3, device=dev)
        t_wg = _event_ms(lambda: ops._wgrad_mma(var, gz, w0, 8, 1, False, 8), dev, flush)
        tf = PB * CONV0_FLOPS / (t_wg * 1e-3) / 1e12
        rooflines["roofline_wgrad_conv0"] = {"kernel": "conv3d_wgrad_mma_kernel (conv0: 32 x 8 x 27 taps over D x H x W, mma.sync m16n8k16)", "bound": "tensor",
                                             "achieved": tf, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops"], "traffic": None,
                                             "ms": t_wg, "peak_source": peaks["source"], "algorithmic_flops": PB * CONV0_FLOPS,
                                             "hbm": {"achieved": PB * CONV0_BYTES(elem) / (t_wg * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                                                     "frac": PB * CONV0_BYTES(elem) / (t_wg * 1e-3) / 1e9 / peaks["hbm_gbs"]}}
        feats = [torch.randn(PB, CHANNELS, HF, WF, device=dev).requires_grad_(True) for _ in range(VIEWS)]
        rt = ops.compose_proj(res["proj_matrices"])
        v = ops.warp_variance(feats[0], feats[1:], rt, res["depth_values"], tdt)
        gv = torch.randn_like(v)
        t_wb = _event_ms(lambda: torch.autograd.grad(v, feats, gv, retain_graph=True), dev, flush)
        bwd_bytes = PB * (elem * CHANNELS * SAMPLES_PER_ITEM + VIEWS * CHANNELS * HF * WF * (elem + 4))
        gbs = bwd_bytes / (t_wb * 1e-3) / 1e9
        rooflines["roofline_sweep_bwd"] = {"kernel": "warp_var_bwd16s_kernel (sweep backward, one thread per (pixel, channel block, source); + unpack of the 5 gradient maps)", "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"],
                                           "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"], "traffic": None, "ms": t_wb, "peak_source": peaks["source"],
                                           "algorithmic_bytes": bwd_bytes}
    dominant = max(rooflines.values(), key=lambda r: r["ms"]) if rooflines else None

    if rank == 0:
        items = world * args.steps * PB
        line = {"metric": "depth-samples/sec through one JDACS training batch (train_sample + train_sample_aug: 2 x forward/backward/Adam), N=5, D=192, 512x640",
                "value": items * SAMPLES_PER_ITEM / (ms_total * 1e-3), "unit": "depth-samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "ms_per_item": ms_total / args.steps / PB, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "%s activations, f32 master weights / accumulation, f64 BatchNorm statistics" % args.train_dtype,
                "data": "synthetic",
                "config": {"workload": "JDACS training batch, photometric loss (BASELINE.json configs[3]; N=5 per SURVEY H5; co-segmentation term out of scope, H10)",
                           "global_batch": world * PB, "per_gpu_batch": PB,
                           "parallelism": "dp%d: batch sharded, one flat-bucket NCCL all-reduce (%.2f MB fp32) per optimiser step, 2 per batch" % (world, step.grads.flat.numel() * 4 / 1e6),
                           "l2": "256 MiB buffer written between timed steps (outside the per-step event pairs)",
                           "launch": "python (no graph)" if graphed is None else "cuda-graph replay of the whole batch (%d C-ABI launches + library kernels per batch)" % per_step_launches},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": items * SAMPLES_PER_ITEM / (ms_e2e * 1e-3), "unit": "depth-samples/s",
                        "h2d_bytes_per_step": world * sum(v.numel() * v.element_size() for v in pinned.values()),
                        "d2h_bytes_per_step": world * loss_host.numel() * 4, "ms_per_step": ms_e2e / args.steps,
                        "how": "every step copies imgs, imgs_aug, cams, proj_matrices, depth_values from pinned host memory and reads both losses back"},
                "roofline": dominant, "stage_ms_first_pass": stages, "losses": {k: float(v) for k, v in losses.items()}}
        line.update(rooflines)
        if world == 1 and not args.no_cpu_baseline:
            try:
                sec, threads, nd = cpu_train_timer()
                line["cpu_baseline"] = {"value": nd * HF * WF / sec, "unit": "depth-samples/s", "cores": threads, "kind": "port",
                                        "sample": "one train_sample-equivalent (forward in train mode + UnSupLoss + backward + Adam) of 1 item at 512x640 with %d of the 192 planes, oracle port on the host cores; the GPU step does two such passes per batch" % nd}
            except Exception as exc:
                line["cpu_baseline"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
        print(json.dumps(line))
    parallel.barrier()


CVP_NSRC, CVP_NSCALE = 4, 3
CVP_SAMPLES = 48 * 128 * 160 + 8 * 256 * 320 + 8 * 512 * 640
CVP_FLOPS = 132192 * CVP_SAMPLES + 2 * 121536 * 512 * 640 * (1 + CVP_NSRC) * (1 + 0.25 + 0.0625)   # 3 regularisations + feature pyramid


def run_cvp(args):
    """BASELINE configs[2]: CVP-MVSNet, 3 pyramid levels, 1 + 4 views of 512x640, bf16 volumes, eval."""
    from types import SimpleNamespace
    import ssmvs_b200
    from ssmvs_b200 import parallel, synth
    from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet
    rank, world, local = parallel.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ssmvs_b200._lib.bind()
    for kv in args.knob:
        ssmvs_b200._lib.set_knob(kv.split("=")[0], int(kv.split("=")[1]))
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype if args.dtype_given else "bf16"]
    PB = args.batch if args.batch_given else 4
    torch.manual_seed(0)
    model = CVPMVSNet(SimpleNamespace(nsrc=CVP_NSRC, nscale=CVP_NSCALE, mode="test"), volume_dtype=dtype)
    synth.randomise_bn(model, 5)
    model = model.to(dev).eval()
    keys = ("ref_img", "src_imgs", "ref_in", "src_in", "ref_ex", "src_ex", "depth_min", "depth_max")
    host = synth.cvp_inputs(PB, CVP_NSRC, HEIGHT, WIDTH, seed=rank)
    pinned = {k: host[k].pin_memory() for k in keys}
    res = {k: v.to(dev) for k, v in pinned.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    out_host = {"depth": torch.empty(PB, HEIGHT, WIDTH).pin_memory(), "conf": torch.empty(PB, HEIGHT, WIDTH).pin_memory()}

    graphed = None
    if not args.no_graph:
        from ssmvs_b200.graph import GraphedForward
        graphed = GraphedForward(lambda *a: model(*a), [res[k] for k in keys])    # the whole three-level forward as one CUDA graph

    def fwd(d):
        if graphed is not None:
            return graphed(*[d[k] for k in keys])          # copies into the static inputs (a no-op for the resident tensors), replays
        with torch.no_grad():
            return model(*[d[k] for k in keys])

    def step_e2e():
        d = pinned if graphed is not None else {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
        o = fwd(d)
        out_host["depth"].copy_(o["depth_est_list"][0], non_blocking=True)
        out_host["conf"].copy_(o["prob_confidence"], non_blocking=True)

    if graphed is not None:
        res = {k: t for k, t in zip(keys, graphed.static_in)}     # resident inputs = the graph's own static buffers
    l0 = ssmvs_b200._lib.launches
    with ClockSampler(local) as clk:
        ms_total = _timed_steps(lambda: fwd(res), args.steps, args.warmup, dev, flush, parallel)
        launches = ssmvs_b200._lib.launches - l0
        ms_e2e = _timed_steps(step_e2e, args.steps, 1, dev, flush, parallel)
    clocks = clk.summary()
    peaks = load_peaks()
    if rank == 0:
        items = world * args.steps * PB
        tf = items * CVP_FLOPS / (ms_total * 1e-3) / 1e12 / world
        line = {"metric": "depth-samples/sec (CVP-MVSNet, 3 levels, nsrc=4, 512x640)", "value": items * CVP_SAMPLES / (ms_total * 1e-3), "unit": "depth-samples/s",
                "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "ms_per_item": ms_total / args.steps / PB,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": {torch.float16: "f16", torch.bfloat16: "bf16", torch.float32: "f32"}[dtype] + " storage, f32 accumulate", "data": "synthetic",
                "config": {"workload": "CVP-MVSNet forward, 3 pyramid levels, 1 + 4 views of 512x640 (BASELINE.json configs[2])", "global_batch": world * PB,
                           "per_gpu_batch": PB, "parallelism": "dp%d (items sharded, no collective)" % world,
                           "l2": "256 MiB buffer written between timed steps (outside the per-step event pairs)",
                           "launch": "python (no graph)" if graphed is None else "cuda-graph replay (%d C-ABI launches per step)" % graphed.launches_per_replay},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": items * CVP_SAMPLES / (ms_e2e * 1e-3), "unit": "depth-samples/s",
                        "h2d_bytes_per_step": world * sum(v.numel() * v.element_size() for v in pinned.values()),
                        "d2h_bytes_per_step": world * sum(v.numel() * 4 for v in out_host.values()), "ms_per_step": ms_e2e / args.steps},
                "roofline": {"kernel": "whole forward: feature pyramid + 3 x CostRegNet on conv3d_tc_kernel (86 % of the step), sweeps, soft-argmin", "bound": "tensor",
                             "achieved": tf, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": tf / peaks["bf16_tflops_sustained"], "traffic": None,
                             "peak_source": peaks["source"], "algorithmic_flops": PB * CVP_FLOPS}}
        print(json.dumps(line))
    parallel.barrier()


def run_output(args):
    """The inference output side at the size eval_dense.py writes (1200 x 1600 maps): the geometric-consistency
    filter of one reference view against 10 source views per step (jdacs/eval_dense.py:177-232) is the timed path; the nearest
    resize of depth + confidence (:150-153) and the fusibile consensus kernel (fusion/fusibile/fusibile.cu:138-277) are timed
    beside it.  cpu_baseline = the NumPy restatement of the reference's own functions (oracle/output_side.py, pinned to them by
    tests/golden/jdacs_output_side.npz) on the host, one pair."""
    import importlib.util
    import numpy as np
    import ssmvs_b200
    from ssmvs_b200 import ops, parallel, synth
    from ssmvs_b200.jdacs.eval_dense import _pair_cams
    from ssmvs_b200.jdacs.fusion import fusibile as fz
    rank, world, local = parallel.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    ssmvs_b200._lib.bind()
    H, W, S = 1200, 1600, 10
    k = synth.intrinsics(W, H).astype(np.float32)
    ex = [synth.extrinsics(v).astype(np.float32) for v in range(5)]
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    d_ref = (620 + 40 * np.sin(xx / 90.0) * np.cos(yy / 70.0) + rank).astype(np.float32)
    d_src = np.stack([(d_ref + np.float32(0.2 * s)) for s in range(S)]).astype(np.float32)
    cams = torch.from_numpy(np.stack([_pair_cams(k, ex[0], k, ex[1 + s % 4]) for s in range(S)])).to(dev)
    pin_ref, pin_src = torch.from_numpy(d_ref).pin_memory(), torch.from_numpy(d_src).pin_memory()
    dr = pin_ref.to(dev).unsqueeze(0).expand(S, -1, -1).contiguous()
    ds = pin_src.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    host_mask = torch.empty(S, H, W, dtype=torch.bool).pin_memory()
    host_depth = torch.empty(S, H, W, dtype=torch.float32).pin_memory()

    def step_resident():
        return ops.geo_consistency(dr, ds, cams)

    def step_e2e():
        r = pin_ref.to(dev, non_blocking=True).unsqueeze(0).expand(S, -1, -1).contiguous()
        o = ops.geo_consistency(r, pin_src.to(dev, non_blocking=True), cams)
        host_mask.copy_(o[0], non_blocking=True)
        host_depth.copy_(o[1], non_blocking=True)

    l0 = ssmvs_b200._lib.launches
    with ClockSampler(local) as clk:
        ms_total = _timed_steps(step_resident, args.steps, args.warmup, dev, flush, parallel)
        launches = ssmvs_b200._lib.launches - l0
        ms_e2e = _timed_steps(step_e2e, args.steps, 1, dev, flush, parallel)
    clocks = clk.summary()
    peaks = load_peaks()
    px = S * H * W
    geo_bytes = px * (4 + 4 + 1 + 5 * 4)          # both depth maps read, the mask and five float maps written
    t_geo = ms_total / args.steps
    small = torch.rand(16, 128, 160, device=dev) * 500 + 400
    t_up = _event_ms(lambda: ops.upsample_nearest(small, (H, W), flip_rows=True), dev, flush)
    up_bytes = 16 * (128 * 160 + H * W) * 4
    V = 10
    nd = fz.constant_normals(torch.from_numpy(np.stack([d_ref] * V)).to(dev))
    fcams = torch.stack([fz.camera_block(k, synth.extrinsics(v % 5)) for v in range(V)]).to(dev)
    t_fuse = _event_ms(lambda: fz.fuse_view(nd, fcams, 0, None, 0.25, 0.52, 3), dev, flush)
    if rank == 0:
        line = {"metric": "pixels/sec through the geometric-consistency filter (reference view vs 10 source views, 1200x1600)",
                "value": world * args.steps * px / (ms_total * 1e-3), "unit": "pixels/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": t_geo, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64 per-pixel algebra, f32 maps (the reference's NumPy promotions)", "data": "synthetic",
                "config": {"workload": "inference output side: reproject_with_depth + check_geometric_consistency, 10 pairs of 1200x1600 maps per step "
                                       "(jdacs/eval_dense.py:177-232)", "pairs_per_step": S,
                           "l2": "256 MiB buffer written between timed steps (outside the per-step event pairs)", "launch": "python (one C-ABI launch per step)"},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": world * args.steps * px / (ms_e2e * 1e-3), "unit": "pixels/s", "h2d_bytes_per_step": world * (1 + S) * H * W * 4,
                        "d2h_bytes_per_step": world * S * H * W * 5, "ms_per_step": ms_e2e / args.steps,
                        "how": "every step uploads the reference and the 10 source depth maps from pinned memory and reads masks + re-projected depths back"},
                "roofline": {"kernel": "geo_consistency_kernel", "bound": "hbm", "achieved": geo_bytes / (t_geo * 1e-3) / 1e9, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": geo_bytes / (t_geo * 1e-3) / 1e9 / peaks["hbm_gbs"], "traffic": None, "ms": t_geo,
                             "peak_source": peaks["source"], "algorithmic_bytes": geo_bytes,
                             "note": "~120 fp64 FMAs per pixel and a scattered 4-tap gather: fp64-issue bound below the HBM roofline"},
                "upsample_nearest": {"maps": 16, "ms": t_up, "algorithmic_bytes": up_bytes, "gbs": up_bytes / t_up / 1e6, "frac_of_hbm_peak": up_bytes / t_up / 1e6 / peaks["hbm_gbs"]},
                "fusibile": {"views": V, "ms_per_reference_view": t_fuse}}
        if world == 1 and not args.no_cpu_baseline:
            spec = importlib.util.spec_from_file_location("output_side_oracle", os.path.join(ROOT, "oracle", "output_side.py"))
            side = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(side)
            t0 = time.perf_counter()
            for s_ in range(3):
                side.check_geometric_consistency(d_ref, k, ex[0], d_src[s_], k, ex[1 + s_ % 4])
            sec = (time.perf_counter() - t0) / 3
            line["cpu_baseline"] = {"value": H * W / sec, "unit": "pixels/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "3 of the 10 pairs: NumPy restatement of reproject_with_depth + check_geometric_consistency (fp64 BLAS matmuls, "
                                              "remap restated in NumPy) on the host"}
        print(json.dumps(line))
    parallel.barrier()


def _reserve_stdout_for_the_json_line():
    """The contract is ONE JSON line on stdout.  Native libraries write there too (NCCL prints its version banner on fd 1 when the
    box sets NCCL_DEBUG=VERSION), so fd 1 is pointed at stderr and Python's sys.stdout keeps the original descriptor for the line."""
    try:
        sys.stdout.flush()
        real = os.dup(1)
        os.dup2(2, 1)
        sys.stdout = os.fdopen(real, "w", buffering=1)
    except OSError:          # no usable stderr / stdout descriptor: leave the streams as they are
        pass


def main():
    _reserve_stdout_for_the_json_line()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=["mvsnet", "cvp", "train", "output"], default="mvsnet",
                    help="mvsnet = the headline metric (BASELINE configs[1]); cvp = configs[2]; train = configs[3] (one JDACS training batch)")
    ap.add_argument("--dtype", choices=["fp16", "bf16", "fp32"], default=None, help="storage dtype of the inference workloads (default fp16; cvp: bf16)")
    ap.add_argument("--train-dtype", choices=["fp16", "bf16", "fp32"], default="bf16")
    ap.add_argument("--batch", type=int, default=None, help="items per GPU per step (independent MVS problems; the reference ran 1 per 11 GB GPU; measured 1 / 2 / 4 / 8 items: 4.4 / 5.3 / 6.3 / 6.4 G samples/s)")
    ap.add_argument("--lanes", type=int, default=1, help="item groups captured on separate streams inside the CUDA graph (GraphedForward)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the end-to-end parity check against the fp32 oracle (rank 0, 1 GPU only)")
    ap.add_argument("--img-dtype", choices=["same", "fp32"], default="same",
                    help="storage of the images handed to MVSNet.forward: 'same' = the volume dtype (16-bit images give bit-identical "
                         "results to fp32 images on the 16-bit path, whose first layer rounds its input anyway, at half the upload)")
    ap.add_argument("--knob", action="append", default=[], metavar="NAME=VALUE", help="kernel-variant knob for tuning runs (mvs_set_knob); not used by the driver")
    ap.add_argument("--no-graph", action="store_true", help="launch every kernel from Python instead of replaying a CUDA graph")
    ap.add_argument("--ncu-range", action="store_true", help="cudaProfilerStart/Stop around the timed steps (ncu --profile-from-start off)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    args.dtype_given, args.batch_given = args.dtype is not None, args.batch is not None
    args.dtype = args.dtype or "fp16"
    args.batch = args.batch if args.batch is not None else 8

    import ssmvs_b200
    from ssmvs_b200 import ops, parallel, synth

    if args.impl == "reference":
        run_reference(args, int(os.environ.get("RANK", "0")))
        return
    if args.workload == "train":
        return run_train(args)
    if args.workload == "cvp":
        return run_cvp(args)
    if args.workload == "output":
        return run_output(args)

    rank, world, local = parallel.init_from_env("nccl")
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (the plane-sweep path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_bound = parallel.bind_to_gpu_numa(local) if world > 1 else False    # before any pinned allocation
    ssmvs_b200._lib.bind()
    for kv in args.knob:
        ssmvs_b200._lib.set_knob(kv.split("=")[0], int(kv.split("=")[1]))
    dtype = {"fp16": torch.float16, "bf16": torch.bfloat16, "fp32": torch.float32}[args.dtype]
    elem = 4 if dtype == torch.float32 else 2
    model = make_model(dtype).to(dev)
    PB = args.batch
    host = synth.mvsnet_inputs(PB, VIEWS, HEIGHT, WIDTH, NDEPTH, seed=rank)
    if args.img_dtype == "same" and dtype != torch.float32:
        host["imgs"] = host["imgs"].to(dtype)
    pinned = {k: host[k].pin_memory() for k in ("imgs", "proj_matrices", "depth_values")}
    res = {k: v.to(dev) for k, v in pinned.items()}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    out_host = {"depth": torch.empty(PB, HF, WF).pin_memory(), "conf": torch.empty(PB, HF, WF).pin_memory()}
    stream = torch.cuda.current_stream(dev)

    graphed = None
    if not args.no_graph:
        from ssmvs_b200.graph import GraphedForward
        graphed = GraphedForward(lambda i, pm, dv: model(i, pm, dv), [res["imgs"], res["proj_matrices"], res["depth_values"]], lanes=args.lanes)

    def step_resident():
        if graphed is not None:
            return graphed(*graphed.static_in)      # inputs already resident in the graph's static buffers
        return model(res["imgs"], res["proj_matrices"], res["depth_values"])

    pipe = None
    if graphed is not None:
        from ssmvs_b200.graph import StreamedForward
        pipe = StreamedForward(lambda i, pm, dv: model(i, pm, dv), [res["imgs"], res["proj_matrices"], res["depth_values"]],
                               ("depth", "photometric_confidence"))
    host_batch = [pinned["imgs"], pinned["proj_matrices"], pinned["depth_values"]]

    def e2e_loop(steps):
        """K steps through the public host-to-host pipeline: every step uploads its inputs from pinned host memory and reads
        depth + confidence back to the host; upload i+1 overlaps the kernels of step i.  One event pair around the K steps
        (no L2 flush needed: a step streams > 1 GB of fresh inputs and intermediates through the 126 MB L2)."""
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            a.record(stream)
            if pipe is not None:
                pipe.copy.wait_event(a)
                for _ in pipe.run(host_batch for _ in range(steps)):
                    pass
            else:
                for _ in range(steps):
                    d = {k: v.to(dev, non_blocking=True) for k, v in pinned.items()}
                    o = model(d["imgs"], d["proj_matrices"], d["depth_values"])
                    out_host["depth"].copy_(o["depth"], non_blocking=True)
                    out_host["conf"].copy_(o["photometric_confidence"], non_blocking=True)
            b.record(stream)
            torch.cuda.synchronize(dev)
        return a.elapsed_time(b)

    def timed(fn, steps, warmup):
        """per-step CUDA-event pairs on the launching stream, L2 flushed between steps (outside the pairs)."""
        with torch.no_grad():
            for _ in range(warmup):
                fn()
            torch.cuda.synchronize(dev)
            parallel.barrier()
            torch.cuda.synchronize(dev)
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
            l0 = ssmvs_b200._lib.launches
            t0 = time.perf_counter()
            if args.ncu_range and fn is step_resident:
                torch.cuda.profiler.start()
            for a, b in ev:
                flush.zero_()
                a.record(stream)
                fn()
                b.record(stream)
            torch.cuda.synchronize(dev)
            if args.ncu_range and fn is step_resident:
                torch.cuda.profiler.stop()
            parallel.barrier()
            wall = time.perf_counter() - t0
            launches = ssmvs_b200._lib.launches - l0
        ms = sum(a.elapsed_time(b) for a, b in ev)
        return parallel.max_over_ranks(ms, dev), wall, launches

    with ClockSampler(local) as clk:
        ms_total, wall, launches = timed(step_resident, args.steps, args.warmup)
        e2e_loop(3)
        torch.cuda.synchronize(dev)
        parallel.barrier()
        ms_e2e = parallel.max_over_ranks(e2e_loop(args.steps), dev)
    clocks = clk.summary()
    if args.ncu_range:
        return

    # ---- per-stage timing for the roofline (same inputs, each stage bracketed by events, L2 flushed before each)
    def stage_times(reps=5):
        t = {"features": [], "warp_var": [], "reg3d": [], "softargmin": [], "conv0": []}
        mk = lambda: torch.cuda.Event(enable_timing=True)
        with torch.no_grad():
            for _ in range(reps):
                e = [mk() for _ in range(10)]
                imgs = res["imgs"]
                flush.zero_(); e[0].record(stream)
                if dtype != torch.float32:   # the path MVSNet.forward takes in eval mode: FeatureNet on the tcgen05 kernel
                    maps = model.feature.forward_maps(imgs, dtype)
                else:
                    f = model.feature(imgs.transpose(0, 1).reshape(VIEWS * PB, 3, HEIGHT, WIDTH))
                    maps = ops.pack_c8_padded(f, dtype)
                    maps = maps.view(VIEWS, PB, *maps.shape[1:])
                e[1].record(stream)
                rt = ops.compose_proj(res["proj_matrices"])
                flush.zero_(); e[2].record(stream)
                var = ops.warp_variance_maps(maps, rt, res["depth_values"], dtype)
                e[3].record(stream)
                flush.zero_(); e[4].record(stream)
                model.cost_regularization.act_dtype = dtype
                reg = model.cost_regularization(var)
                e[5].record(stream)
                flush.zero_(); e[6].record(stream)
                ops.soft_argmin(reg, res["depth_values"])
                e[7].record(stream)
                # the single largest launch of the stack: conv0 (32 -> 8 channels at full D x H x W), timed alone
                flush.zero_(); e[8].record(stream)
                model.cost_regularization.conv0(var)
                e[9].record(stream)
                torch.cuda.synchronize(dev)
                for k, (i, j) in zip(t, ((0, 1), (2, 3), (4, 5), (6, 7), (8, 9))):
                    t[k].append(e[i].elapsed_time(e[j]))
        return {k: statistics.median(v) for k, v in t.items()}

    stages = stage_times()
    peaks = load_peaks()
    try:   # DRAM bytes per launch from the committed `ncu --set full` captures (profiles/), per item
        with open(os.path.join(ROOT, "profiles", "r02_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        traffic = {}
    traffic = {k: (v * PB if isinstance(v, (int, float)) else v) for k, v in traffic.items()}   # captures are for one item
    wv_gbs = PB * warp_var_bytes(elem) / (stages["warp_var"] * 1e-3) / 1e9
    reg_tfs = PB * REG_FLOPS / (stages["reg3d"] * 1e-3) / 1e12
    conv0_tfs = PB * CONV0_FLOPS / (stages["conv0"] * 1e-3) / 1e12
    roof_wv = {"kernel": "warp_var_fwd_tma_kernel (fused warp + bilinear gather + variance, TMA-staged source windows)", "bound": "hbm", "achieved": wv_gbs, "peak": peaks["hbm_gbs"],
               "unit": "GB/s", "frac": wv_gbs / peaks["hbm_gbs"], "traffic": traffic.get("warp_var_fwd"), "ms": stages["warp_var"],
               "peak_source": peaks["source"], "algorithmic_bytes": PB * warp_var_bytes(elem),
               "note": "bound on chip three ways at once: 4 taps x 64 B per (voxel, source) = 31.5 M wavefronts per item through the 128 B/clk L1/shared pipe (ncu: 73 %), FMA pipe 69 % (HFMA2 issues at half rate), issue 67 %; traffic = DRAM reads + SM->L2 write sectors of one ncu --set full capture: 1.008 x the algorithmic bytes (the volume is written exactly once)"}
    roof_conv0 = {"kernel": "conv3d_tc_kernel (conv0: 32->8, 3x3x3, full D x H x W)", "bound": "tensor", "achieved": conv0_tfs,
                  "peak": peaks["bf16_tflops"], "unit": "TFLOP/s", "frac": conv0_tfs / peaks["bf16_tflops"],
                  "traffic": traffic.get("conv0"), "ms": stages["conv0"], "peak_source": peaks["source"],
                  "algorithmic_flops": PB * CONV0_FLOPS,
                  "hbm": {"achieved": PB * CONV0_BYTES(elem) / (stages["conv0"] * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                          "frac": PB * CONV0_BYTES(elem) / (stages["conv0"] * 1e-3) / 1e9 / peaks["hbm_gbs"], "algorithmic_bytes": PB * CONV0_BYTES(elem)},
                  "note": "thin layer (32 -> 8 channels): 6.9 kFLOP per 80 B of volume, below the ridge; each tcgen05.mma pays a fixed 128 x 16 A-operand fetch for N <= 96 columns (tools/umma_bench.cu: 32 + N/4 cycles), so both the HBM and the tensor fractions are reported"}
    roof_reg = {"kernel": "CostRegNet conv stack (12 launches)", "bound": "tensor", "achieved": reg_tfs, "peak": peaks["bf16_tflops_sustained"],
                "unit": "TFLOP/s", "frac": reg_tfs / peaks["bf16_tflops_sustained"], "traffic": None, "ms": stages["reg3d"],
                "peak_source": peaks["source"]}
    dominant = roof_conv0 if stages["conv0"] >= stages["warp_var"] else roof_wv

    if rank == 0:
        items = world * args.steps * PB
        value = items * SAMPLES_PER_ITEM / (ms_total * 1e-3)
        e2e_value = items * SAMPLES_PER_ITEM / (ms_e2e * 1e-3)
        h2d = world * sum(v.numel() * v.element_size() for v in pinned.values())      # whole job: every rank uploads its own items
        d2h = world * sum(v.numel() * v.element_size() for v in out_host.values())
        line = {"metric": METRIC, "value": value, "unit": "depth-samples/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": {"fp16": "f16", "bf16": "bf16", "fp32": "f32"}[args.dtype] + " storage, f32 accumulate",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "global_batch": world * PB, "per_gpu_batch": PB, "parallelism": "dp%d (items sharded, no collective)" % world,
                           "l2": "256 MiB buffer written between timed steps (outside the per-step event pairs)",
                           "launch": "python" if graphed is None else "cuda-graph replay (%d C-ABI launches per step, %d stream lane(s))" % (graphed.launches_per_replay, graphed.lanes),
                           "wall_s_incl_flush": wall, "numa_bound": numa_bound},
                "clocks": clocks, "gpu_launches": launches,
                "e2e": {"value": e2e_value, "unit": "depth-samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps,
                        "how": "StreamedForward.run: pinned host batch -> H2D (copy stream, one step ahead, straight into the static inputs of one of two captured graphs) -> graph replay -> D2H of depth + confidence, every step; one event pair around the K steps, max over ranks"},
                "roofline": dominant, "roofline_warp_var": roof_wv, "roofline_conv0": roof_conv0, "roofline_reg3d": roof_reg,
                "stage_ms": stages}
        if world == 1 and not args.no_parity:
            line["parity"] = parity_check(dev, dtype, args.dtype)
        if world == 1 and not args.no_cpu_baseline:
            times, threads = cpu_forward_timer(3, 1)
            line["cpu_baseline"] = {"value": SAMPLES_PER_ITEM * len(times) / sum(times), "unit": "depth-samples/s", "cores": threads,
                                    "kind": "port", "sample": "3 full forward passes of 1 item after 1 warm-up (oracle port of the reference's torch-CPU path, fp32; best of the probed thread counts, host has %d cores)" % (os.cpu_count() or 1)}
            try:
                gt = gpu_library_timer(dev)
                line["gpu_library_baseline"] = {"value": SAMPLES_PER_ITEM * len(gt) / sum(gt), "unit": "depth-samples/s", "kind": "port",
                                                "sample": "3 forward passes of 1 item after 1 warm-up: the reference's op sequence (oracle port) run by PyTorch on this GPU (ATen grid_sample, cuDNN conv3d, fp32 with torch's default TF32 convolutions)"}
            except Exception as exc:   # informational only
                line["gpu_library_baseline"] = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
        print(json.dumps(line))
    parallel.barrier()


if __name__ == "__main__":
    main()
