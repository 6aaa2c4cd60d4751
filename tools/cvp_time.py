"""BASELINE.json configs[2]: CVP-MVSNet forward, 3 pyramid levels, 1 + 4 views of 512x640, bf16 volumes: time per item and per stage."""
import os
import sys
from types import SimpleNamespace

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ssmvs_b200  # noqa: E402
from ssmvs_b200 import synth  # noqa: E402
from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet  # noqa: E402

dev = torch.device("cuda:0")
ssmvs_b200._lib.bind()
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1
dt = {"bf16": torch.bfloat16, "fp16": torch.float16, "fp32": torch.float32}[sys.argv[2] if len(sys.argv) > 2 else "bf16"]
torch.manual_seed(0)
model = CVPMVSNet(SimpleNamespace(nsrc=4, nscale=3, mode="test"), volume_dtype=dt).eval().to(dev)
inp = {k: v.to(dev) for k, v in synth.cvp_inputs(B, 4, 512, 640, seed=0).items()}
args = [inp[k] for k in ("ref_img", "src_imgs", "ref_in", "src_in", "ref_ex", "src_ex", "depth_min", "depth_max")]
samples = B * (48 * 128 * 160 + 8 * 256 * 320 + 8 * 512 * 640)
with torch.no_grad():
    for _ in range(3):
        model(*args)
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record(); model(*args); b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)[2]
    print("CVP-MVSNet forward B=%d %s: %.3f ms per step, %.1f M depth-samples/s (4,259,840 per item)" % (B, dt, ms, samples / ms / 1e3))
    # stage split: feature pyramid alone (the path the eval forward takes: tcgen05 kernel for 16-bit volumes)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    imgs = torch.cat((inp["ref_img"].unsqueeze(1), inp["src_imgs"]), 1)
    a.record()
    if dt != torch.float32:
        model.featurePyramid.forward_maps(imgs, 3, dt)
    else:
        for i in range(5):
            model.featurePyramid(imgs[:, i], 3)
    b.record()
    torch.cuda.synchronize()
    print("  FeaturePyramid, 5 views x 3 levels: %.3f ms" % a.elapsed_time(b))
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        model(*args)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
