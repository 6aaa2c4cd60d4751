"""BASELINE.json configs[4]: cost-volume sweep D in {48,96,192,256} x image sizes up to 1600x1200 (N=5, C=32, fp16):
the fused warp+variance kernel (algorithmic GB/s vs the measured HBM peak) and conv0 of CostRegNet, the layer that consumes the
volume (TFLOP/s vs the measured tensor peak, and its own GB/s).  CUDA events, L2 flushed before every timed launch.
    python tools/cost_volume_sweep.py > profiles/r01_cost_volume_sweep.txt        (on a B200)"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ssmvs_b200  # noqa: E402
from ssmvs_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
ssmvs_b200._lib.bind()
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tens = peaks["hbm_gbs"], peaks["bf16_tflops"]
except Exception:
    hbm, tens = 6650.0, 1590.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("N=5 views, C=32, fp16 storage; HBM peak %.0f GB/s, tensor peak %.0f TFLOP/s (MEASURED_PEAKS.json)" % (hbm, tens))
print("%-11s %4s %9s | %8s %8s %6s | %8s %8s %6s %8s" % ("image", "D", "samples", "warp ms", "GB/s", "frac", "conv0 ms", "TFLOP/s", "frac", "GB/s"))
for (ih, iw) in ((512, 640), (864, 1152), (1184, 1600)):
    h, w = ih // 4, iw // 4
    for d in (48, 96, 192, 256):
        inp = synth.feature_inputs(1, 5, 32, h, w, d, seed=0)
        maps = ops.pack_c8_padded(inp["features"].flatten(0, 1).to(dev), torch.float16)
        maps = maps.view(5, 1, *maps.shape[1:])
        rt = ops.compose_proj(inp["proj_matrices"].to(dev))
        dv = inp["depth_values"].to(dev)
        g = ops.pack_conv3d_weight(0.05 * torch.randn(8, 32, 3, 3, 3, device=dev), False)
        cache = {}
        tw, tc = [], []
        for it in range(7):
            e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
            flush.zero_(); e[0].record()
            var = ops.warp_variance_maps(maps, rt, dv, torch.float16)
            e[1].record()
            flush.zero_(); e[2].record()
            y = ops.conv3d_raw(var, g, 8, relu=True, tile_cache=cache)
            e[3].record()
            torch.cuda.synchronize()
            if it >= 2:
                tw.append(e[0].elapsed_time(e[1])); tc.append(e[2].elapsed_time(e[3]))
            del var, y
        n = d * h * w
        wms, cms = statistics.median(tw), statistics.median(tc)
        wbytes = 2 * (32 * n + 5 * 32 * h * w) + 4 * d
        cflops, cbytes = 2 * 27 * 32 * 8 * n, 2 * 40 * n
        print("%-11s %4d %9d | %8.3f %8.0f %6.3f | %8.3f %8.1f %6.3f %8.0f" % ("%dx%d" % (iw, ih), d, n, wms, wbytes / wms / 1e6, wbytes / wms / 1e6 / hbm,
                                                                              cms, cflops / cms / 1e9, cflops / cms / 1e9 / tens, cbytes / cms / 1e6))
