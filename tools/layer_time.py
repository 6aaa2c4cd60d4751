"""Time the full-resolution U-Net layers of MVSNet (conv0, conv11 with its skip, prob) and the stride-2 / transposed mid layers
alone, 8 items, fp16, CUDA events with an L2 flush: `python tools/layer_time.py [path/to/variant/libmvs_b200.so]` for A/B runs
of kernel variants built with other compile-time switches."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssmvs_b200
from ssmvs_b200 import ops

ssmvs_b200._lib.bind(next((a for a in sys.argv[1:] if a.endswith(".so")), None))
for a in sys.argv[1:]:
    if "=" in a:
        ssmvs_b200._lib.set_knob(a.split("=")[0], int(a.split("=")[1]))
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
B = 8


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


def layer(name, cin, cout, stride, tr, shape, skip):
    d, h, w = shape
    x8 = torch.randn(B, cin // 8, d, h, w, 8, device=dev).half()
    wt = 0.1 * (torch.randn(cin, cout, 3, 3, 3, device=dev) if tr else torch.randn(cout, cin, 3, 3, 3, device=dev))
    g = ops.pack_conv3d_weight(wt, tr)
    sc, sh = torch.ones(cout, device=dev), torch.zeros(cout, device=dev)
    sk = None
    if skip:
        m = 2 if (tr and stride == 2) else 1
        sk = torch.randn(B, cout // 8, d * m, h * m, w * m, 8, device=dev).half()
    cache = {}
    us = timeit(lambda: ops.conv3d_raw(x8, g, cout, stride, tr, sc if cout > 1 else None, sh, sk, relu=cout > 1, algo=2, tile_cache=cache))
    print("%-8s %2d -> %2d s%d %s %-16s skip=%d : %8.1f us" % (name, cin, cout, stride, "T" if tr else " ", shape, int(skip), us), flush=True)


layer("conv0", 32, 8, 1, False, (192, 128, 160), False)
layer("conv4", 32, 32, 1, False, (48, 32, 40), False)
layer("conv6", 64, 64, 1, False, (24, 16, 20), False)
layer("conv7", 64, 32, 2, True, (24, 16, 20), True)
layer("conv1", 8, 16, 2, False, (192, 128, 160), False)
layer("conv2", 16, 16, 1, False, (96, 64, 80), False)
layer("conv3", 16, 32, 2, False, (96, 64, 80), False)
layer("conv9", 32, 16, 2, True, (48, 32, 40), True)
layer("conv11", 16, 8, 2, True, (96, 64, 80), True)
layer("conv11ns", 16, 8, 2, True, (96, 64, 80), False)
layer("prob", 8, 1, 1, False, (192, 128, 160), False)
