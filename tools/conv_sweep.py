"""GPU micro-benchmark of the tcgen05 convolution over channel counts / tiles (CUDA-event timed, L2 flushed)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssmvs_b200
from ssmvs_b200 import ops

ssmvs_b200._lib.bind()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1e3)
    return sorted(ts)[len(ts) // 2]


def run(cin, cout, stride, tr, shape, nm=0):
    if nm:
        ssmvs_b200._lib.set_knob("tc_nm", nm)
    else:
        ssmvs_b200._lib.set_knob("tc_nm", -1)
    b, d, h, w = shape
    x8 = ops.pack_c8(torch.randn(b, cin, d, h, w, device=dev), torch.float16)
    wt = 0.1 * (torch.randn(cin, cout, 3, 3, 3, device=dev) if tr else torch.randn(cout, cin, 3, 3, 3, device=dev))
    g = ops.pack_conv3d_weight(wt, tr)
    us = timeit(lambda: ops.conv3d_raw(x8, g, cout, stride, tr, relu=cout > 1, algo=2))
    vox = d * h * w * (8 if (tr and stride == 2) else 1) // (8 if (stride == 2 and not tr) else 1)
    macs = vox * 27 * cin * cout if not (tr and stride == 2) else d * h * w * 27 * cin * cout
    print("cin %2d cout %2d s%d %s %-18s nM=%d : %8.1f us  %6.1f TMAC/s" % (cin, cout, stride, "T" if tr else " ", shape, nm, us, macs / us / 1e6))


full = (1, 192, 128, 160)
if len(sys.argv) > 1 and sys.argv[1] == "prof":   # short list for ncu captures
    run(32, 8, 1, False, full)
    run(32, 64, 1, False, full)
    run(16, 8, 2, True, (1, 96, 64, 80))
    sys.exit(0)
for cout in (8, 16, 32, 64):
    run(32, cout, 1, False, full)
for nm in (1, 2, 4):
    run(32, 8, 1, False, full, nm)
run(16, 16, 1, False, full)
run(16, 8, 1, False, full)
run(64, 64, 1, False, (1, 96, 64, 80))
run(8, 1, 1, False, full)
run(8, 8, 1, False, full)
run(16, 8, 2, True, (1, 96, 64, 80))
run(8, 16, 2, False, full)
for shape in ((1, 24, 16, 20), (1, 48, 32, 40)):
    run(64, 64, 1, False, shape)
    run(32, 32, 1, False, shape)
