"""Timeline of one CTA of the tcgen05 convolution (debug build with -DMVS_TC_TRACE): prints, per depth step, when the
producer issued the slot, when the issuer passed its waits / finished issuing, and when the epilogue started / finished."""
import ctypes
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
pkg = os.path.join(ROOT, "self-supervised-mvs_b200")
lib = os.environ.get("MVS_TRACE_LIB", os.path.join(pkg, "libmvs_b200_trace.so"))
srcs = ["core.cu", "warp.cu", "softargmin.cu", "conv3d_simt.cu", "conv3d_tc.cu", "invwarp.cu", "loss.cu", "output.cu", "fusion.cu", "featnet_front.cu", "train.cu"]
if "--build" in sys.argv:
    subprocess.check_call(["/usr/local/cuda/bin/nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-Xcompiler", "-fPIC",
                           "-DMVS_TC_TRACE"] + [a for a in sys.argv if a.startswith("-D")] + ["-shared", "-o", lib] + [os.path.join(pkg, "csrc", s) for s in srcs] + ["-lcudart", "-lcuda", "--expt-relaxed-constexpr"])
    sys.exit(0)
import torch
import ssmvs_b200
from ssmvs_b200 import ops
ssmvs_b200._lib.bind(lib)
dev = torch.device("cuda:0")
if len(sys.argv) > 1 and sys.argv[1] == "2d":      # 2d cin cout ksize stride : a FeatureNet layer over 5 images of 512x640 / stride history
    cin, cout, k, stride = (int(a) for a in sys.argv[2:6])
    h, w = (512, 640) if cin <= 8 and k == 3 else ((256, 320) if cin <= 16 and not (cin == 16 and k == 5 and False) else (128, 160))
    if len(sys.argv) > 7:
        h, w = int(sys.argv[6]), int(sys.argv[7])
    xs = torch.randn(max(cin, 8) // 8, 5, h, w, 8, device=dev).half()
    g = ops.pack_conv2d_weight(0.1 * torch.randn(cout, cin, k, k, device=dev))
    for _ in range(3):
        ops.conv2d_raw(xs, g, cout, k, stride, None, None, True)
else:
    cin, cout = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (32, 8)
    stride = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    tr = len(sys.argv) > 4 and sys.argv[4] == "T"
    shape = (1, cin, 96, 64, 80) if (tr and stride == 2) else (1, cin, 192, 128, 160)
    x8 = ops.pack_c8(torch.randn(*shape, device=dev), torch.float16)
    g = ops.pack_conv3d_weight(0.1 * (torch.randn(cin, cout, 3, 3, 3, device=dev) if tr else torch.randn(cout, cin, 3, 3, 3, device=dev)), tr)
    skip = None
    if tr and stride == 2 and "noskip" not in sys.argv:
        skip = ops.pack_c8(torch.randn(1, cout, 192, 128, 160, device=dev), torch.float16)
    for _ in range(3):
        ops.conv3d_raw(x8, g, cout, stride, tr, skip=skip, relu=cout > 1, algo=2)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * (10 * 1024))()
fn = ssmvs_b200._lib.lib().mvs_debug_tc_trace
fn.argtypes, fn.restype = [ctypes.c_void_p], ctypes.c_int
assert fn(buf) == 0
t = [[buf[r * 1024 + i] for i in range(1024)] for r in range(10)]
t0 = min(v for v in t[0][:8] if v)
print("step: producer_issue | issuer: at_wait, past_wait, issued | epilogue: start, done   (cycles since first TMA)")
for i in range(40):
    print("%3d: %8d | %8d %8d %8d | %8d %8d | mma_issued %8d" % (i, t[0][i] - t0, t[1][i] - t0, t[2][i] - t0, t[3][i] - t0, t[4][i] - t0, t[5][i] - t0, t[6][i] - t0))

print("epilogue of warp 6, first channel block per call: enter -> loads issued -> tmem ready -> exit (cycles)")
for i in range(4, 24):
    print("%3d: enter %8d  +setup %6d  +tmem %6d  +math/stores(all cb) %6d" % (i, t[8][i] - t0, t[6][i] - t[8][i], t[7][i] - t[6][i], t[9][i] - t[7][i]))
print("epilogue rows of warp 6 (per M-tile): skip loads issued -> TMEM data ready, and gap to the next M-tile")
for i in range(4, 24):
    print("%3d: issued %8d  tmem_ready +%6d  next_mtile +%6d" % (i, t[6][i] - t0, t[7][i] - t[6][i], t[6][i + 1] - t[7][i]))
