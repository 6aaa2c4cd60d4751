"""FeatureNet.forward_maps at the headline size (8 items x 5 views of 512x640, fp16) under kernel-variant knobs:
`python tools/featnet_time.py knob=value ...` (e.g. tc_kwfold_max=48)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssmvs_b200
from ssmvs_b200.jdacs.models.mvsnet import FeatureNet

ssmvs_b200._lib.bind()
for a in sys.argv[1:]:
    k, v = a.split("=")
    if k != "fused":
        ssmvs_b200._lib.set_knob(k, int(v))
dev = torch.device("cuda:0")
net = FeatureNet().to(dev).eval()
net.fused_front = "fused=0" not in sys.argv
imgs = torch.randn(8, 5, 3, 512, 640, device=dev).half()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with torch.no_grad():
    for _ in range(3):
        net.forward_maps(imgs, torch.float16)
    ts = []
    for _ in range(7):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); net.forward_maps(imgs, torch.float16); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
print(" ".join(sys.argv[1:]) or "defaults", ": %.3f ms" % sorted(ts)[len(ts) // 2])
