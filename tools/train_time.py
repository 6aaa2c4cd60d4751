"""BASELINE.json configs[3]: one JDACS training step (photometric loss, N = 5, 512x640, D = 192): forward / backward / Adam times and the
top kernels.  `python tools/train_time.py [fp32|bf16|fp16]` selects the train dtype (default: the model default, bf16)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ssmvs_b200  # noqa: E402
from ssmvs_b200 import synth  # noqa: E402
from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss  # noqa: E402
from ssmvs_b200.jdacs.models.mvsnet import MVSNet  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = "bench" in sys.argv       # jdacs/train.py:35 sets it
BATCH = 2 if "b2" in sys.argv else 1
ssmvs_b200._lib.bind()
torch.manual_seed(0)
tdt = {"fp32": torch.float32, "bf16": torch.bfloat16, "fp16": torch.float16}.get(sys.argv[1] if len(sys.argv) > 1 else "", None)
model = MVSNet(refine=False, train_dtype=tdt).to(dev).train()
opt = torch.optim.Adam(model.parameters(), lr=1e-3)
crit = UnSupLoss()
inp = {k: v.to(dev) for k, v in synth.mvsnet_inputs(BATCH, 5, 512, 640, 192, seed=2).items()}
ev = lambda: torch.cuda.Event(enable_timing=True)
rows = []
for it in range(4):
    e = [ev() for _ in range(5)]
    e[0].record()
    out = model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])
    e[1].record()
    loss = crit(inp["imgs"], inp["cams"], out["depth"])
    e[2].record()
    opt.zero_grad()
    loss.backward()
    e[3].record()
    opt.step()
    e[4].record()
    torch.cuda.synchronize()
    rows.append([e[i].elapsed_time(e[i + 1]) for i in range(4)])
r = rows[-1]
print("train step (%d item(s), cudnn.benchmark=%s)" % (BATCH, torch.backends.cudnn.benchmark), end=" "); print("train step: forward %.1f ms, loss %.1f ms, backward %.1f ms, Adam %.1f ms, total %.1f ms; peak memory %.1f GB" % (
    r[0], r[1], r[2], r[3], sum(r), torch.cuda.max_memory_allocated() / 2**30))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = model(inp["imgs"], inp["proj_matrices"], inp["depth_values"])
    loss = crit(inp["imgs"], inp["cams"], out["depth"])
    opt.zero_grad(); loss.backward(); opt.step()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=40, max_name_column_width=90))
