"""Run the two training-only kernels alone at BASELINE configs[3] size (1 item) for ncu: conv0's weight gradient
(conv3d_wgrad_mma_kernel) and the plane-sweep backward (warp_var_bwd16_kernel)."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ssmvs_b200
from ssmvs_b200 import ops, synth
ssmvs_b200._lib.bind()
dev = torch.device("cuda:0")
dt = torch.bfloat16
var = torch.randn(1, 4, 192, 128, 160, 8, device=dev).to(dt)
gz = torch.randn(1, 1, 192, 128, 160, 8, device=dev).to(dt)
w0 = torch.zeros(8, 32, 3, 3, 3, device=dev)
inp = synth.mvsnet_inputs(1, 5, 512, 640, 192, seed=0)
feats = [torch.randn(1, 32, 128, 160, device=dev).requires_grad_(True) for _ in range(5)]
rt = ops.compose_proj(inp["proj_matrices"].to(dev))
v = ops.warp_variance(feats[0], feats[1:], rt, inp["depth_values"].to(dev), dt)
gv = torch.randn_like(v)
for _ in range(2):
    ops._wgrad_mma(var, gz, w0, 8, 1, False, 8)
    torch.autograd.grad(v, feats, gv, retain_graph=True)
torch.cuda.synchronize()
print("done")
