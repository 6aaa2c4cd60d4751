#!/usr/bin/env python
"""Same-box LIBRARY bars per stage of the MVSNet forward (BASELINE configs[1], 1 item): what PyTorch's own kernels (ATen
grid_sample / elementwise, cuDNN convolutions) take for the stage our kernels replace, next to our kernels.

  * reference layout and precision: fp32 NCDHW / NCHW, torch defaults (cuDNN may use TF32)
  * the strongest library configuration: fp16 channels_last(_3d), BatchNorm folded into the weights (inference)

Prints one JSON document; CUDA-event timed, L2 flushed between repetitions, median of 5."""
import json
import os
import statistics
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import ssmvs_b200  # noqa: E402
from ssmvs_b200 import ops, synth  # noqa: E402
from ssmvs_b200.jdacs.models.mvsnet import MVSNet  # noqa: E402

ssmvs_b200._lib.bind()
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
V, H, W, D = 5, 512, 640, 192


def timeit(fn, reps=5):
    with torch.no_grad():
        fn(); fn(); torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); fn(); b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
    return statistics.median(ts)


torch.manual_seed(0)
model = MVSNet(refine=False)
synth.randomise_bn(model, 5)
model = model.to(dev).eval()
inp = {k: v.to(dev) for k, v in synth.mvsnet_inputs(1, V, H, W, D, seed=0).items()}
doc = {"gpu": torch.cuda.get_device_name(0), "workload": "MVSNet forward N=5 512x640 D=192, 1 item", "unit": "ms", "stages": {}}

# ---------------------------------------------------------------- FeatureNet
x = inp["imgs"].transpose(0, 1).reshape(V, 3, H, W).contiguous()
st = {"library_fp32_nchw": timeit(lambda: model.feature(x)),
      "library_fp16_channels_last_folded": timeit(lambda: model.feature.forward_folded(x, torch.float16)),
      "ours_fp16_tcgen05": timeit(lambda: model.feature.forward_maps(inp["imgs"], torch.float16))}
doc["stages"]["feature_net"] = st

# ---------------------------------------------------------------- cost volume (warp + variance)
with torch.no_grad():
    feats = model.feature(x)
f = [feats[v:v + 1].contiguous() for v in range(V)]
proj, dv = inp["proj_matrices"], inp["depth_values"]


def lib_variance(half):
    fe = [t.half() for t in f] if half else f
    b, c, h, w = fe[0].shape
    ref = fe[0].unsqueeze(2).repeat(1, 1, D, 1, 1)
    s1, s2 = ref, ref ** 2
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32, device=dev), torch.arange(w, dtype=torch.float32, device=dev), indexing="ij")
    pix = torch.stack((xs.reshape(-1), ys.reshape(-1), torch.ones(h * w, device=dev)))
    for s in range(1, V):
        rel = proj[:, s] @ torch.inverse(proj[:, 0])
        pts = (rel[:, :3, :3] @ pix.unsqueeze(0)).unsqueeze(2) * dv.view(b, 1, D, 1) + rel[:, :3, 3].view(b, 3, 1, 1)
        xy = pts[:, :2] / pts[:, 2:3]
        grid = torch.stack((xy[:, 0] / ((w - 1) / 2) - 1, xy[:, 1] / ((h - 1) / 2) - 1), dim=3).view(b, D * h, w, 2)
        wv = F.grid_sample(fe[s], grid.to(fe[s].dtype), mode="bilinear", padding_mode="zeros", align_corners=False).view(b, c, D, h, w)
        s1 += wv
        s2 += wv.pow_(2)
    return s2.div_(V).sub_(s1.div_(V).pow_(2))


rt = ops.compose_proj(proj)
maps16 = ops.pack_c8_padded(feats, torch.float16).view(V, 1, 4, H // 4 + 3, W // 4 + 2, 8)
doc["stages"]["warp_variance"] = {"library_fp32 (grid_sample + in-place elementwise, the reference's eval branch)": timeit(lambda: lib_variance(False)),
                                  "library_fp16 (same ops on half tensors)": timeit(lambda: lib_variance(True)),
                                  "ours_fp16_fused": timeit(lambda: ops.warp_variance_maps(maps16, rt, dv, torch.float16))}

# ---------------------------------------------------------------- CostRegNet
reg = model.cost_regularization
sd = {k: v.detach() for k, v in reg.state_dict().items()}
with torch.no_grad():
    var = lib_variance(False)


def fold(w, pre, transposed):
    s = sd[pre + "weight"] * torch.rsqrt(sd[pre + "running_var"] + 1e-5)
    sh = sd[pre + "bias"] - sd[pre + "running_mean"] * s
    return (w * (s.view(1, -1, 1, 1, 1) if transposed else s.view(-1, 1, 1, 1, 1))), sh


def make_stack(dtype, cl):
    fmt = torch.channels_last_3d if cl else torch.contiguous_format
    L = {}
    for n in ("conv0", "conv1", "conv2", "conv3", "conv4", "conv5", "conv6"):
        w, b = fold(sd[n + ".conv.weight"], n + ".bn.", False)
        L[n] = (w.to(dtype).contiguous(memory_format=fmt), b.to(dtype))
    for n in ("conv7", "conv9", "conv11"):
        w, b = fold(sd[n + ".0.weight"], n + ".1.", True)
        L[n] = (w.to(dtype).contiguous(memory_format=fmt), b.to(dtype))
    L["prob"] = (sd["prob.weight"].to(dtype).contiguous(memory_format=fmt), sd["prob.bias"].to(dtype))
    x0 = var.to(dtype).contiguous(memory_format=fmt)

    def run():
        c = lambda t, n, s: F.relu_(F.conv3d(t, L[n][0], L[n][1], s, 1))
        d = lambda t, n: F.relu_(F.conv_transpose3d(t, L[n][0], L[n][1], 2, 1, 1))
        c0 = c(x0, "conv0", 1); c2 = c(c(c0, "conv1", 2), "conv2", 1); c4 = c(c(c2, "conv3", 2), "conv4", 1)
        y = c(c(c4, "conv5", 2), "conv6", 1)
        y = c4 + d(y, "conv7"); y = c2 + d(y, "conv9"); y = c0 + d(y, "conv11")
        return F.conv3d(y, L["prob"][0], L["prob"][1], 1, 1)
    return run, x0, L


def ref_stack():   # the reference's own modules: conv3d + batch_norm (eval) + relu, fp32 NCDHW, torch defaults
    oracle_spec = __import__("importlib.util").util.spec_from_file_location("o", os.path.join(ROOT, "oracle", "planesweep.py"))
    o = __import__("importlib.util").util.module_from_spec(oracle_spec); oracle_spec.loader.exec_module(o)
    return lambda: o.cost_reg_mvsnet(var, sd, False)


st = {}
st["library_fp32_ncdhw_reference_modules (conv3d + batch_norm + relu, torch defaults)"] = timeit(ref_stack())
run32, _, _ = make_stack(torch.float32, False)
st["library_fp32_ncdhw_bn_folded"] = timeit(run32)
for name, dt in (("fp16", torch.float16), ("bf16", torch.bfloat16)):
    try:
        run_cl, x0, L = make_stack(dt, True)
        st["library_%s_channels_last_3d_bn_folded" % name] = timeit(run_cl)
        st["library_%s_channels_last_3d conv0 alone" % name] = timeit(lambda: F.relu_(F.conv3d(x0, L["conv0"][0], L["conv0"][1], 1, 1)))
    except Exception as exc:
        st["library_%s_channels_last_3d_bn_folded" % name] = "failed: %s" % exc
    var8 = ops.pack_c8(var, dt)
    reg.act_dtype = dt
    st["ours_%s_tcgen05" % name] = timeit(lambda: reg(var8))
    st["ours_%s_tcgen05 conv0 alone" % name] = timeit(lambda: reg.conv0(var8))
doc["stages"]["cost_reg_net"] = st

# ---------------------------------------------------------------- softmax + regression + confidence
cost = torch.randn(1, D, H // 4, W // 4, device=dev)


def lib_tail():
    p = F.softmax(cost, 1)
    depth = torch.sum(p * dv.view(1, D, 1, 1), 1)
    win = 4 * F.avg_pool3d(F.pad(p.unsqueeze(1), (0, 0, 0, 0, 1, 2)), (4, 1, 1), stride=1).squeeze(1)
    idx = torch.sum(p * torch.arange(D, device=dev, dtype=torch.float32).view(1, D, 1, 1), 1).long()
    return depth, torch.gather(win, 1, idx.unsqueeze(1)).squeeze(1)


doc["stages"]["softargmin"] = {"library_fp32": timeit(lib_tail), "ours_fused": timeit(lambda: ops.soft_argmin(cost, dv))}
print(json.dumps(doc, indent=1))
