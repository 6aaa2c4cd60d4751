"""The end-to-end parity harness itself (tests/parity_util.py), on the CPU box: it must produce a peaky probability volume
(hazard H11) and the fp32 product path (host-emulation build of the kernel sources) must agree with the oracle on it, index
included.  The 16-bit comparisons at BASELINE sizes are tests/test_gpu_fullsize.py (-m gpu)."""
import torch

import parity_util as pu


def test_mvsnet_peaky_volume_and_fp32_parity(emu):
    model, inp, want, cond = pu.peaky_mvsnet("cpu", 3, 64, 96, 16, seed=1, target_peak=0.3)
    span = float(inp["depth_values"][0, -1] - inp["depth_values"][0, 0])
    assert cond["peak"] >= 0.3 and cond["depth_std"] >= 0.1 * span, cond      # not the uniform softmax of default weights
    r = pu.depth_parity(pu.product_mvsnet(model, inp, torch.float32), want)
    assert r["depth_rel_max"] < 1e-5 and r["conf_abs_max"] < 1e-4, r
    assert r["index_mismatch"] == r["index_mismatch_near_integer"], r
    # what 16-bit storage alone costs on this input: reported, and far above fp32 round-off (so the GPU test is not vacuous)
    ideal = pu.depth_parity(pu.ideal_storage_mvsnet(model, inp, torch.float16), want)
    assert 10 * r["depth_rel_max"] < ideal["depth_rel_max"] < 1e-2, ideal


def test_cvp_peaky_volume_and_fp32_parity(emu):
    model, inp, want, cond = pu.peaky_cvp("cpu", 3, 2, 64, 96, seed=3, target_peak=0.3)
    assert cond["peak"] >= 0.3 and cond["depth_std"] >= 10.0, cond
    r = pu.cvp_parity(pu.product_cvp(model, inp, torch.float32), want)
    assert r["level0"]["depth_rel_max"] < 1e-4 and r["level1"]["depth_rel_max"] < 1e-4 and r["conf_abs_max"] < 1e-3, r
