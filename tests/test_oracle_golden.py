"""The oracle (oracle/planesweep.py) against the golden fixtures produced by the UNMODIFIED reference
(oracle/gen_golden.py, run in the build container where /root/reference exists).  CPU only."""
import torch

from conftest import rel_err, state_dict_of


def test_homo_warping(oracle, golden):
    g = golden("jdacs_warp")
    for restated in (False, True):
        out = oracle.homo_warping(g["src_fea"], g["src_proj"], g["ref_proj"], g["depth_values"], restated_sampler=restated)
        assert rel_err(out, g["warped"]) < 1e-6
    out = oracle.homo_warping(g["src_fea"], g["src_proj2"], g["ref_proj"], g["depth_far"], restated_sampler=True)
    assert torch.allclose(out, g["warped_far"], atol=1e-6)


def test_mvsnet_stages(oracle, golden):
    g = golden("jdacs_mvsnet")
    sd = state_dict_of(g)
    st = {}
    with torch.no_grad():
        out = oracle.mvsnet_forward(g["imgs"], g["proj_matrices"], g["depth_values"], sd, False, False, st)
    assert rel_err(st["features"], g["features"]) < 1e-6
    assert rel_err(st["variance"], g["variance"]) < 1e-6
    assert rel_err(st["cost_reg"], g["cost_reg"]) < 1e-5
    assert rel_err(out["depth"], g["depth"]) < 1e-6
    assert rel_err(out["photometric_confidence"], g["photometric_confidence"]) < 1e-5
    assert torch.equal(st["index"], g["index"])  # bit-exact expected-depth index


def test_mvsnet_training_gradients(oracle, golden):
    g = golden("jdacs_mvsnet")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "running" not in k) for k, v in state_dict_of(g).items()}
    out = oracle.mvsnet_forward(g["imgs"], g["proj_matrices"], g["depth_values"], sd, True)
    assert rel_err(out["depth"], g["train_depth"]) < 1e-5
    (out["depth"] * g["loss_weight"]).sum().backward()
    for k, v in g.items():
        if k.startswith("grad."):
            assert rel_err(sd[k[5:]].grad, v) < 1e-3, k


def test_ms_warp_and_hypotheses(oracle, golden):
    g = golden("ms_warp")
    hyp = oracle.sweeping_depth_hypos(g["depth_min"], g["depth_max"], 1)
    assert hyp.shape == g["sweep_hypos"].shape and rel_err(hyp, g["sweep_hypos"]) < 1e-6
    w = oracle.homo_warping_ms(g["src_fea"], g["ref_in_l1"], g["src_in_l1"][:, 0], g["ref_ex"], g["src_ex"][:, 0], g["sweep_hypos"])
    assert rel_err(w, g["warped"]) < 1e-6
    h = oracle.depth_hypos_refine(g["depth_up"], g["ref_in"], g["src_in"][:, 0], g["ref_ex"], g["src_ex"][:, 0])
    assert rel_err(h, g["refine_hypos"]) < 1e-6
    rp = oracle.compose_projection(g["ref_in"], g["ref_ex"])
    sp = [oracle.compose_projection(g["src_in"][:, i], g["src_ex"][:, i]) for i in range(2)]
    c = oracle.variance_volume(g["ref_fea"], [g["src_fea0"], g["src_fea1"]], rp, sp, g["refine_hypos"], True)
    assert rel_err(c, g["proj_cost"]) < 1e-6


def test_cvp_forward(oracle, golden):
    g = golden("ms_cvp")
    with torch.no_grad():
        out = oracle.cvp_forward(g, state_dict_of(g), 2)
    for i, d in enumerate(out["depth_est_list"]):
        assert rel_err(d, g["depth_est_list.%d" % i]) < 1e-5
    assert rel_err(out["prob_confidence"], g["prob_confidence"]) < 1e-4


def test_inverse_warping(oracle, golden):
    g = golden("jdacs_invwarp")
    d = g["depth"].clone().requires_grad_(True)
    w, m = oracle.inverse_warping(g["img"], g["cams"][:, 0], g["cams"][:, 2], d)
    (w * g["weight"]).sum().backward()
    assert rel_err(w, g["warped"]) < 1e-6 and torch.equal(m, g["mask"])
    assert rel_err(d.grad, g["grad_depth"]) < 1e-5
    w, m = oracle.inverse_warping(g["img"], g["cams"][:, 0], g["cams"][:, 1], g["depth_far"])
    assert rel_err(w, g["warped_far"]) < 1e-6 and torch.equal(m, g["mask_far"])


def test_unsup_loss(oracle, golden):
    for name, down, ws in (("jdacs_unsup_loss", True, 0.18), ("ms_unsup_loss", False, 0.05)):
        g = golden(name)
        d = g["depth"].clone().requires_grad_(True)
        out = oracle.unsup_loss(g["imgs"], g["cams"], d, down, ws)
        out["total"].backward()
        for k in ("total", "reconstr", "ssim", "smooth"):
            assert rel_err(out[k], g[k]) < 1e-6, (name, k)
        assert rel_err(d.grad, g["grad_depth"]) < 1e-5
