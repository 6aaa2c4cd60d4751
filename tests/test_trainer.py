"""trainer.TrainStep (one JDACS batch: train_sample + train_sample_aug, jdacs/train.py:189-291) on the host-emulation build:
the device-side mask box and the gather-free augmentation loss equal the reference's formulation, and a step trains."""
import pytest
import torch
import torch.nn.functional as F



def test_mask_box_matches_slicing():
    from ssmvs_b200.trainer import draw_mask_box, mask_reference_view
    g = torch.Generator().manual_seed(3)
    imgs = torch.randn(2, 4, 3, 24, 36, generator=g)
    for _ in range(5):
        box = draw_mask_box(24, 36, g)
        x, y = int(box[0]), int(box[1])
        want_mask = torch.ones(2, 3, 24, 36)
        want_mask[:, :, y:y + 8, x:x + 12] = 0          # models/augmentations.py:107-124
        out, mask = mask_reference_view(imgs, box)
        assert torch.equal(mask.expand(2, 3, 24, 36), want_mask)
        assert torch.equal(out[:, 0], imgs[:, 0] * want_mask) and torch.equal(out[:, 1:], imgs[:, 1:])


def test_train_step_runs_and_matches_gather_form(emu):
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    from ssmvs_b200.trainer import TrainStep, mask_reference_view
    torch.manual_seed(0)
    model = MVSNet(refine=False, train_dtype=torch.float32)
    step = TrainStep(model, UnSupLoss(), lr=1e-3)
    inp = synth.mvsnet_inputs(1, 4, 32, 64, 8, seed=1)
    inp["imgs_aug"] = inp["imgs"] + 0.05 * torch.randn(inp["imgs"].shape, generator=torch.Generator().manual_seed(2))
    before = [p.detach().clone() for p in model.parameters()]
    out = step(inp["imgs"], inp["imgs_aug"], inp["cams"], inp["proj_matrices"], inp["depth_values"])
    assert torch.isfinite(out["loss"]) and torch.isfinite(out["augment_loss"]) and out["loss"] > 0
    assert sum(int(not torch.equal(a, b)) for a, b in zip(before, model.parameters())) > 10     # both optimiser steps moved weights
    # the augmentation loss without the boolean gather == the reference's depth_aug[mask] form (train.py:262-266)
    g = torch.Generator().manual_seed(5)
    da, de = torch.rand(1, 8, 16, generator=g) * 50 + 400, torch.rand(1, 8, 16, generator=g) * 50 + 400
    _, fmask = mask_reference_view(inp["imgs"], torch.tensor([10, 5]))
    fm = F.interpolate(fmask.float(), scale_factor=0.25)[:, 0] > 0.5
    want = F.smooth_l1_loss(da[fm], de[fm])
    got = (F.smooth_l1_loss(da, de, reduction="none") * fm.float()).sum() / fm.float().sum()
    assert abs(float(want) - float(got)) < 1e-5 * float(want)


@pytest.mark.gpu
def test_graphed_train_step_equals_the_eager_step(gpu):
    """GraphedTrainStep (the whole batch -- two forward / fused loss / backward / Adam passes -- captured in one CUDA graph) against
    the eager TrainStep from the same initial weights, on the same batches and mask boxes (the warm-up steps the graphed step takes
    before the capture are mirrored): same losses and parameters up to the run-to-run spread of the atomic gradient reductions,
    over several replays (the graph must pick up new inputs and new boxes)."""
    import copy
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    from ssmvs_b200.trainer import GraphedTrainStep, TrainStep, draw_mask_box
    torch.manual_seed(0)
    dev = gpu.device
    model_a = MVSNet(refine=False, train_dtype=torch.bfloat16).to(dev)
    model_b = copy.deepcopy(model_a)
    init = [p.detach().clone() for p in model_a.parameters()]
    batches = []
    for seed in (1, 2, 3):
        b = {k: v.to(dev) for k, v in synth.mvsnet_inputs(1, 4, 64, 96, 8, seed=seed).items()}
        b["imgs_aug"] = b["imgs"] + 0.05 * torch.randn_like(b["imgs"])
        batches.append(b)
    keys = ("imgs", "imgs_aug", "cams", "proj_matrices", "depth_values")
    eager = TrainStep(model_a, UnSupLoss(), lr=1e-4)
    graphed_step = TrainStep(model_b, UnSupLoss(), lr=1e-4)
    eager.gen.manual_seed(7)
    graphed_step.gen.manual_seed(7)
    warm = 2
    graphed = GraphedTrainStep(graphed_step, batches[0], warmup=warm)        # draws one box, takes `warm` eager steps with it, captures
    model_a.train()
    box0 = draw_mask_box(64, 96, eager.gen).to(dev)
    for _ in range(warm):
        eager.run(*[batches[0][k] for k in keys], box0)
    for b in batches:
        la = eager(*[b[k] for k in keys])
        lb = graphed(*[b[k] for k in keys])
        assert torch.isfinite(lb["loss"]) and abs(float(la["loss"]) - float(lb["loss"])) < 3e-2 * abs(float(la["loss"])) + 1e-3
        assert abs(float(la["augment_loss"]) - float(lb["augment_loss"])) < 0.2 * abs(float(la["augment_loss"])) + 1e-3
    # both trained, and to (nearly) the same place: distance between the two runs against the distance travelled, over all
    # parameters (Adam turns the noise of a near-zero gradient into full-size steps, so single entries may differ)
    moved = sum(float((p.detach() - p0).double().pow(2).sum()) for p, p0 in zip(model_b.parameters(), init)) ** 0.5
    apart = sum(float((pa.detach() - pb.detach()).double().pow(2).sum()) for pa, pb in zip(model_a.parameters(), model_b.parameters())) ** 0.5
    assert moved > 1e-3 and apart < 0.5 * moved, (moved, apart)


def test_feature_layer_embedding_is_plain_tensor_algebra():
    """The host-side re-expressions behind FeatureNet.forward_train_tc need no kernel to check: a 3x3 Conv2d is the kd = 1 slice of
    the embedded 3x3x3 weight (input channels padded to 8), a 5x5 stride-2 pad-2 Conv2d is the 3x3 stride-1 convolution of the
    parity planes (space_to_depth_c8) with the re-ordered weight -- both differentiable w.r.t. the original weight."""
    from ssmvs_b200 import ops
    torch.manual_seed(0)
    x = torch.randn(2, 8, 12, 20)
    w5 = torch.randn(16, 8, 5, 5, requires_grad=True)
    want = F.conv2d(x, w5, None, 2, 2)
    x8 = x.permute(0, 2, 3, 1).reshape(2, 1, 1, 12, 20, 8).permute(2, 1, 0, 3, 4, 5).contiguous()      # C8 image volume [1, 1, M, H, W, 8]
    xs = ops.space_to_depth_c8(x8)
    assert xs.shape == (1, 4, 2, 6, 10, 8)
    w3 = ops.embed_conv2d_weight(w5, 2)
    assert w3.shape == (16, 32, 3, 3, 3) and float(w3[:, :, 0].abs().max()) == 0 and float(w3[:, :, 2].abs().max()) == 0
    got = F.conv2d(xs[0].permute(1, 0, 4, 2, 3).reshape(2, 32, 6, 10), w3[:, :, 1], None, 1, 1)
    assert (got - want).abs().max() < 1e-4
    got.sum().backward()                                                  # the embedding is differentiable: same weight gradient
    g_embed = w5.grad.clone()
    w5.grad = None
    want.sum().backward()
    assert (g_embed - w5.grad).abs().max() < 1e-3
    w33 = torch.randn(8, 3, 3, 3)
    e = ops.embed_conv2d_weight(w33, 1)
    assert e.shape == (8, 8, 3, 3, 3) and torch.equal(e[:, :3, 1], w33) and float(e[:, 3:].abs().max()) == 0
    with pytest.raises(ValueError):
        ops.embed_conv2d_weight(torch.randn(8, 8, 7, 7), 1)
    # several views at once: the parity split keeps the leading (view) axis
    xv = torch.randn(3, 2, 4, 6, 8, 8)
    assert ops.space_to_depth_c8(xv).shape == (3, 8, 4, 3, 4, 8)
    assert torch.equal(ops.space_to_depth_c8(xv)[1], ops.space_to_depth_c8(xv[1:2])[0])


def test_augmentation_mirror_matches_reference_semantics():
    """models/augmentations.py mirror: the same mask as slicing at the drawn corner, the same loss as the boolean gather."""
    import numpy as np
    from ssmvs_b200.jdacs.models.augmentations import aug_loss, random_image_mask
    g = torch.Generator().manual_seed(0)
    img = torch.randn(2, 3, 24, 36, generator=g)
    np.random.seed(3)
    x, y = np.random.randint(0, 36 - 12), np.random.randint(0, 24 - 8)      # the draws random_image_mask is about to make
    np.random.seed(3)
    out, mask = random_image_mask(img, (8, 12))
    want = torch.ones_like(img)
    want[:, :, y:y + 8, x:x + 12] = 0.0
    assert torch.equal(mask, want) and torch.equal(out, img * want)
    out2, mask2 = random_image_mask(img, (8, 12), box=torch.tensor([x, y]))
    assert torch.equal(mask2, want) and torch.equal(out2, out)
    same, none = random_image_mask(img, (24, 36))
    assert none is None and same is img
    a, b = torch.rand(2, 6, 9, generator=g) * 50 + 400, torch.rand(2, 6, 9, generator=g) * 50 + 400
    m = torch.rand(2, 6, 9, generator=g)
    assert abs(float(aug_loss(a, b, m)) - float(F.smooth_l1_loss(a[m > 0.5], b[m > 0.5]))) < 1e-5


def test_self_supervised_training_trajectory_matches_the_oracle(emu, oracle):
    """The reference's train_sample loop (jdacs/train.py:189-203: model.train(), forward, UnSupLoss on the estimated depth,
    backward, Adam step) run twice in lockstep from the same weights on the same batch: once through the product modules (fp32
    arithmetic, host build of the kernels), once through the oracle's functional restatement of the same modules.  Six optimiser
    steps: the loss and its three terms must agree at every step -- forward values, batch-statistics BatchNorm, every gradient and
    the parameter updates they drive all enter -- and the loss must go down."""
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    torch.manual_seed(0)
    model = MVSNet(refine=False, volume_dtype=torch.float32, train_dtype=torch.float32)
    params = {k: v.detach().clone().requires_grad_(v.dtype.is_floating_point and k in dict(model.named_parameters()))
              for k, v in model.state_dict().items()}
    start = {k: v.detach().clone() for k, v in model.named_parameters()}
    criterion = UnSupLoss()
    opt_a = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))
    opt_b = torch.optim.Adam([params[k] for k, _ in model.named_parameters()], lr=1e-3, betas=(0.9, 0.999))
    s = synth.mvsnet_inputs(1, 5, 64, 96, 8, seed=3)
    got, want = [], []
    for _ in range(6):
        model.train()
        opt_a.zero_grad()
        out = model(s["imgs"], s["proj_matrices"], s["depth_values"])
        loss = criterion(s["imgs"], s["cams"], out["depth"])
        loss.backward()
        opt_a.step()
        got.append([float(loss.detach()), float(criterion.reconstr_loss), float(criterion.ssim_loss), float(criterion.smooth_loss)])
        opt_b.zero_grad()
        ref = oracle.mvsnet_forward(s["imgs"], s["proj_matrices"], s["depth_values"], params, training=True)
        terms = oracle.unsup_loss(s["imgs"], s["cams"], ref["depth"], True, 0.18)
        terms["total"].backward()
        opt_b.step()
        want.append([float(terms[k].detach()) for k in ("total", "reconstr", "ssim", "smooth")])
    got, want = torch.tensor(got), torch.tensor(want)
    assert torch.isfinite(got).all()
    # Adam divides by the running gradient magnitude, so last-bit differences in near-zero gradients become full-size steps and the
    # two runs drift apart slowly (measured after 6 steps: total 8e-4, colour 2e-5, SSIM 4e-5, smoothness -- the smallest and most
    # curvature-sensitive term -- 4.5e-3); a wrong gradient anywhere shows up at the first update, two orders above these bounds
    drift = ((got - want).abs() / want.abs()).max(0).values
    assert drift[0] < 3e-3 and drift[1] < 5e-4 and drift[2] < 5e-4 and drift[3] < 1.5e-2, (drift, got, want)
    assert ((got[1] - want[1]).abs() / want[1].abs()).max().item() < 2e-4            # after the first update
    assert ((got[0] - want[0]).abs() / want[0].abs()).max().item() < 1e-5           # before any update: plain forward parity
    assert got[-1, 0] < got[0, 0] - 0.2
    # (parameters are not compared: at this size the deep layers see 6 voxels, most of their gradients are cancellation residue,
    # and Adam turns residue into +-lr steps -- prob.bias, whose true gradient is exactly zero under the softmax, wanders by 30 %)
    assert max(float((p.detach() - start[k]).abs().max()) for k, p in model.named_parameters()) > 3e-3      # and it did train


def test_cvp_self_supervised_trajectory_matches_the_oracle(emu, oracle):
    """The same lockstep check for JDACS-MS (jdacs-ms/train.py:200-258): CVP-MVSNet forward (2 levels, per-pixel hypotheses at the
    fine one), every level's depth resized to the image (nearest), UnSupLoss per level, summed, backward, Adam -- four steps
    against the oracle's cvp_forward / unsup_loss from the same weights."""
    from types import SimpleNamespace
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs_ms.losses.unsup_loss import UnSupLoss
    from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet
    torch.manual_seed(1)
    nsrc, nscale, h, w = 3, 2, 32, 48
    model = CVPMVSNet(SimpleNamespace(nsrc=nsrc, nscale=nscale, mode="train"), volume_dtype=torch.float32, train_dtype=torch.float32)
    names = [k for k, _ in model.named_parameters()]
    params = {k: v.detach().clone().requires_grad_(k in names) for k, v in model.state_dict().items()}
    criterion = UnSupLoss()
    opt_a = torch.optim.Adam(model.parameters(), lr=1e-3, betas=(0.9, 0.999))
    opt_b = torch.optim.Adam([params[k] for k in names], lr=1e-3, betas=(0.9, 0.999))
    c = synth.cvp_inputs(1, nsrc, h, w, seed=5)
    keys = ("ref_img", "src_imgs", "ref_in", "src_in", "ref_ex", "src_ex", "depth_min", "depth_max")
    imgs = torch.cat([c["ref_img"].unsqueeze(1), c["src_imgs"]], 1)
    cams = torch.zeros(1, nsrc + 1, 2, 4, 4)
    cams[:, 0, 0], cams[:, 1:, 0] = c["ref_ex"], c["src_ex"]
    cams[:, 0, 1, :3, :3], cams[:, 1:, 1, :3, :3] = c["ref_in"], c["src_in"]

    def total(ests, loss_fn):
        return sum(loss_fn(F.interpolate(d.unsqueeze(1), size=[h, w]).squeeze(1)) for d in ests)

    got, want = [], []
    for _ in range(4):
        model.train()
        opt_a.zero_grad()
        loss = total(model(*[c[k] for k in keys])["depth_est_list"], lambda d: criterion(imgs, cams, d))
        loss.backward()
        opt_a.step()
        got.append(float(loss.detach()))
        opt_b.zero_grad()
        ref = total(oracle.cvp_forward(c, params, nscale, training=True)["depth_est_list"],
                    lambda d: oracle.unsup_loss(imgs, cams, d, False, 0.05)["total"])
        ref.backward()
        opt_b.step()
        want.append(float(ref.detach()))
    got, want = torch.tensor(got), torch.tensor(want)
    assert torch.isfinite(got).all() and abs(got[0] - want[0]) / want[0] < 1e-5
    assert ((got - want).abs() / want).max().item() < 3e-3, (got, want)
    assert (got - got[0]).abs().max() > 0.05          # the weights moved the loss (on random images four steps need not lower it:
    #                                                   the oracle's run goes the same way, which is the point of the comparison)
