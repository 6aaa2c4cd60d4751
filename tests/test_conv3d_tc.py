"""tcgen05 implicit-GEMM convolution (algo = 2) against the fp32 SIMT path (algo = 1) and ATen on identical inputs.

The inputs are rounded to the storage dtype first, so the only differences are accumulation order (fp32 in both)
and the rounding of the output; tolerances are a few ulp of the storage type."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # cin, cout, stride, transposed, (B, D, H, W) of the INPUT
    (32, 8, 1, False, (1, 8, 16, 30)),      # exactly one tile
    (32, 8, 1, False, (2, 5, 19, 37)),      # ragged everything, batch 2
    (16, 16, 1, False, (1, 12, 24, 40)),
    (32, 32, 1, False, (1, 9, 20, 33)),
    (64, 64, 1, False, (1, 6, 12, 31)),     # streamed weight tiles, nM = 2
    (64, 32, 1, True, (1, 6, 10, 20)),      # stride-1 ConvTranspose3d (CVP conv5)
    (16, 1, 1, False, (1, 10, 17, 45)),     # single-channel output (prob0 of CVP)
    (32, 64, 1, False, (1, 4, 9, 31)),
    (8, 1, 1, False, (1, 9, 18, 35)),       # Cin = 8: paired taps (prob of MVSNet)
    (8, 8, 1, False, (2, 6, 17, 31)),
    (8, 16, 2, False, (1, 12, 20, 64)),     # stride 2, Cin = 8 (conv1)
    (16, 32, 2, False, (2, 10, 18, 62)),    # stride 2 (conv3)
    (32, 64, 2, False, (1, 8, 16, 36)),     # stride 2 (conv5)
    (64, 32, 2, True, (1, 5, 9, 17)),       # transposed stride 2 (conv7)
    (32, 16, 2, True, (2, 6, 13, 33)),      # conv9
    (16, 8, 2, True, (1, 10, 21, 47)),      # conv11
]


CASES += [
    (8, 8, 1, True, (1, 7, 18, 33)),        # stride-1 ConvTranspose3d with <= 8 output channels (kd-folded, flipped taps)
    (32, 8, 1, False, (1, 40, 64, 90)),     # several depth segments and (h, w) tiles per persistent CTA
    (16, 16, 1, False, (2, 24, 40, 70)),
    (16, 8, 2, True, (1, 20, 24, 50)),
]


@pytest.mark.parametrize("dtype,force_nm,kdfold", [(torch.float16, 0, 1), (torch.bfloat16, 0, 1), (torch.float16, 4, 1), (torch.float16, 2, 1),
                                                   (torch.float16, 0, 0), (torch.float16, 4, 0)])
@pytest.mark.parametrize("cin,cout,stride,tr,shape", CASES)
def test_tc_matches_simt_and_aten(gpu, knob, cin, cout, stride, tr, shape, dtype, force_nm, kdfold):
    from ssmvs_b200 import ops
    if not kdfold and not (stride == 1 and cout <= 8):
        pytest.skip("MVS_TC_KDFOLD only changes stride-1 layers with <= 8 output channels")
    knob("tc_kdfold", kdfold)       # library test knob: fold the kd taps into N as well (default on)
    if force_nm:
        knob("tc_nm", force_nm)     # library test knob: M-tiles per CTA (default: by volume size)
    else:
        knob("tc_nm", -1)
    torch.manual_seed(cin * 100 + cout + stride)
    b, d, h, w = shape
    dev = gpu.device
    x = torch.randn(b, cin, d, h, w, device=dev)
    wt = 0.1 * (torch.randn(cin, cout, 3, 3, 3, device=dev) if tr else torch.randn(cout, cin, 3, 3, 3, device=dev))
    wt = wt.to(dtype).float()                                   # weights representable in the storage dtype
    g = ops.pack_conv3d_weight(wt, tr)
    x8 = ops.pack_c8(x, dtype)
    xr = ops.unpack_c8(x8)
    want = F.conv_transpose3d(xr, wt, None, stride, 1, stride - 1) if tr else F.conv3d(xr, wt, None, stride, 1)
    scale = (torch.rand(cout, device=dev) + 0.5) if cout > 1 else None
    shift = torch.randn(cout, device=dev)
    skip = ops.pack_c8(torch.randn_like(want), dtype) if cout > 1 else None
    y_tc = ops.conv3d_raw(x8, g, cout, stride, tr, scale, shift, skip, relu=cout > 1, algo=2)
    y_ref = ops.conv3d_raw(x8, g, cout, stride, tr, scale, shift, skip, relu=cout > 1, algo=1)
    torch.cuda.synchronize()
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    err = (y_tc.float() - y_ref.float()).abs().max().item()
    ref = y_ref.float().abs().max().item()
    assert err <= tol * ref, (err, ref)
    # and against ATen on the same rounded inputs (independent of our SIMT kernel)
    if cout > 1:
        want = F.relu(want * scale.view(1, -1, 1, 1, 1) + shift.view(1, -1, 1, 1, 1)) + ops.unpack_c8(skip)
        got = ops.unpack_c8(y_tc)
    else:
        want = (want + shift.view(1, 1, 1, 1, 1)).squeeze(1)
        got = y_tc
    assert rel_err(got, want) < 2 * tol


def test_tc_is_selected_automatically(gpu):
    """algo = 0 must pick the tensor-core kernel whenever it applies (and give the same bits as algo = 2)."""
    from ssmvs_b200 import ops
    torch.manual_seed(0)
    x8 = ops.pack_c8(torch.randn(1, 32, 8, 16, 32, device=gpu.device), torch.float16)
    g = ops.pack_conv3d_weight(0.1 * torch.randn(8, 32, 3, 3, 3, device=gpu.device), False)
    assert torch.equal(ops.conv3d_raw(x8, g, 8, algo=0), ops.conv3d_raw(x8, g, 8, algo=2))


# ---------------------------------------------------------------------------------------------- 2-D feature layers
CASES_2D = [
    # cin (real), cout, ksize, stride, (M, H, W), out_padded, relu
    (3, 8, 3, 1, (2, 16, 30), False, True),       # FeatureNet conv0: 3 input channels padded to 8, exactly one tile
    (3, 8, 3, 1, (3, 37, 70), False, True),       # ragged
    (8, 8, 3, 1, (5, 64, 96), False, True),       # conv1
    (8, 16, 5, 2, (3, 44, 68), False, True),      # conv2: 5x5 stride 2, Cin = 8 (paired taps)
    (16, 16, 3, 1, (2, 33, 47), False, True),     # conv3 / conv4
    (16, 32, 5, 2, (3, 40, 92), False, True),     # conv5
    (32, 32, 3, 1, (2, 21, 35), False, True),     # conv6
    (32, 32, 3, 1, (5, 32, 40), True, False),     # feature: bias only, zero-bordered image-major output
    (32, 32, 3, 1, (10, 128, 160), True, False),  # headline feature-map size, two items x 5 views
    (3, 8, 3, 1, (5, 128, 160), False, True),
    # FeaturePyramid of CVP-MVSNet: bias + LeakyReLU(0.1); 64 output channels run the unfolded 9-entry program
    (3, 64, 3, 1, (2, 24, 40), False, 0.1),
    (64, 64, 3, 1, (3, 37, 61), False, 0.1),
    (64, 32, 3, 1, (2, 32, 30), False, 0.1),
    (32, 16, 3, 1, (2, 19, 33), False, 0.1),
    (16, 16, 3, 1, (5, 64, 80), True, 0.1),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cin,cout,k,stride,shape,padded,relu", CASES_2D)
def test_tc_conv2d_matches_aten(gpu, cin, cout, k, stride, shape, padded, relu, dtype):
    """The 2-D mode of the tcgen05 kernel (image stack as a volume, no taps across images) against F.conv2d on inputs and
    weights rounded to the storage type."""
    from ssmvs_b200 import ops
    torch.manual_seed(cin * 7 + cout + k)
    m, h, w = shape
    dev = gpu.device
    cinp = (cin + 7) // 8 * 8
    x = torch.randn(m, cin, h, w, device=dev).to(dtype).float()
    wt = (0.2 * torch.randn(cout, cin, k, k, device=dev)).to(dtype).float()
    leaky = isinstance(relu, float)
    scale = (torch.rand(cout, device=dev) + 0.5) if (relu and not leaky) else None
    shift = torch.randn(cout, device=dev)
    xs = torch.zeros(cinp // 8, m, h, w, 8, device=dev, dtype=dtype)           # C8 image stack [Cin/8][M][H][W][8]
    xp = torch.zeros(m, cinp, h, w, device=dev)
    xp[:, :cin] = x
    xs.copy_(xp.view(m, cinp // 8, 8, h, w).permute(1, 0, 3, 4, 2))
    y = ops.conv2d_raw(xs, ops.pack_conv2d_weight(wt), cout, k, stride, scale, shift, relu, out_padded=padded)
    torch.cuda.synchronize()
    want = F.conv2d(x, wt, None, stride, k // 2)
    want = want * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1) if scale is not None else want + shift.view(1, -1, 1, 1)
    if leaky:
        want = F.leaky_relu(want, relu)
    elif relu:
        want = F.relu(want)
    ho, wo = want.shape[2:]
    if padded:
        assert y.shape == (m, cout // 8, ho + 3, wo + 2, 8)
        yf = y.float()
        got = yf[:, :, 1:ho + 1, 1:wo + 1].permute(0, 1, 4, 2, 3).reshape(m, cout, ho, wo)
        border = yf.clone()
        border[:, :, 1:ho + 1, 1:wo + 1] = 0
        assert torch.count_nonzero(border) == 0
    else:
        assert y.shape == (cout // 8, m, ho, wo, 8)
        got = y.float().permute(1, 0, 4, 2, 3).reshape(m, cout, ho, wo)
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    assert rel_err(got, want) < 2 * tol


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 6e-2)])
def test_feature_net_on_tcgen05(gpu, oracle, dtype, tol):
    """FeatureNet.forward_maps (8 launches of the tcgen05 kernel) against the oracle's fp32 FeatureNet (jdacs/models/mvsnet.py:17-34)."""
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.models.mvsnet import MVSNet
    torch.manual_seed(0)
    model = MVSNet(refine=False).eval()
    with torch.no_grad():
        for mod in model.feature.modules():          # non-trivial BatchNorm statistics
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.7, 1.3); mod.bias.normal_(0, 0.2)
    inp = synth.mvsnet_inputs(2, 3, 64, 96, 8, seed=3)
    sd = {k[len("feature."):]: v.detach().clone() for k, v in model.state_dict().items() if k.startswith("feature.")}
    with torch.no_grad():
        want = torch.stack([oracle.feature_net(inp["imgs"][:, v], sd) for v in range(3)], 0)      # [N,B,32,16,24]
        maps = model.to(gpu.device).feature.forward_maps(inp["imgs"].to(gpu.device), dtype)
    assert maps.shape == (3, 2, 4, 16 + 3, 24 + 2, 8)
    got = maps.float()[:, :, :, 1:17, 1:25].permute(0, 1, 2, 5, 3, 4).reshape(3, 2, 32, 16, 24).cpu()
    assert rel_err(got, want) < tol


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("batch,views,h,w,img16", [(2, 3, 64, 96, False), (1, 2, 52, 76, True), (1, 5, 128, 160, False), (3, 1, 36, 34, False)])
def test_fused_featnet_front_matches_the_layered_path_and_aten(gpu, dtype, batch, views, h, w, img16):
    """mvs_featnet_front (conv0 + conv1 + conv2 of FeatureNet in one kernel: mma.sync stages over shared-memory tiles with
    recomputed halos) against (a) the three tcgen05 launches it replaces -- same operands, same fp32 accumulation, so equal up
    to the summation order: a few 16-bit ulps -- and (b) ATen in fp32 on the rounded image / weights.  Ragged tiles (extents not
    multiples of 32 x 16), one and several images per item, fp32 and 16-bit images."""
    from ssmvs_b200 import ops
    from ssmvs_b200.jdacs.models.mvsnet import FeatureNet
    torch.manual_seed(h + w)
    net = FeatureNet().eval()
    with torch.no_grad():
        for mod in net.modules():
            if isinstance(mod, torch.nn.BatchNorm2d):
                mod.running_mean.normal_(0, 0.2); mod.running_var.uniform_(0.5, 1.5); mod.weight.uniform_(0.7, 1.3); mod.bias.normal_(0, 0.2)
            if isinstance(mod, torch.nn.Conv2d):
                mod.weight.copy_(mod.weight.to(dtype).float())
    net = net.to(gpu.device)
    imgs = torch.randn(batch, views, 3, h, w, device=gpu.device).to(dtype)
    layers = []
    for blk in (net.conv0, net.conv1, net.conv2):
        scale, shift = ops.fold_bn(blk.bn)
        layers.append((ops.pack_conv2d_weight(blk.conv.weight), blk.conv.out_channels, blk.conv.kernel_size[0], blk.conv.stride[0], scale, shift))
    frag, aff = ops.featnet_front_pack(net.conv0.conv.weight, net.conv1.conv.weight, net.conv2.conv.weight, [(l[4], l[5]) for l in layers], dtype)
    got = ops.featnet_front(imgs if img16 else imgs.float(), frag, aff, dtype)
    x = ops.pack_images_c8(imgs.float(), dtype)
    for g, cout, k, stride, scale, shift in layers:
        x = ops.conv2d_raw(x, g, cout, k, stride, scale, shift, True)
    assert got.shape == x.shape == (2, views * batch, h // 2, w // 2, 8)
    ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
    assert rel_err(got, x) < 4 * ulp, rel_err(got, x)
    # (b) ATen fp32 on the same rounded inputs, rounding every layer's output like the kernels do
    y = imgs.float().transpose(0, 1).reshape(views * batch, 3, h, w)
    with torch.no_grad():
        for blk in (net.conv0, net.conv1, net.conv2):
            y = F.relu(blk.bn(blk.conv(y))).to(dtype).float()
    want = y.view(views * batch, 2, 8, h // 2, w // 2).permute(1, 0, 3, 4, 2)
    assert rel_err(got, want) < 8 * ulp, rel_err(got, want)


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
def test_feature_net_fused_front_switch_is_transparent(gpu, dtype):
    """FeatureNet.forward_maps with and without the fused front: the same zero-bordered maps up to 16-bit rounding."""
    from ssmvs_b200 import synth
    from ssmvs_b200.jdacs.models.mvsnet import FeatureNet
    torch.manual_seed(0)
    net = FeatureNet().to(gpu.device).eval()
    synth.randomise_bn(net, 3)
    imgs = synth.mvsnet_inputs(2, 3, 64, 96, 8, seed=3)["imgs"].to(gpu.device)
    with torch.no_grad():
        a = net.forward_maps(imgs, dtype)
        net.fused_front = False
        b = net.forward_maps(imgs, dtype)
    assert a.shape == b.shape and rel_err(a, b) < (4e-3 if dtype == torch.float16 else 3e-2)


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 1e-2), (torch.bfloat16, 8e-2)])
def test_feature_pyramid_on_tcgen05(gpu, oracle, dtype, tol):
    """FeaturePyramid.forward_maps (9 launches per level) against the oracle's fp32 pyramid (jdacs-ms/models/network.py:16-41)."""
    from types import SimpleNamespace
    from ssmvs_b200.jdacs_ms.models.network import CVPMVSNet
    torch.manual_seed(1)
    model = CVPMVSNet(SimpleNamespace(nsrc=2, nscale=3, mode="test")).eval()
    imgs = torch.randn(2, 3, 3, 64, 96)
    sd = {k[len("featurePyramid."):]: v.detach().clone() for k, v in model.state_dict().items() if k.startswith("featurePyramid.")}
    with torch.no_grad():
        want = [oracle.feature_pyramid(imgs[:, v], sd, 3) for v in range(3)]           # [view][level] -> [B,16,h,w]
        maps = model.to(gpu.device).featurePyramid.forward_maps(imgs.to(gpu.device), 3, dtype)
    for level, m in enumerate(maps):
        h, w = 64 >> level, 96 >> level
        assert m.shape == (3, 2, 2, h + 3, w + 2, 8)
        got = m.float()[:, :, :, 1:h + 1, 1:w + 1].permute(0, 1, 2, 5, 3, 4).reshape(3, 2, 16, h, w).cpu()
        ref = torch.stack([want[v][level] for v in range(3)], 0)
        assert rel_err(got, ref) < tol, (level, rel_err(got, ref))
