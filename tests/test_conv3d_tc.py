"""tcgen05 implicit-GEMM convolution (algo = 2) against the fp32 SIMT path (algo = 1) on identical inputs.

The inputs are rounded to the storage dtype first, so the only differences are accumulation order (fp32 in both)
and the rounding of the output; tolerances are a few ulp of the storage type."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err

pytestmark = pytest.mark.gpu

CASES = [
    # cin, cout, transposed, (B, D, H, W)
    (32, 8, False, (1, 8, 16, 30)),      # exactly one tile
    (32, 8, False, (2, 5, 19, 37)),      # ragged everything, batch 2
    (16, 16, False, (1, 12, 24, 40)),
    (32, 32, False, (1, 9, 20, 33)),
    (64, 64, False, (1, 6, 12, 31)),     # streamed weight tiles, nM = 2
    (64, 32, True, (1, 6, 10, 20)),      # stride-1 ConvTranspose3d (CVP conv5)
    (16, 1, False, (1, 10, 17, 45)),     # single-channel output (prob0 of CVP)
    (32, 64, False, (1, 4, 9, 31)),
]


@pytest.mark.parametrize("dtype", [torch.float16, torch.bfloat16])
@pytest.mark.parametrize("cin,cout,tr,shape", CASES)
def test_tc_matches_simt(gpu, cin, cout, tr, shape, dtype):
    from ssmvs_b200 import ops
    torch.manual_seed(cin * 100 + cout)
    b, d, h, w = shape
    dev = gpu.device
    x = torch.randn(b, cin, d, h, w, device=dev)
    wt = 0.1 * (torch.randn(cin, cout, 3, 3, 3, device=dev) if tr else torch.randn(cout, cin, 3, 3, 3, device=dev))
    wt = wt.to(dtype).float()                                   # weights representable in the storage dtype
    g = ops.pack_conv3d_weight(wt, tr)
    x8 = ops.pack_c8(x, dtype)
    scale = (torch.rand(cout, device=dev) + 0.5) if cout > 1 else None
    shift = torch.randn(cout, device=dev)
    skip = ops.pack_c8(torch.randn(b, cout, d, h, w, device=dev), dtype) if cout > 1 else None
    y_tc = ops.conv3d_raw(x8, g, cout, 1, tr, scale, shift, skip, relu=cout > 1, algo=2)
    y_ref = ops.conv3d_raw(x8, g, cout, 1, tr, scale, shift, skip, relu=cout > 1, algo=1)
    torch.cuda.synchronize()
    tol = 2e-3 if dtype == torch.float16 else 1.6e-2
    err = (y_tc.float() - y_ref.float()).abs().max().item()
    ref = y_ref.float().abs().max().item()
    assert err <= tol * ref, (err, ref)
    # and against ATen on the same rounded inputs (independent of our SIMT kernel)
    xr = ops.unpack_c8(x8)
    want = F.conv_transpose3d(xr, wt, None, 1, 1) if tr else F.conv3d(xr, wt, None, 1, 1)
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
