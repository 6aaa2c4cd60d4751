"""The C-ABI boundary: header <-> exports <-> ctypes table, error behaviour, no-fallback behaviour.  CPU only."""
import ctypes
import os
import re

import pytest
import torch

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "mvs_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mvs_[a-z0-9_]+)\s*\(", src)))


def test_header_matches_ctypes_table():
    import ssmvs_b200
    assert sorted(ssmvs_b200._lib.EXPORTS) == _declared()


def test_library_exports_every_declared_symbol():
    """libmvs_b200.so (built by __graft_entry__.build()) must load on a box without a GPU and export the ABI."""
    import ssmvs_b200
    path = ssmvs_b200._lib.DEFAULT_PATH
    if not os.path.exists(path):
        import __graft_entry__
        __graft_entry__.build()
    lib = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(lib, name), name
    lib.mvs_version.restype = ctypes.c_int
    assert lib.mvs_version() == 102
    lib.mvs_is_emulation.restype = ctypes.c_int
    assert lib.mvs_is_emulation() == 0


def test_product_library_refuses_cpu_tensors():
    """No CPU fallback: with the real library bound, host tensors raise instead of silently computing."""
    import ssmvs_b200
    ssmvs_b200._lib.bind()
    with pytest.raises(RuntimeError, match="CUDA"):
        ssmvs_b200.ops.pack_c8(torch.zeros(1, 8, 4, 4))


def test_missing_library_fails_loudly(tmp_path):
    import ssmvs_b200
    with pytest.raises(RuntimeError, match="not found"):
        ssmvs_b200._lib.bind(str(tmp_path / "nope.so"))
    ssmvs_b200._lib.bind()


def test_error_codes_and_messages(emu):
    import ssmvs_b200
    from ssmvs_b200 import ops
    with pytest.raises(ValueError):
        ops.pack_c8(torch.zeros(1, 7, 4, 4))                      # C % 8
    lib = ssmvs_b200._lib.lib()
    rc = lib.mvs_pack_c8(None, None, 1, 8, 16, 0, None)
    assert rc == -1 and b"null" in lib.mvs_last_error()
    rc = lib.mvs_compose_proj(1, 1, 1, 12, None)                   # too many views (pointers never dereferenced)
    assert rc == -2 and b"views" in lib.mvs_last_error()
    x = ops.pack_c8(torch.zeros(1, 8, 3, 4, 4))
    g = torch.zeros(27, 8, 8)
    with pytest.raises(ValueError, match="even"):
        ops.conv3d_raw(x, g, 8, stride=2)                          # odd extent under stride 2
    with pytest.raises(RuntimeError, match="emulation"):
        ops.conv3d_raw(ops.pack_c8(torch.zeros(1, 8, 4, 4, 4)), g, 8, algo=2)


def test_new_entry_points_validate_before_touching_the_device():
    """Argument checks of the 2-D convolution / padded packing / image packing entry points run on a box without a GPU
    (they return before any CUDA call): null pointers, inconsistent extents, unknown layouts."""
    import ssmvs_b200
    from ssmvs_b200._lib import Conv2dDesc, DEFAULT_PATH, _SIGS
    lib = ctypes.CDLL(DEFAULT_PATH)          # a private handle: the session-wide binding (emulation for the CPU tests) stays as it is
    for name in ("mvs_conv2d_fwd", "mvs_conv2d_workspace_bytes", "mvs_pack_c8_padded", "mvs_pack_images_c8", "mvs_last_error"):
        getattr(lib, name).argtypes, getattr(lib, name).restype = _SIGS[name]
    d = Conv2dDesc(2, 8, 8, 16, 16, 16, 16, 3, 1, 1, 1, 0, 0, 0.0)
    assert lib.mvs_conv2d_fwd(ctypes.byref(d), None, None, None, None, None, None, None) == -1 and b"null" in lib.mvs_last_error()
    d.Hout = 8                                                      # stride 1 must keep the extent
    assert lib.mvs_conv2d_fwd(ctypes.byref(d), 1, 1, None, None, 1, None, None) == -2 and b"extent" in lib.mvs_last_error()
    assert lib.mvs_pack_c8_padded(None, None, 1, 8, 4, 4, 0, 0, None) == -1
    assert lib.mvs_pack_c8_padded(1, 1, 1, 7, 4, 4, 0, 0, None) == -2 and b"multiple of 8" in lib.mvs_last_error()
    assert lib.mvs_pack_c8_padded(1, 1, 1, 8, 4, 4, 5, 0, None) == -1 and b"src_layout" in lib.mvs_last_error()
    assert lib.mvs_pack_images_c8(None, 0, None, 1, 1, 4, 4, 1, None) == -1
    # unsupported 2-D layer shapes are reported, not silently computed some other way
    bad = Conv2dDesc(2, 8, 8, 16, 16, 16, 16, 7, 1, 1, 1, 0, 0, 0.0)
    assert lib.mvs_conv2d_workspace_bytes(ctypes.byref(bad)) == 0


def test_round2_entry_points_validate_before_touching_the_device():
    """Argument checks of the round-2 entry points (fused loss, FeatureNet front, 2-D weight gradient, output side, fusion) on
    the PRODUCT library without a GPU: they return an error code and a message before any CUDA call -- null pointers, fewer than
    three source views (the reference's top-k with k = 3, hazard H5), inconsistent extents, unsupported storage types."""
    from ssmvs_b200._lib import Conv3dDesc, DEFAULT_PATH, _SIGS
    lib = ctypes.CDLL(DEFAULT_PATH)
    for name in ("mvs_unsup_loss_fwd", "mvs_unsup_loss_bwd", "mvs_featnet_front", "mvs_featnet_front_pack", "mvs_featnet_front_workspace_bytes",
                 "mvs_conv2d_wgrad_mma", "mvs_upsample_nearest", "mvs_depth_preview_u8", "mvs_geo_consistency", "mvs_fusibile", "mvs_last_error"):
        getattr(lib, name).argtypes, getattr(lib, name).restype = _SIGS[name]
    p = 1        # a non-null "pointer" that is never dereferenced on the failing paths
    assert lib.mvs_unsup_loss_fwd(None, p, p, 1, 5, 64, 80, 16, 20, 1.0, 0.18, p, p, p, p, p, p, p, None) == -1 and b"null" in lib.mvs_last_error()
    assert lib.mvs_unsup_loss_fwd(p, p, p, 1, 3, 64, 80, 16, 20, 1.0, 0.18, p, p, p, p, p, p, p, None) == -2 and b"3 source views" in lib.mvs_last_error()
    assert lib.mvs_unsup_loss_fwd(p, p, p, 1, 5, 60, 80, 16, 20, 1.0, 0.18, p, p, p, p, p, p, p, None) == -2 and b"4x" in lib.mvs_last_error()
    assert lib.mvs_unsup_loss_bwd(p, p, p, p, p, p, p, p, None, 1, 5, 16, 20, 1.0, 0.18, None) == -1
    assert lib.mvs_featnet_front_workspace_bytes() == 38 * 64 * 4
    assert lib.mvs_featnet_front(p, 0, p, p, p, 1, 5, 64, 80, 0, None) == -4 and b"16-bit" in lib.mvs_last_error()      # fp32 storage: not on this kernel
    assert lib.mvs_featnet_front(p, 2, p, p, p, 1, 5, 64, 80, 1, None) == -4 and b"volume dtype" in lib.mvs_last_error()  # bf16 images into fp16 volumes
    assert lib.mvs_featnet_front(p, 1, p, p, p, 1, 5, 63, 80, 1, None) == -2
    assert lib.mvs_featnet_front_pack(p, p, None, p, 1, None) == -1
    d = Conv3dDesc(1, 8, 8, 2, 16, 16, 2, 16, 16, 1, 0, 0, 0, 0, 0)                                                          # fp32 storage
    assert lib.mvs_conv2d_wgrad_mma(ctypes.byref(d), p, p, p, 8, None) == -4 and b"16-bit" in lib.mvs_last_error()
    assert lib.mvs_upsample_nearest(p, None, 1, 4, 4, 8, 8, 0, None) == -1
    assert lib.mvs_upsample_nearest(p, p, 1, 4, 4, 0, 8, 0, None) == -2
    assert lib.mvs_depth_preview_u8(p, p, 16, 500.0, 0.0, None) == -2
    assert lib.mvs_geo_consistency(p, None, p, p, p, p, p, p, p, 1, 8, 8, 1.0, 0.01, 1, None) == -1
    assert lib.mvs_geo_consistency(p, p, p, p, p, p, p, p, p, 1, 8, 40000, 1.0, 0.01, 1, None) == -2 and b"32767" in lib.mvs_last_error()
    assert lib.mvs_fusibile(p, None, p, p, 3, 4, 8, 8, 7, 0.25, 0.52, 3, p, p, None) == -2                               # ref view out of range
    assert lib.mvs_fusibile(None, None, p, p, 3, 4, 8, 8, 0, 0.25, 0.52, 3, p, p, None) == -1


def test_concurrent_callers_get_their_own_results_and_error_strings(emu):
    """SURVEY 8b threading contract: the library keeps no mutable state between calls beyond a THREAD-LOCAL error string, so
    nn.DataParallel's per-device worker threads may call it concurrently.  Checked on the host build (ctypes drops the GIL
    around each call, so the two threads really overlap): results equal the serial ones, and an error raised on one thread
    does not show up in the other thread's mvs_last_error()."""
    import threading
    import ssmvs_b200
    from ssmvs_b200 import ops, synth
    cases = []
    for seed in range(2):
        li = synth.mvsnet_inputs(1, 3, 64, 96, 8, seed=seed)
        feats = [torch.randn(1, 8, 16, 24, generator=torch.Generator().manual_seed(10 * seed + v)) for v in range(3)]
        cases.append((feats, ops.compose_proj(li["proj_matrices"]), li["depth_values"]))
    serial = [ops.warp_variance(f[0], f[1:], rt, dv, torch.float32, False, False) for f, rt, dv in cases]
    got, errs = [None, None], [None, None]
    gate = threading.Barrier(2)
    lib = ssmvs_b200._lib.lib()

    def worker(i):
        f, rt, dv = cases[i]
        gate.wait()
        for _ in range(4):
            got[i] = ops.warp_variance(f[0], f[1:], rt, dv, torch.float32, False, False)
        gate.wait()
        if i == 0:
            assert lib.mvs_compose_proj(1, 1, 1, 12, None) == -2           # too many views: sets this thread's message only
        gate.wait()
        errs[i] = bytes(lib.mvs_last_error())

    th = [threading.Thread(target=worker, args=(i,)) for i in range(2)]
    [t.start() for t in th]
    [t.join(120) for t in th]
    assert all(not t.is_alive() for t in th)
    for i in range(2):
        assert torch.equal(got[i], serial[i])
    assert b"views" in errs[0] and b"views" not in errs[1]


def test_plain_c_client_of_the_header(emu, tmp_path):
    """include/mvs_b200.h is the boundary a binding in the reference's host language would compile against: it must be valid
    C99 on its own (pedantic), and a C program using only that header must link and run.  tests/cabi_client.c is run against
    the host build (host pointers) and LINKED against the product library (every symbol it uses resolves; no GPU needed)."""
    import subprocess
    import ssmvs_b200
    inc = os.path.join(ROOT, "include")
    src = os.path.join(ROOT, "tests", "cabi_client.c")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", "-x", "c", HEADER], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    from build_emu import build_emu
    emu_lib = build_emu()
    for lib_path, run in ((emu_lib, True), (ssmvs_b200._lib.DEFAULT_PATH, False)):
        if not os.path.exists(lib_path):
            continue
        exe = str(tmp_path / ("client_" + os.path.basename(lib_path)))
        cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", inc, src, "-o", exe, lib_path, "-lm",
               "-Wl,-rpath," + os.path.dirname(lib_path)]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
        if run:
            r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
            assert r.returncode == 0 and r.stdout.strip().endswith("ok"), r.stdout + r.stderr
