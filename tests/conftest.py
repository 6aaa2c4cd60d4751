"""Test configuration.

`-m "not gpu"`: oracle vs golden fixtures, host logic, C-ABI export check, and the kernel sources compiled for the
host (tests/emu) checked against the oracle.  `-m gpu`: parity of libmvs_b200.so on a B200 through the C ABI.
"""
import importlib.util
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _load(name, path):
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.fixture(scope="session")
def oracle():
    return _load("planesweep_oracle", os.path.join(ROOT, "oracle", "planesweep.py"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return {k: torch.from_numpy(z[k]) for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def state_dict_of(gold, prefix="sd."):
    return {k[len(prefix):]: v for k, v in gold.items() if k.startswith(prefix)}


class Backend:
    """Which library the ops are bound to and where tensors live."""

    def __init__(self, device):
        self.device = torch.device(device)

    def to(self, t):
        if isinstance(t, dict):
            return {k: self.to(v) for k, v in t.items()}
        if isinstance(t, (list, tuple)):
            return type(t)(self.to(v) for v in t)
        return t.detach().clone().to(self.device) if isinstance(t, torch.Tensor) else t


_EMU_PATH = []


@pytest.fixture
def emu():
    """Bind the host-emulation build of the kernel sources (CPU tensors).  Re-bound per test: other tests bind the product library."""
    import ssmvs_b200
    if not _EMU_PATH:
        sys.path.insert(0, os.path.join(ROOT, "tests", "emu"))
        from build_emu import build_emu
        _EMU_PATH.append(build_emu())
    if ssmvs_b200._lib._lib is None or not ssmvs_b200._lib._emulation:
        ssmvs_b200._lib.bind(_EMU_PATH[0])
    assert ssmvs_b200._lib.is_emulation()
    return Backend("cpu")


@pytest.fixture
def gpu():
    """Bind libmvs_b200.so (CUDA tensors)."""
    import ssmvs_b200
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    if ssmvs_b200._lib._lib is None or ssmvs_b200._lib._emulation:
        ssmvs_b200._lib.bind()
    assert not ssmvs_b200._lib.is_emulation()
    # parity tests compare fp32 with fp32: keep the library (cuDNN / cuBLAS) parts of the models off TF32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return Backend("cuda:0")


@pytest.fixture
def knob():
    """Set test / tuning knobs of the bound library (mvs_set_knob); every knob touched is reset afterwards."""
    import ssmvs_b200
    touched = []

    def setter(name, value):
        touched.append(name)
        ssmvs_b200._lib.set_knob(name, value)
    yield setter
    for name in touched:
        ssmvs_b200._lib.set_knob(name, -1)


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()
