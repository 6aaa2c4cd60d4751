"""Build the host-emulation library used by the CPU unit tests (TEST INFRASTRUCTURE, never shipped).

The same SIMT kernel sources as libmvs_b200.so are compiled with g++ and -DMVS_CPU_EMU: a launch becomes a
serial loop over (block, thread), fp32 storage only, no tcgen05 path.  This lets `pytest -m "not gpu"` check
index arithmetic, layouts, gradient formulas and the whole Python host stack against the oracle on a box
without a GPU.  The package never loads this library on its own; tests bind it explicitly.
"""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "self-supervised-mvs_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libmvs_emu.so")
SOURCES = ["core.cu", "warp.cu", "softargmin.cu", "conv3d_simt.cu", "invwarp.cu", "loss.cu", "output.cu", "fusion.cu", "featnet_front.cu", "train.cu"]


def build_emu() -> str:
    """MVS_EMU_ASAN=1: the same build with AddressSanitizer + UBSan (tools/asan_emu.sh runs the CPU kernel tests under it; the
    interpreter must then be started with libasan preloaded)."""
    os.makedirs(OUT, exist_ok=True)
    asan = os.environ.get("MVS_EMU_ASAN", "0") == "1"
    LIB = os.path.join(OUT, "libmvs_emu_asan.so" if asan else "libmvs_emu.so")
    srcs = [os.path.join(CSRC, s) for s in SOURCES]
    deps = srcs + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".h")] + [os.path.join(ROOT, "include", "mvs_b200.h")]
    if os.path.exists(LIB) and all(os.path.getmtime(d) <= os.path.getmtime(LIB) for d in deps):
        return LIB
    cmd = ["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-DMVS_CPU_EMU", "-Wno-unknown-pragmas", "-o", LIB]
    if asan:
        cmd[3:3] = ["-g", "-fno-omit-frame-pointer", "-fsanitize=address,undefined", "-fno-sanitize-recover=undefined"]
    for s in srcs:
        cmd += ["-x", "c++", s]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulation build failed:\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build_emu())
