#!/bin/bash
# The CPU kernel tests with the host-emulation build of the SIMT kernel sources under AddressSanitizer + UBSan: every launch of
# the emulated kernels on the tests' shapes (the hypothesis sweeps included) is checked for reads / writes outside the torch
# tensors it was handed (torch's CPU allocator goes through the interposed posix_memalign, so each tensor gets red zones).
# Complements the compute-sanitizer runs of the real kernels on the GPU (profiles/r02_sanitizer_*.log).
#   tools/asan_emu.sh [pytest args]      (default: the kernel-level test files)
set -e
cd "$(dirname "$0")/.."
ASAN=$(gcc -print-file-name=libasan.so)
export MVS_EMU_ASAN=1 ASAN_OPTIONS=detect_leaks=0:abort_on_error=0:halt_on_error=1 LD_PRELOAD="$ASAN"
exec python -m pytest -q -m "not gpu" -p no:cacheprovider "${@:-tests/test_kernels.py tests/test_loss.py tests/test_output_side.py tests/test_fusion.py tests/test_training_tc.py}"
