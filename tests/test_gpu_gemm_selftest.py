"""Runs the stand-alone C++ self-test of the tcgen05 GEMM kernels (tfkaldi_b200/csrc/selftest_gemm.cu):
every operand-major / epilogue / split-K / bf16x3 combination of both the 1-CTA and the CTA-pair kernel
against a double-precision host reference, on small, ragged and multi-tile shapes."""
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tfkaldi_b200", "csrc", "build", "selftest_gemm")


@pytest.mark.timeout(900)
def test_gemm_selftest_quick(cuda_device):
    if not os.path.exists(BIN):
        import __graft_entry__

        __graft_entry__.build()
    out = subprocess.run([BIN, "quick"], capture_output=True, text=True, timeout=600)
    tail = "\n".join(out.stdout.splitlines()[-45:])
    assert out.returncode == 0 and "ALL PASS" in out.stdout, tail + out.stderr[-2000:]
    assert out.stdout.count("[PASS]") >= 40 and "[FAIL]" not in out.stdout, tail
