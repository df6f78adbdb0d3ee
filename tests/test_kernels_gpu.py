"""Single-kernel parity on the B200: every check of tools/diag_kernels.py as a test. Reference = torch fp32 CPU math
on the same bf16-rounded operands; tolerances: 1e-4 where the kernel output is fp32 (accumulation order only),
4e-3 where it is rounded to bf16 (one rounding, 2^-9 relative)."""
import os
import sys

import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))

pytestmark = pytest.mark.gpu

BF16_OUT = {"y", "y_eval", "dx", "dpre1_bf", "nhwc", "nchw", "E"}     # result keys stored in bf16 by the kernel
BF16_CHECKS = {"conv3x3_bf16out", "conv3x3_splitk_odd", "conv3x3_persistent_n256"}


def _checks():
    import diag_kernels
    return diag_kernels.CHECKS


def pytest_generate_tests(metafunc):
    if "check" in metafunc.fixturenames:
        cs = _checks()
        metafunc.parametrize("check", cs, ids=[c[0] for c in cs])


def test_kernel_parity(check):
    from torchsr_b200 import ops
    name, fn = check
    res = fn()
    ops.check_watchdog()
    for key, err in res.items():
        tol = 4e-3 if (key in BF16_OUT or name in BF16_CHECKS or (name == "upsample" and key == "dx")) else 1e-4
        assert err <= tol, f"{name}.{key}: rel-L2 {err:.3e} > {tol:.0e}"
