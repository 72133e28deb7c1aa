"""Kernel-level GPU checks through the C ABI against plain fp64 torch on the same inputs: the
contraction core in both cta_group variants (ragged M/N/K, batching, shared operands, fused epilogues,
triangular k-range skipping), the factor SYRK on ragged / boundary shapes (d' = 176 | 177 SIMT /
tensor-core switch, unaligned rows, batch sizes that are not multiples of 8), implicit-im2col conv
factors with padding and stride, the diagonal kernels, Philox addressing and the batched Cholesky
inversion from d = 1 to d = 1025.  The cases live in tools/gpu_check_core.py (also a stand-alone
bring-up tool); this test runs them and fails on any case out of tolerance."""
import importlib.util
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def core():
    assert torch.cuda.is_available()
    spec = importlib.util.spec_from_file_location("gpu_check_core", ROOT / "tools" / "gpu_check_core.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["gpu_check_core"] = mod
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("cta_group", [0, 1, 2])
@pytest.mark.parametrize("case", ["t_gemm_basic", "t_gemm_epilogue", "t_gemm_tri"])
def test_contraction_core(core, case, cta_group):
    core.FAIL.clear()
    core.L.bk_set_cta_group(cta_group)
    try:
        getattr(core, case)()
    finally:
        core.L.bk_set_cta_group(0)
    assert core.FAIL == []


@pytest.mark.parametrize("case", ["t_syrk", "t_syrk_edges", "t_conv", "t_diag", "t_philox"])
def test_factor_and_elementwise_kernels(core, case):
    core.FAIL.clear()
    getattr(core, case)()
    assert core.FAIL == []


def test_batched_cholesky_inversion(core):
    core.FAIL.clear()
    core.t_chol_small()
    assert core.FAIL == []
