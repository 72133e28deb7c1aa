"""GPU parity tests, round 2 second half: the remaining SURVEY §8 rows (f2 kernel-block-diagonal helpers, e4 / e5 /
e6 sharded entry points at world size 1, wide-conv tensor-core factors, fused Philox sampling GEMM), all through
the public Python API -> C ABI -> sm_100a kernels, against outputs of the reference itself (tests/golden/) or the
CPU oracle on identical seeded inputs.  Tolerance: BASELINE.json's 1e-3 unless a test states another one."""
import numpy as np
import pytest
import torch

from conftest import ROOT, relerr
from oracle import kfac_oracle as O

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from bnn_kfac_b200 import _lib
    _lib.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def golden_kd():
    return dict(np.load(ROOT / "tests" / "golden" / "reference_golden_kernel_diag.npz"))


# ------------------------------------------------------------------------------------------ f2, second half
@pytest.mark.parametrize("name", ["kd748", "kd141", "kdreg5", "kdreg30"])
def test_kernel_diag_vs_reference_golden(golden_kd, dev, name):
    """sampling_free/utils.py:63-211 against outputs of the reference's own generate_kernel_diag* functions,
    including the in-place `H += tau I` and the overlapping first-layer blocks of the regression variant."""
    from bnn_kfac_b200 import dense
    g = golden_kd
    P, n, tau, scale, n_hid = g[f"{name}_meta"]
    P, n_hid, tau, scale = int(P), (None if n_hid < 0 else int(n_hid)), float(tau), float(scale)
    G = torch.tensor(g[f"{name}_G"]).double()
    H64 = G.t() @ G / G.shape[0]
    H = H64.float().to(dev)
    if name == "kd748":
        res, inv = dense.generate_kernel_diag_748(H, tau)
    elif name == "kd141":
        res, inv = dense.generate_kernel_diag_141(H, tau, scale)
    else:
        res, inv = dense.generate_kernel_diag(H, tau, scale, n_hid)
    assert relerr(H.cpu(), (H64 + tau * torch.eye(P, dtype=torch.float64)).numpy()) < 1e-6   # side effect kept
    assert relerr(res.cpu(), g[f"{name}_res"]) < 1e-6
    assert relerr(inv.cpu(), g[f"{name}_inv"]) < TOL
    # exact zeros outside the block union, like the reference's torch.zeros_like + slice assignment
    assert torch.equal(res.cpu() == 0, torch.tensor(g[f"{name}_res"]) == 0)
    if name == "kd141":
        H0 = H64.float().to(dev)
        d_res, d_inv = dense.generate_diag(H0, tau)
        assert relerr(d_res.cpu(), g["kd141_diag_res"]) < 1e-6 and relerr(d_inv.cpu(), g["kd141_diag_inv"]) < 1e-6
        h_reg, h_inv = dense.generate_H(H0, tau)
        assert relerr(h_reg.cpu(), g["kd141_H_reg"]) < 1e-6 and relerr(h_inv.cpu(), g["kd141_H_inv"]) < TOL
        h_reg2, h_inv2 = dense.generate_H_true(H0, tau)
        assert torch.equal(h_inv2, h_inv)
        assert abs(dense.calculate_dominance(H0) - float(g["kd141_dominance"])) < 1e-5
        assert relerr(H0.cpu(), H64.numpy()) < 1e-6                                          # not modified


def test_kernel_diag_15080_blocks_vs_fp64(dev):
    """BASELINE config 3 size (P = 15 080, 109 blocks up to 160 wide): every block of the inverse against an fp64
    torch.inverse of the same block; NotImplementedError for any other size, like the reference."""
    from bnn_kfac_b200 import dense
    P, n = 15080, 512
    gen = torch.Generator().manual_seed(5)
    G = (0.3 * torch.randn(n, P, generator=gen)).to(dev)
    H = dense.dense_fisher(G, precision="bf16x3")
    H = H.contiguous()
    H0 = H.clone()
    res, inv = dense.generate_kernel_diag_15080(H, 0.04)
    coords = O.kernel_coords(P)
    worst = 0.0
    mask = torch.zeros(P, P, dtype=torch.bool, device=dev)
    for a, b in coords:
        mask[a:b, a:b] = True
        blk = H0[a:b, a:b].double() + 0.04 * torch.eye(b - a, device=dev, dtype=torch.float64)
        assert float((res[a:b, a:b].double() - blk).norm() / blk.norm()) < 1e-6
        ref = torch.linalg.inv(blk)
        worst = max(worst, float((inv[a:b, a:b].double() - ref).norm() / ref.norm()))
    assert worst < 1e-4, worst
    assert not res[~mask].any() and not inv[~mask].any()
    with pytest.raises(NotImplementedError):
        dense.generate_kernel_diag_15080(H[:100, :100].contiguous(), 0.04)
