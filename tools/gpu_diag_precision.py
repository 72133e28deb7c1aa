"""Development diagnostic: where does end-to-end inverse error come from on ill-conditioned factors?"""
import sys

import torch

sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib  # noqa: E402
from bnn_kfac_b200.curvatures import invert_factors  # noqa: E402

L = _lib.load()
dev = torch.device("cuda:0")


def relerr(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def syrk(x, prec):
    n, d = x.shape
    st = torch.zeros(d + 1, d + 1, device=dev)
    wsb = L.bk_syrk_workspace_bytes(n, d, 1, prec)
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    _lib.check(L.bk_syrk_accum(st.data_ptr(), d + 1, x.data_ptr(), d, n, d, 1, 1.0, 1.0 / n, 0.0, prec,
                               ws.data_ptr(), wsb, _lib.stream_ptr()), "syrk")
    return st


def ref_inv(F, add, mult):
    Fd = F.double()
    R = mult ** 0.5 * Fd + add ** 0.5 * torch.eye(F.shape[0], dtype=torch.float64, device=F.device)
    R = (R + R.T) / 2
    return torch.linalg.cholesky(torch.linalg.inv(R)), torch.linalg.cond(R).item()


g = torch.Generator().manual_seed(1234)
for name, x in [("U(0,1) images 784", torch.rand(256, 784, generator=g)),
                ("relu(N(0,1)) 1024", torch.relu(torch.randn(256, 1024, generator=g)))]:
    x = x.to(dev)
    xa = torch.cat([x, torch.ones(256, 1, device=dev)], 1).double()
    A64 = xa.T @ xa / 256
    for add, mult in [(0.04, 200.0), (1.0, 200.0)]:
        Lref, cond = ref_inv(A64, add, mult)
        print(f"== {name} add={add} mult={mult} cond(R)={cond:.3e}")
        for pname, prec in [("fp32", 0), ("bf16x3", 3), ("bf16", 1)]:
            A = syrk(x, prec)
            (Lg,) = invert_factors([A], [add], [mult])
            Lstage, _ = ref_inv(A, add, mult)           # fp64 inverse of the GPU's fp32 factor
            (L64in,) = invert_factors([A64.float()], [add], [mult])
            print(f"  {pname:7s} factor_err={relerr(A, A64):.2e}  e2e_inv_err={relerr(Lg, Lref):.2e}  "
                  f"chol_stage_err={relerr(Lg, Lstage):.2e}  factor_only_err={relerr(Lstage, Lref):.2e}  "
                  f"fp32cast_chol_err={relerr(L64in, Lref):.2e}")
