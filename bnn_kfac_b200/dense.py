"""Dense-Fisher Laplace (BASELINE config 3).  Reference recipes (paths relative to /root/reference):

    flat_gradient(model)            hessian/classification_ll_dense_kernel_diag.py:79-84
    dense_fisher(grads)             hessian/classification_ll_dense_kernel_diag.py:85-89
                                    (sampling_free/classification/classification_ll_dense.py:96-104)
    dominance(H, coords, tau)       hessian/utils.py:4-23   (coords: hessian/utils.py:67-95)
    dense_inverse(H, tau)           sampling_free/utils.py:47-53, classification_ll_dense.py:108-109
    dense_variance(J, H_inv)        classification_ll_dense.py:160-161

The reference accumulates H by one rank-1 update of the P x P matrix per mini-batch (2*P^2*4 bytes of
HBM traffic each, 1.8 GB at P = 15 080); here the flat gradients are stacked [n, P] and H is ONE
tensor-core SYRK (the same kernel as the Kronecker factors).  H + tau*I is SPD for tau > 0, so the
script's pseudo-inverse is the inverse and comes from the batched Cholesky path (L L^T)."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
from torch import Tensor

from . import _lib
from .curvatures import _Workspace, _alloc_factor, _round8, invert_factors, stage_operand


def flat_gradient(model: torch.nn.Module, params: Optional[Sequence[torch.nn.Parameter]] = None) -> Tensor:
    """cat over modules()[1:] of cat(flatten(p.grad)) — the parameter order of the dense scripts.
    `params` restricts it to a subset (e.g. the last layer's parameters)."""
    if params is not None:
        return torch.cat([torch.flatten(p.grad.data) for p in params], dim=0)
    g = []
    for layer in list(model.modules())[1:]:
        for p in layer.parameters():
            g.append(torch.flatten(p.grad.data))
    return torch.cat(g, dim=0)


def dense_fisher(grads: Tensor, state: Optional[Tensor] = None, precision: str = "bf16x3",
                 ws: Optional[_Workspace] = None, normalise: Optional[float] = None) -> Tensor:
    """H (+)= grads^T grads / n for stacked flat gradients [n, P] (one SYRK launch sequence).
    `state` accumulates over chunks of gradients (pass normalise = total count to keep one scale)."""
    lib = _lib.load()
    _lib.require_device()
    g = grads.float().contiguous()
    n, P = g.shape
    beta = 1.0
    if state is None:
        state = _alloc_factor(P, g.device)
        beta = 0.0
    prec = {"fp32": _lib.BK_PREC_FP32, "bf16": _lib.BK_PREC_BF16, "bf16x3": _lib.BK_PREC_BF16X3}[precision]
    nbytes = lib.bk_syrk_workspace_bytes(n, P, 0, prec)
    ws = ws or _Workspace()
    buf = ws.get(nbytes, g.device)
    _lib.check(lib.bk_syrk_accum(state.data_ptr(), state.stride(0), g.data_ptr(), g.stride(0), n, P, 0, 1.0,
                                 1.0 / float(normalise if normalise is not None else n), beta, prec,
                                 buf.data_ptr(), nbytes, _lib.stream_ptr()), "bk_syrk_accum")
    return state


def kernel_block_coords_basenet15k() -> List[Tuple[int, int]]:
    """Per-kernel diagonal block ranges of BaseNet_15k's flat parameter vector (hessian/utils.py:67-95)."""
    coords, curr = [], 0
    for count, size, bias in ((5, 25, 5), (10, 125, 10), (80, 160, 80), (10, 80, 10)):
        for _ in range(count):
            coords.append((curr, curr + size))
            curr += size
        coords.append((curr, curr + bias))
        curr += bias
    return coords


def dominance(H: Tensor, coords: Sequence[Tuple[int, int]], tau: float = 1e-5) -> Tuple[float, float]:
    """(sum|diag| / sum|all|, sum|kernel blocks| / sum|all|) of H + tau*I in one pass over H."""
    lib = _lib.load()
    _lib.require_device()
    Hf = H.float()
    if Hf.stride(1) != 1:
        Hf = Hf.contiguous()
    dev = Hf.device
    coords = sorted(coords)
    bb = torch.tensor([c[0] for c in coords], dtype=torch.int32, device=dev)
    be = torch.tensor([c[1] for c in coords], dtype=torch.int32, device=dev)
    out = torch.empty(3, dtype=torch.float64, device=dev)
    _lib.check(lib.bk_dominance(Hf.data_ptr(), Hf.stride(0), Hf.shape[0], float(tau), bb.data_ptr(),
                                be.data_ptr(), len(coords), out.data_ptr(), _lib.stream_ptr()), "bk_dominance")
    s_diag, s_all, s_blk = out.tolist()
    return s_diag / s_all, s_blk / s_all


def dense_inverse(H: Tensor, tau: float, ws: Optional[_Workspace] = None) -> Tensor:
    """(H + tau*I)^-1 = L L^T with L = chol(inv(H + tau I)) from the batched Cholesky kernels."""
    from .predictive import inverse_from_chol
    (Lc,) = invert_factors([H], [float(tau) ** 2], [1.0], ws)
    return inverse_from_chol(Lc)


def dense_variance(J: Tensor, H_inv: Tensor, precision: int = _lib.BK_PREC_BF16X3) -> Tensor:
    """|J_b H_inv J_b^T| for Jacobian rows J [B, P]: one GEMM T = J H_inv + a fused row-wise dot."""
    lib = _lib.load()
    J = J.float().contiguous()
    Bn, P = J.shape
    dev = J.device
    x3 = precision == _lib.BK_PREC_BF16X3
    j_hi, j_lo, ldj = stage_operand(J)
    h_hi, h_lo, ldh = stage_operand(H_inv)          # symmetric: H_inv^T == H_inv
    T = torch.empty(Bn, P, device=dev, dtype=torch.float32)
    _lib.check(lib.bk_gemm_nt(j_hi.data_ptr(), j_lo.data_ptr() if x3 else 0, ldj, 0,
                              h_hi.data_ptr(), h_lo.data_ptr() if x3 else 0, ldh, 0,
                              Bn, P, P, 1, precision, 0, 1.0, 0.0, T.data_ptr(), P, 0, 0, 0, 0, 0, 0, 0,
                              _lib.stream_ptr()), "bk_gemm_nt(J H^-1)")
    out = torch.empty(Bn, device=dev, dtype=torch.float32)
    _lib.check(lib.bk_frob_dot(out.data_ptr(), J.data_ptr(), P, T.data_ptr(), P, P, Bn, 1, 0,
                               _lib.stream_ptr()), "bk_frob_dot")
    return out
