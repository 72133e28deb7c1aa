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


# ------------------------------------------------------------------------------------------------------
# Kernel-block-diagonal approximations (sampling_free/utils.py:42-211).  Same names, arguments and return
# pairs as the reference's helpers; `H` is a device tensor and, exactly like the reference's `H += diag * tau`,
# the generate_kernel_diag* functions add tau to the caller's diagonal IN PLACE.
_KERNEL_SPECS = {
    # (blocks, block size, advance, bias block) per layer, in flat-parameter order
    15080: ((5, 25, 25, 5), (10, 125, 125, 10), (80, 160, 160, 80), (10, 80, 80, 10)),   # utils.py:65-93
    748: ((3, 9, 9, 3), (6, 27, 27, 6), (10, 54, 54, 10)),                               # utils.py:108-129
    141: ((10, 1, 1, 10), (10, 10, 10, 10), (1, 10, 10, 1)),                             # utils.py:146-167
}


def kernel_block_coords(spec) -> List[Tuple[int, int]]:
    """[a, b) ranges from (count, size, advance, bias) rows: `count` blocks of `size` whose start advances by
    `advance`, then one bias block.  The regression variant (utils.py:181-201) advances its first group by 1
    while its blocks are n_hid wide - overlapping squares, reproduced as they are."""
    coords, curr = [], 0
    for count, size, advance, bias in spec:
        for _ in range(count):
            coords.append((curr, curr + size))
            curr += advance
        coords.append((curr, curr + bias))
        curr += bias
    return coords


def kernel_block_coords_regression(n_hid: int) -> List[Tuple[int, int]]:
    """Coordinates of generate_kernel_diag(H, tau, n, n_hid) (utils.py:181-201)."""
    return kernel_block_coords(((n_hid, n_hid, 1, n_hid), (n_hid, n_hid, n_hid, n_hid), (1, n_hid, n_hid, 1)))


def band_bounds(coords: Sequence[Tuple[int, int]], P: int):
    """Host bookkeeping for bk_band_mask / bk_block_inverse: per-row extreme bounds of the blocks containing
    the row, and the connected components of the union of the blocks (merged overlapping intervals)."""
    import numpy as np
    lo = np.zeros(P, dtype=np.int32)
    hi = np.zeros(P, dtype=np.int32)
    seen = np.zeros(P, dtype=bool)
    for a, b in coords:
        a, b = max(int(a), 0), min(int(b), P)
        if b <= a:
            continue
        rows = slice(a, b)
        lo[rows] = np.where(seen[rows], np.minimum(lo[rows], a), a)
        hi[rows] = np.where(seen[rows], np.maximum(hi[rows], b), b)
        seen[rows] = True
    comps: List[List[int]] = []
    for a, b in sorted((max(int(a), 0), min(int(b), P)) for a, b in coords):
        if b <= a:
            continue
        if comps and a < comps[-1][1]:
            comps[-1][1] = max(comps[-1][1], b)
        else:
            comps.append([a, b])
    return lo, hi, [(a, b) for a, b in comps]


def masked_blocks(H: Tensor, coords: Sequence[Tuple[int, int]], tau: float = 0.0, in_place: bool = True) -> Tensor:
    """res with res[a:b, a:b] = (H + tau I)[a:b, a:b] for every block, zero elsewhere (one pass, bk_band_mask).
    in_place=True also leaves H += tau I in the caller's tensor (the reference's side effect)."""
    lib = _lib.load()
    _lib.require_device()
    if H.dtype != torch.float32 or H.stride(1) != 1:
        raise ValueError("H must be an fp32 row-major device matrix (it is updated in place)")
    P = H.shape[0]
    lo, hi, _ = band_bounds(coords, P)
    dev = H.device
    lo_d = torch.from_numpy(lo).to(dev)
    hi_d = torch.from_numpy(hi).to(dev)
    res = torch.empty(P, P, device=dev, dtype=torch.float32)
    _lib.check(lib.bk_band_mask(H.data_ptr(), H.stride(0), P, float(tau), int(in_place), lo_d.data_ptr(),
                                hi_d.data_ptr(), res.data_ptr(), res.stride(0), _lib.stream_ptr()), "bk_band_mask")
    return res


def block_inverse(res: Tensor, coords: Sequence[Tuple[int, int]], n: float = 1.0,
                  ws: Optional[_Workspace] = None) -> Tensor:
    """inverse(n * res) of a block-masked matrix: one fp64 shared-memory Gauss-Jordan per connected component of
    the block union (bk_block_inverse); components wider than BK_BLOCK_INV_MAX_DIM go through the batched
    Cholesky path (they must then be positive definite).  Rows outside every block would make n * res singular
    (torch.inverse raises): here that is a ValueError."""
    from .predictive import inverse_from_chol
    lib = _lib.load()
    P = res.shape[0]
    dev = res.device
    _, _, comps = band_bounds(coords, P)
    covered = sum(b - a for a, b in comps)
    if covered != P:
        raise ValueError("the blocks do not cover every row: n * res is singular")
    small = [(a, b) for a, b in comps if b - a <= _lib.BK_BLOCK_INV_MAX_DIM]
    large = [(a, b) for a, b in comps if b - a > _lib.BK_BLOCK_INV_MAX_DIM]
    out = torch.empty(P, P, device=dev, dtype=torch.float32)
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    st = _lib.stream_ptr()
    if small:
        cb = torch.tensor([a for a, _ in small], dtype=torch.int32, device=dev)
        ce = torch.tensor([b for _, b in small], dtype=torch.int32, device=dev)
        _lib.check(lib.bk_block_inverse(res.data_ptr(), res.stride(0), P, cb.data_ptr(), ce.data_ptr(), len(small),
                                        max(b - a for a, b in small), float(n), out.data_ptr(), out.stride(0), 1,
                                        status.data_ptr(), st), "bk_block_inverse")
    else:
        out.zero_()
    for a, b in large:
        (Lc,) = invert_factors([res[a:b, a:b]], [0.0], [float(n) ** 2], ws)
        out[a:b, a:b].copy_(inverse_from_chol(Lc))
    code = int(status.item())
    if code:
        raise RuntimeError(f"block {code >> 16} of the kernel-diagonal matrix is singular (pivot {code & 0xffff})")
    return out


def generate_kernel_diag_15080(H: Tensor, tau: float = 0):
    """(res, inverse(res)) for BaseNet_15k's flat parameter vector (sampling_free/utils.py:63-103)."""
    if H.numel() != 15080 ** 2:
        raise NotImplementedError
    coords = kernel_block_coords(_KERNEL_SPECS[15080])
    res = masked_blocks(H, coords, tau)
    return res, block_inverse(res, coords)


def generate_kernel_diag_748(H: Tensor, tau: float = 0):
    """BaseNet_750 (sampling_free/utils.py:105-138)."""
    if H.numel() != 748 ** 2:
        raise NotImplementedError
    coords = kernel_block_coords(_KERNEL_SPECS[748])
    res = masked_blocks(H, coords, tau)
    return res, block_inverse(res, coords)


def generate_kernel_diag_141(H: Tensor, tau: float = 0, n: float = 1):
    """Regression net with 10 hidden units (sampling_free/utils.py:140-177): returns (res, inverse(n * res))."""
    if H.numel() != 141 ** 2:
        raise NotImplementedError
    coords = kernel_block_coords(_KERNEL_SPECS[141])
    res = masked_blocks(H, coords, tau)
    return res, block_inverse(res, coords, n)


def generate_kernel_diag(H: Tensor, tau: float = 0, n: float = 1, n_hid: int = 1):
    """Regression net with n_hid hidden units (sampling_free/utils.py:179-211), overlapping first-layer blocks
    included (regression_ll_kernel.py:134 calls it with n_hid = 30)."""
    coords = kernel_block_coords_regression(int(n_hid))
    res = masked_blocks(H, coords, tau)
    return res, block_inverse(res, coords, n)


def generate_diag(H: Tensor, tau: float = 0):
    """(diag(diag(H) + tau), diag(1 / (diag(H) + tau)))  (sampling_free/utils.py:42-45); H is not modified."""
    P = H.shape[0]
    coords = [(i, i + 1) for i in range(P)]
    res = masked_blocks(H, coords, tau, in_place=False)
    return res, block_inverse(res, coords)


def generate_H(H: Tensor, tau: float = 0, ws: Optional[_Workspace] = None):
    """(H + tau I, pinv(H + tau I))  (sampling_free/utils.py:47-53).  H + tau I is positive definite for the
    Fisher matrices this is called on (tau > 0), so the pseudo-inverse is the inverse (batched Cholesky path)."""
    P = H.shape[0]
    reg = masked_blocks(H, [(0, P)], tau, in_place=False)
    return reg, dense_inverse(reg, 0.0, ws)


def generate_H_true(H: Tensor, tau: float = 0, ws: Optional[_Workspace] = None):
    """(H + tau I, inverse(H + tau I))  (sampling_free/utils.py:55-61); ValueError when not invertible."""
    try:
        return generate_H(H, tau, ws)
    except RuntimeError as exc:
        raise ValueError('H + tau*eye not invertible!') from exc


def calculate_dominance(H: Tensor, tau: float = 0.00001) -> float:
    """sum|diag| / sum|all| of H + tau I  (sampling_free/utils.py:6-21, without the prints)."""
    return dominance(H, [], tau)[0]
