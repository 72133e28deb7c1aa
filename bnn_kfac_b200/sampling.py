"""Matrix-normal posterior weight sampling, batched over samples.

Reference semantics (`KFAC.sample`, /root/reference/models/curvatures.py:400-405):

    z = randn(d_in', d_out);   sample = (L_A @ z @ L_G.t()).t()          -> [d_out, d_in']

Here the two products are two launches of the tcgen05 contraction core for ALL samples at once,
with the triangular structure of L_A / L_G used to skip the zero k-blocks:

    T_s  = L_A  Z_s          A = L_A [d_in', d_in'] (lower, shared), B = Z_s^T [d_out, d_in']
    S_s  = L_G  T_s^T        A = L_G [d_out, d_out] (lower, shared), B = T_s   [d_in', d_out]

Z_s^T is produced directly in operand layout by the Philox kernel (element (o, i) of sample s is a
pure function of (seed, sample id, layer id, o*d_in' + i)), or staged from a caller-supplied `z`
(parity mode: the reference and this engine consume identical noise).
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import Tensor

from . import _lib


def _round8(v: int) -> int:
    return (v + 7) // 8 * 8


def philox_normal_t(seed: int, sample0: int, stream_id: int, d_in_p: int, d_out: int, n_samples: int,
                    device) -> Tensor:
    """The noise the fused path consumes, as fp32 Z^T [S, d_out, d_in'] (tests / parity runs)."""
    lib = _lib.load()
    zf = torch.empty(n_samples, d_out, d_in_p, device=device, dtype=torch.float32)
    _lib.check(lib.bk_philox_normal(seed, sample0, stream_id, d_out, d_in_p, n_samples, zf.data_ptr(),
                                    d_in_p, d_out * d_in_p, 0, 0, 0, 0, _lib.stream_ptr()),
               "bk_philox_normal")
    return zf


def matrix_normal_samples(stA, stG, d_in_p: int, d_out: int, n_samples: int, *, precision: int,
                          seed: int, sample0: int, stream_id: int, z: Optional[Tensor] = None,
                          ws=None, max_chunk_bytes: int = 4 << 30) -> Tensor:
    """Returns fp32 samples [S, d_out, d_in'].

    stA / stG: (hi, lo, ld) bf16 staged Cholesky factors (see curvatures.stage_operand).
    z: optional external noise [S, d_in', d_out] (reference orientation, curvatures.py:404)."""
    lib = _lib.load()
    a_hi, a_lo, lda = stA
    g_hi, g_lo, ldg = stG
    dev = a_hi.device
    st = _lib.stream_ptr()
    x3 = precision == _lib.BK_PREC_BF16X3
    ldz = _round8(d_in_p)   # Z^T rows are d_in' long
    ldt = _round8(d_out)    # T rows are d_out long
    out = torch.empty(n_samples, d_out, d_in_p, device=dev, dtype=torch.float32)
    per_sample = (d_out * ldz + d_in_p * ldt) * 2 * (2 if x3 else 1)
    chunk = max(1, min(n_samples, max_chunk_bytes // max(per_sample, 1)))
    zt_hi = torch.empty(chunk, d_out, ldz, dtype=torch.bfloat16, device=dev)
    zt_lo = torch.empty_like(zt_hi) if x3 else None
    t_hi = torch.empty(chunk, d_in_p, ldt, dtype=torch.bfloat16, device=dev)
    t_lo = torch.empty_like(t_hi) if x3 else None
    if z is not None:
        z = z.to(dev, torch.float32).contiguous()
        assert z.shape == (n_samples, d_in_p, d_out), "z must be [S, d_in', d_out]"
    for s0 in range(0, n_samples, chunk):
        sc = min(chunk, n_samples - s0)
        if z is None:
            _lib.check(lib.bk_philox_normal(seed, sample0 + s0, stream_id, d_out, d_in_p, sc, 0, 0, 0,
                                            zt_hi.data_ptr(), _lib.ptr(zt_lo), ldz, d_out * ldz, st),
                       "bk_philox_normal")
        else:
            for s in range(sc):  # parity mode only; z_s [d_in', d_out] -> Z_s^T [d_out, d_in']
                zs = z[s0 + s]
                _lib.check(lib.bk_transpose_split(zs.data_ptr(), d_out, d_in_p, d_out, 1.0, 0,
                                                  zt_hi[s].data_ptr(),
                                                  zt_lo[s].data_ptr() if x3 else 0, ldz, st),
                           "bk_transpose_split")
        # T_s = L_A Z_s  (bf16 split output, K-major operand of the next product)
        _lib.check(lib.bk_gemm_nt(a_hi.data_ptr(), a_lo.data_ptr() if x3 else 0, lda, 0,
                                  zt_hi.data_ptr(), _lib.ptr(zt_lo), ldz, d_out * ldz,
                                  d_in_p, d_out, d_in_p, sc, precision, _lib.GEMM_TRI_A, 1.0, 0.0,
                                  0, 0, 0, 0, 0,
                                  t_hi.data_ptr(), _lib.ptr(t_lo), ldt, d_in_p * ldt, st),
                   "bk_gemm_nt(L_A Z)")
        # S_s = L_G T_s^T -> [d_out, d_in']
        o = out[s0:s0 + sc]
        _lib.check(lib.bk_gemm_nt(g_hi.data_ptr(), g_lo.data_ptr() if x3 else 0, ldg, 0,
                                  t_hi.data_ptr(), _lib.ptr(t_lo), ldt, d_in_p * ldt,
                                  d_out, d_in_p, d_out, sc, precision, _lib.GEMM_TRI_A, 1.0, 0.0,
                                  o.data_ptr(), d_in_p, d_out * d_in_p, 0, 0,
                                  0, 0, 0, 0, st),
                   "bk_gemm_nt(L_G T^T)")
    return out
