"""Helpers of the reference's `models/utilities.py` that belong to the Kronecker-factored Laplace path
(paths relative to /root/reference):

    kron(a, b)                      models/utilities.py:387-409 (einsum Kronecker product)
    get_eigenvectors(factors)       models/utilities.py:144-159 (symeig of F + F^T per factor)
    get_eigenvalues(factors)        models/utilities.py:120-141 (ger of the factors' eigenvalues)

`torch.symeig` no longer exists, so the reference versions crash on a current torch; these run the
batched one-sided Jacobi eigensolver of libbk_kfac.so (bk_eigh_batched).  Metrics / plotting / argparse
helpers of that file are out of scope."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch
from torch import Tensor
from torch.nn import Module

from . import _lib
from .curvatures import _Workspace


def kron(a: Tensor, b: Tensor) -> Tensor:
    """Kronecker product of two 2-D matrices (same result as the reference's einsum)."""
    lib = _lib.load()
    _lib.require_device()
    af, bf = a.float().contiguous(), b.float().contiguous()
    out = torch.empty(a.size(0) * b.size(0), a.size(1) * b.size(1), device=af.device, dtype=torch.float32)
    _lib.check(lib.bk_kron(af.data_ptr(), af.size(0), af.size(1), bf.data_ptr(), bf.size(0), bf.size(1),
                           out.data_ptr(), _lib.stream_ptr()), "bk_kron")
    return out.to(a.dtype) if a.dtype != torch.float32 else out


def eigh_factors(mats: Sequence[Tensor], sym_scale: float = 0.5, vectors: bool = True,
                 max_sweeps: int = 30, ws: Optional[_Workspace] = None
                 ) -> Tuple[List[Tensor], List[Optional[Tensor]]]:
    """Eigendecomposition of S_i = sym_scale * (F_i + F_i^T) for a batch of square fp32 matrices, one
    batched launch sequence: (eigenvalues ascending [d], eigenvectors as columns [d, d])."""
    lib = _lib.load()
    _lib.require_device()
    n = len(mats)
    if n == 0:
        return [], []
    fs = [m.float() if m.stride(-1) == 1 else m.float().contiguous() for m in mats]
    dev = fs[0].device
    evals = [torch.empty(f.shape[0], device=dev, dtype=torch.float32) for f in fs]
    evecs = [torch.empty(f.shape[0], f.shape[0], device=dev, dtype=torch.float32) if vectors else None
             for f in fs]
    dims = (C.c_int * n)(*[f.shape[0] for f in fs])
    lds = (C.c_longlong * n)(*[f.stride(0) for f in fs])
    fptr = (C.c_void_p * n)(*[f.data_ptr() for f in fs])
    wptr = (C.c_void_p * n)(*[e.data_ptr() for e in evals])
    vptr = (C.c_void_p * n)(*[(v.data_ptr() if v is not None else 0) for v in evecs])
    nbytes = lib.bk_eigh_workspace_bytes(dims, n)
    ws = ws or _Workspace()
    buf = ws.get(nbytes, dev)
    rc = lib.bk_eigh_batched(fptr, lds, wptr, vptr, dims, n, float(sym_scale), int(max_sweeps),
                             buf.data_ptr(), nbytes, _lib.stream_ptr())
    _lib.check(rc, "bk_eigh_batched")
    if rc > 0:
        raise RuntimeError(f"eigensolver did not converge for factor {rc - 1} (d={fs[rc - 1].shape[0]})")
    return evals, evecs


def get_eigenvectors(factors: Dict[Module, Sequence[Tensor]]) -> Dict[Module, Tuple[Tensor, Tensor]]:
    """layer -> (eigenvectors of xxt + xxt^T, eigenvectors of ggt + ggt^T), ascending eigenvalue order
    (models/utilities.py:144-159).  All factors of the model are solved in one batch."""
    layers = list(factors.keys())
    mats = [m for layer in layers for m in factors[layer]]
    _, vecs = eigh_factors(mats, sym_scale=1.0, vectors=True)
    return {layer: (vecs[2 * i], vecs[2 * i + 1]) for i, layer in enumerate(layers)}


def get_eigenvalues(factors: List[Sequence[Tensor]], verbose: bool = False) -> Tensor:
    """Concatenated eigenvalues of KFAC factor pairs (outer product of the two factors' eigenvalues,
    flattened) or of diagonal factors (the factor itself, flattened).  models/utilities.py:120-141."""
    del verbose
    pairs = [f for f in factors if len(f) == 2]
    mats = [m for f in pairs for m in f]
    vals, _ = eigh_factors(mats, sym_scale=0.5, vectors=False) if mats else ([], [])
    out, k = [], 0
    for f in factors:
        if len(f) == 2:
            xa, xg = vals[2 * k], vals[2 * k + 1]
            k += 1
            out.append(kron(xa.view(-1, 1), xg.view(1, -1)).contiguous().view(-1))
        else:
            out.append(f.contiguous().view(-1))
    return torch.cat(out) if out else torch.Tensor()
