"""Helpers of the reference's `models/utilities.py` that belong to the Kronecker-factored Laplace path
(paths relative to /root/reference):

    kron(a, b)                      models/utilities.py:387-409 (einsum Kronecker product)
    get_eigenvectors(factors)       models/utilities.py:144-159 (symeig of F + F^T per factor)
    get_eigenvalues(factors)        models/utilities.py:120-141 (ger of the factors' eigenvalues)

    accuracy, confidence, negative_log_likelihood, predictive_entropy,
    expected_calibration_error, calibration_curve, binned_kl_distance
                                    models/utilities.py:178-366 (calibration metrics)

`torch.symeig` no longer exists, so the reference versions crash on a current torch; these run the
batched one-sided Jacobi eigensolver of libbk_kfac.so (bk_eigh_batched).  The metrics take the class
probabilities as a device tensor (or numpy array) and reduce them on the GPU in one pass
(bk_calibration_rows + bk_binned_stats); only the per-bin tail (<= 256 numbers) is combined on the host.
They return what the reference returns (Python floats / numpy arrays).  Plotting / argparse helpers of
that file are out of scope."""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
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


# ------------------------------------------------------------------------------- calibration metrics
def _device_probs(probabilities) -> Tensor:
    p = torch.as_tensor(probabilities)
    if not p.is_cuda:
        p = p.cuda()
    p = p.float()
    if p.dim() != 2:
        raise ValueError("probabilities must be [n, classes]")
    return p if p.stride(1) == 1 else p.contiguous()


def calibration_rows(probabilities, labels=None) -> Dict[str, Tensor]:
    """One pass over probs [n, classes]: per-row confidence (max p), correctness of the arg-max (0/1),
    -log(p[label] + 1e-12), entropy of the row, arg-max — device tensors — plus `totals` = their four
    sums over rows (fp64, host).  Everything below is derived from this."""
    lib = _lib.load()
    _lib.require_device()
    p = _device_probs(probabilities)
    n, classes = p.shape
    lab = None
    if labels is not None:
        lab = torch.as_tensor(labels).to(p.device, torch.int64).contiguous()
        assert lab.shape == (n,)
    out = {k: torch.empty(n, device=p.device, dtype=torch.float32) for k in ("conf", "correct", "nll", "entropy")}
    out["pred"] = torch.empty(n, device=p.device, dtype=torch.int32)
    totals = torch.empty(4, device=p.device, dtype=torch.float64)
    _lib.check(lib.bk_calibration_rows(p.data_ptr(), p.stride(0), _lib.ptr(lab), n, classes,
                                       out["conf"].data_ptr(), out["correct"].data_ptr(), out["nll"].data_ptr(),
                                       out["entropy"].data_ptr(), out["pred"].data_ptr(), totals.data_ptr(),
                                       _lib.stream_ptr()), "bk_calibration_rows")
    out["totals"] = totals.cpu()
    out["n"] = n
    return out


def _binned(x: Tensor, w1: Optional[Tensor], w2: Optional[Tensor], edges: Tensor, mode: int) -> np.ndarray:
    """[3, nbins] fp64: count, sum w1, sum w2 per bin (bk_binned_stats; edges fp64 on the device)."""
    lib = _lib.load()
    nbins = edges.numel() - 1
    out = torch.empty(3, nbins, device=x.device, dtype=torch.float64)
    _lib.check(lib.bk_binned_stats(x.data_ptr(), _lib.ptr(w1), _lib.ptr(w2), x.numel(), edges.data_ptr(), nbins,
                                   mode, out.data_ptr(), _lib.stream_ptr()), "bk_binned_stats")
    return out.cpu().numpy()


def accuracy(probabilities, labels) -> float:
    """Top-1 accuracy in percent (models/utilities.py:178-189)."""
    rows = calibration_rows(probabilities, labels)
    return 100.0 * float(rows["totals"][0]) / rows["n"]


def confidence(probabilities, mean: bool = True) -> Union[float, np.ndarray]:
    """Maximum predicted class probability (models/utilities.py:220-233)."""
    rows = calibration_rows(probabilities)
    if mean:
        return float(rows["totals"][1]) / rows["n"]
    return rows["conf"].cpu().numpy()


def negative_log_likelihood(probabilities, labels) -> float:
    """-mean(log(p[label] + 1e-12)) (models/utilities.py:236-247)."""
    rows = calibration_rows(probabilities, labels)
    return float(rows["totals"][2]) / rows["n"]


def predictive_entropy(probabilities, mean: bool = False) -> Union[np.ndarray, float]:
    """H(y) = -sum_c y_c ln y_c per prediction, rows normalised as scipy.stats.entropy does
    (models/utilities.py:335-353)."""
    rows = calibration_rows(probabilities)
    if mean:
        return float(rows["totals"][3]) / rows["n"]
    return rows["entropy"].cpu().numpy()


def expected_calibration_error(probabilities, labels, bins: int = 10):
    """ECE over `bins` equally spaced confidence bins (lo, hi] (models/utilities.py:300-332): returns
    (ece, per-bin conf - acc, per-bin accuracy, per-bin confidence); empty bins report 0."""
    rows = calibration_rows(probabilities, labels)
    edges = torch.linspace(0, 1, bins + 1, dtype=torch.float64).to(rows["conf"].device)
    cnt, s_ok, s_conf = _binned(rows["conf"], rows["correct"], rows["conf"], edges, 0)
    n = rows["n"]
    bin_ace, bin_accuracy, bin_confidence = [], [], []
    ece = 0
    for i in range(bins):
        if cnt[i] > 0:
            acc, conf = s_ok[i] / cnt[i], s_conf[i] / cnt[i]
            ece += cnt[i] / n * abs(conf - acc)
            bin_ace.append(conf - acc)
            bin_accuracy.append(acc)
            bin_confidence.append(conf)
        else:
            bin_ace.append(0)
            bin_accuracy.append(0)
            bin_confidence.append(0)
    return ece, np.array(bin_ace), np.array(bin_accuracy), np.array(bin_confidence)


def calibration_curve(probabilities, labels, bins: int = 20):
    """ECE over equal-mass bins whose edges are every `step`-th sorted confidence, intervals open on
    both sides (models/utilities.py:250-297): (ece, mean confidence, accuracy, proportion) of the
    non-empty bins."""
    rows = calibration_rows(probabilities, labels)
    conf = rows["conf"]
    n = rows["n"]
    step = (n + bins - 1) // bins
    srt = torch.sort(conf).values
    edges = srt[::step]
    if n % step != 1:
        edges = torch.cat([edges, srt[-1:]])
    cnt, s_ok, s_conf = _binned(conf, rows["correct"], conf, edges.double().contiguous(), 1)
    ece = 0.0
    xs, ys, zs = [], [], []
    for i in range(cnt.shape[0]):
        if cnt[i] > 0:
            prop, acc, avg = cnt[i] / n, s_ok[i] / cnt[i], s_conf[i] / cnt[i]
            ece += abs(avg - acc) * prop
            xs.append(avg)
            ys.append(acc)
            zs.append(prop)
    return ece, np.array(xs), np.array(ys), np.array(zs)


def binned_kl_distance(dist1, dist2, smooth: float = 1e-7, bins=None) -> float:
    """Symmetrised discrete KL divergence between two sample sets, histogrammed over `bins`
    (default logspace(-7, 1, 200)) with additive smoothing (models/utilities.py:192-217)."""
    _lib.require_device()
    if bins is None:
        bins = np.logspace(-7, 1, num=200)
    edges = torch.as_tensor(np.asarray(bins, dtype=np.float64)).cuda()
    pdfs = []
    for d in (dist1, dist2):
        x = torch.as_tensor(d)
        x = (x if x.is_cuda else x.cuda()).float().contiguous().view(-1)
        h = _binned(x, None, None, edges, 2)[0] + smooth
        pdfs.append(h / h.sum())
    p, q = pdfs
    return float(np.sum(p * np.log(p / q)) + np.sum(q * np.log(q / p)))


# ------------------------------------------------------------------------------------------------------
# Script helpers of models/utilities.py:22-118, 366-386, 410-430 (the same functions are duplicated in
# sampling_free/utils.py:213-236) that the linearised-predictive and dense-Fisher scripts call.
def gradient(y: Tensor, x: Tensor, grad_outputs: Optional[Tensor] = None) -> Tensor:
    """dy/dx contracted with grad_outputs (ones by default), graph kept (models/utilities.py:29-34)."""
    if grad_outputs is None:
        grad_outputs = torch.ones_like(y)
    return torch.autograd.grad(y, [x], grad_outputs=grad_outputs, create_graph=True, retain_graph=True,
                               allow_unused=True)[0]


def jacobian(y: Tensor, x: Tensor, device=None) -> Tensor:
    """[classes, numel(x)]: row i = gradient of sum_b y[b, i] w.r.t. x (models/utilities.py:36-47, which runs one
    backward pass per class; here the classes are one batched backward, `is_grads_batched`, same values)."""
    n_cls = y.shape[1]
    go = torch.zeros((n_cls,) + tuple(y.shape), dtype=y.dtype, device=y.device)
    idx = torch.arange(n_cls, device=y.device)
    go[idx, :, idx] = 1
    try:
        g = torch.autograd.grad(y, [x], grad_outputs=go, retain_graph=True, allow_unused=True,
                                is_grads_batched=True)[0]
        jac = g.reshape(n_cls, -1)
    except RuntimeError:        # an op without a batching rule: the reference's loop
        jac = torch.stack([torch.flatten(gradient(y, x, go[i])) for i in range(n_cls)])
    return jac.detach().to(device if device is not None else y.device)


def get_near_psd(A: Tensor, epsilon: float) -> Tensor:
    """Nearest-PSD repair: eigenvalues of (A + A^T)/2 below epsilon are raised to epsilon
    (models/utilities.py:22-26).  The reference calls the GENERAL eigensolver on the symmetric part in fp64 and
    returns a complex tensor; the symmetric part has a real orthogonal eigenbasis, so this returns the real fp32
    matrix V max(w, eps) V^T from the Jacobi eigensolver (documented deviation: dtype)."""
    from .curvatures import mm_nt
    (w,), (v,) = eigh_factors([A], sym_scale=0.5)
    w = torch.where(w < epsilon, torch.full_like(w, float(epsilon)), w)
    return mm_nt((v * w.unsqueeze(0)).contiguous(), v)


def generate_kernel_coords() -> List[Tuple[int, int]]:
    """Per-kernel diagonal block ranges of BaseNet_15k (models/utilities.py:93-118, hessian/utils.py:67-95)."""
    from .dense import kernel_block_coords_basenet15k
    return kernel_block_coords_basenet15k()


def calculateDominance(H: Tensor, regParam: float = 0.00001) -> Tuple[float, float]:
    """(diagonal dominance, kernel-block dominance) of H + regParam I for BaseNet_15k's 15 080 x 15 080 Fisher
    (models/utilities.py:50-70; other sizes raise NotImplementedError like the reference), one pass on the device."""
    from .dense import dominance
    if H.numel() != 15080 ** 2:
        raise NotImplementedError
    return dominance(H, generate_kernel_coords(), regParam)


def ram() -> float:
    """Utilised system memory in percent (models/utilities.py:369-375)."""
    import psutil
    return psutil.virtual_memory()[2]


def vram() -> float:
    """Device memory allocated by this process in GB (models/utilities.py:378-384)."""
    return torch.cuda.memory_allocated() / (1024.0 ** 3)


def seed_all_rng(seed: Optional[int] = None) -> None:
    """Seeds torch, numpy and python RNGs (models/utilities.py:389-407).  The engine's own noise is Philox keyed by
    the estimator's `seed=` argument and does not depend on these global generators."""
    import os
    import random
    from datetime import datetime
    if seed is None:
        seed = os.getpid() + int(datetime.now().strftime("%S%f")) + int.from_bytes(os.urandom(2), "big")
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    random.seed(seed)
