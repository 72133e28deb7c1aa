"""Posterior predictive: Monte-Carlo over weight samples and the sampling-free linearised form.

The reference keeps these loops in scripts; they are first-class functions here, with semantics
pinned to the cited script lines (paths relative to /root/reference):

  mc_predict            sampling/classification_sampling.py:71-80 (mean of softmax over S samples),
                        sampling/regression_sampling.py:81-88 (per-input mean / std, ddof = 0);
                        each sample = Curvature.sample_and_replace + forward (models/wrapper.py:35-44)
  linearised_kfac       sampling_free/classification/classification_ll_block.py:114-135,
                        sampling_free/regression/regression_ll_block.py:120-140
  linearised_diag       sampling_free/classification/classification_ll_diagonal.py:104-131,
                        sampling_free/regression/regression_ll_diagonal.py:116-139

`mc_predict` never writes sampled weights back into the model: all S weight samples of a layer are
drawn in one batched launch, turned into bf16 GEMM operands by one fused kernel, and the S forward
passes run as ONE batched tcgen05 GEMM per Linear layer (bias + ReLU in the epilogue) / one SIMT
launch per conv layer.  Sample s always uses Philox subsequence s, so sharding samples over GPUs does
not change the result.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor
from torch.nn import Module

from . import _lib
from .curvatures import KFAC, Diagonal, _round8, gemm_precision, invert_factors, stage_operand


# ------------------------------------------------------------------------------------------------
@dataclass
class Op:
    kind: str                 # "conv" | "linear" | "flatten"
    layer: Optional[Module] = None
    relu: bool = False
    pool: bool = False


def program_for(model: Module) -> List[Op]:
    """Forward program of the supported architectures (reference CNNs, LeNet-5, MLP stacks)."""
    name = model.__class__.__name__
    if name == "BaseNet_15k":
        return [Op("conv", model.conv1, True, True), Op("conv", model.conv2, True, True), Op("flatten"),
                Op("linear", model.fc1, True), Op("linear", model.fc2, False)]
    if name == "BaseNet_750":
        return [Op("conv", model.conv1, True, True), Op("conv", model.conv2, True, True), Op("flatten"),
                Op("linear", model.fc1, False)]
    if name == "LeNet5":
        return [Op("conv", model.conv1, True, True), Op("conv", model.conv2, True, True), Op("flatten"),
                Op("linear", model.fc1, True), Op("linear", model.fc2, True), Op("linear", model.fc3, False)]
    if name == "MLP" and hasattr(model, "layers"):
        n = len(model.layers)
        return [Op("flatten")] + [Op("linear", l, i + 1 < n) for i, l in enumerate(model.layers)]
    # generic: attribute-ordered Linear layers fc1..fcK with ReLU between (reference regression nets)
    fcs = [m for m in model.children() if m.__class__.__name__ == "Linear"]
    if fcs and len(fcs) == len(list(model.children())):
        return [Op("linear", l, i + 1 < len(fcs)) for i, l in enumerate(fcs)]
    raise NotImplementedError(f"no forward program for {name}; pass program=[Op(...), ...]")


def _linear_forward(lib, act, act_shared: bool, rows: int, d_in: int, w_hi, w_lo, ldw: int, bias, S: int,
                    d_out: int, relu: bool, last: bool, prec: int):
    """act: (hi, lo, ld) bf16 staged [S or 1, rows, ld].  Returns fp32 [S, rows, d_out] when `last`,
    else the next layer's staged (hi, lo, ld)."""
    a_hi, a_lo, lda = act
    x3 = prec == _lib.BK_PREC_BF16X3
    dev = a_hi.device
    flags = _lib.GEMM_RELU if relu else 0
    if last:
        out = torch.empty(S, rows, d_out, device=dev, dtype=torch.float32)
        c_ptr, o_hi, o_lo, ldo = out.data_ptr(), None, None, 0
    else:
        ldo = _round8(d_out)
        o_hi = torch.zeros(S, rows, ldo, dtype=torch.bfloat16, device=dev)
        o_lo = torch.zeros_like(o_hi) if x3 else None
        out, c_ptr = None, 0
    _lib.check(lib.bk_gemm_nt(a_hi.data_ptr(), a_lo.data_ptr() if x3 else 0, lda,
                              0 if act_shared else rows * lda,
                              w_hi.data_ptr(), w_lo.data_ptr() if x3 else 0, ldw, d_out * ldw,
                              rows, d_out, d_in, S, prec, flags, 1.0, 0.0,
                              c_ptr, d_out, rows * d_out,
                              bias.data_ptr(), d_out,
                              _lib.ptr(o_hi), _lib.ptr(o_lo), ldo, rows * ldo, _lib.stream_ptr()),
               "bk_gemm_nt(forward)")
    return out if last else (o_hi, o_lo, ldo)


def _gemm(lib, a, sa, b, sb, m, n, k, batch, prec, flags=0, alpha=1.0, beta=0.0, c=None, ldc=0, sc=0,
          o=None, so=0, what="bk_gemm_nt"):
    """bk_gemm_nt on staged operands: a / b / o are (hi, lo, ld) triples, c an fp32 tensor."""
    x3 = prec == _lib.BK_PREC_BF16X3
    a_hi, a_lo, lda = a
    b_hi, b_lo, ldb = b
    o_hi, o_lo, ldo = o if o is not None else (None, None, 0)
    _lib.check(lib.bk_gemm_nt(a_hi.data_ptr(), a_lo.data_ptr() if x3 else 0, lda, sa,
                              b_hi.data_ptr(), b_lo.data_ptr() if x3 else 0, ldb, sb,
                              m, n, k, batch, prec, flags, alpha, beta,
                              _lib.ptr(c), ldc, sc, 0, 0,
                              _lib.ptr(o_hi), _lib.ptr(o_lo) if x3 else 0, ldo, so, _lib.stream_ptr()), what)


def _stage_activation(lib, x2d: Tensor, x3: bool, ones_col: bool):
    """fp32 [rows, d] -> bf16 (hi, lo, ld) with an optional trailing column of ones (the reference's
    bias augmentation, models/curvatures.py:346-348, on the activation side)."""
    rows, d = x2d.shape
    ld = _round8(d + int(ones_col))
    hi = torch.zeros(rows, ld, dtype=torch.bfloat16, device=x2d.device)
    lo = torch.zeros_like(hi) if x3 else hi
    xc = x2d.float().contiguous()
    _lib.check(lib.bk_convert_split(xc.data_ptr(), xc.stride(0), rows, d, 1.0, 0, hi.data_ptr(),
                                    lo.data_ptr() if x3 else 0, ld, _lib.stream_ptr()), "bk_convert_split")
    if ones_col:
        hi[:, d] = 1.0
    return hi, lo, ld


def _ptr(t: Optional[Tensor], off_elems: int = 0) -> int:
    return 0 if t is None else t.data_ptr() + 2 * off_elems


def _gemm_raw(lib, a_hi, a_lo, lda, sa, b_hi, b_lo, ldb, sb, m, n, k, batch, prec, flags, c, ldc, sc,
              o_hi, o_lo, ldo, so, what, beta=0.0):
    x3 = prec == _lib.BK_PREC_BF16X3
    _lib.check(lib.bk_gemm_nt(a_hi, a_lo if x3 else 0, lda, sa, b_hi, b_lo if x3 else 0, ldb, sb,
                              m, n, k, batch, prec, flags, 1.0, beta, c, ldc, sc, 0, 0,
                              o_hi, o_lo if x3 else 0, ldo, so, _lib.stream_ptr()), what)


def _new_cat(S: int, B: int, kx: int, d_out: int, x3: bool, dev):
    """K-concatenated activation buffer [S, B, kx + round8(d_out)] of a Linear layer (zero-filled:
    padding columns must stay zero)."""
    ldcat = kx + _round8(d_out)
    hi = torch.empty(S, B, ldcat, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi) if x3 else None
    for t in (hi, lo):
        if t is not None and ldcat != kx + d_out:
            t[:, :, kx + d_out:].zero_()
    return hi, lo, ldcat


def _implicit_linear(lib, est: KFAC, layer: Module, li: int, cat, x_shared: bool, B: int, S: int,
                     sample0: int, z: Optional[Tensor], relu: bool, last: bool, prec: int, nxt,
                     prefetched=None):
    """One Linear layer of the MC forward WITHOUT materialising the sampled weights.

    With x~ = [x, 1] and the sample (L_A Z L_G^T)^T added to M~ = [W | b] (models/curvatures.py:
    67-82, 403-405):   x~ W_s~^T = x~ M~^T + ((x~ L_A) Z_s) L_G^T  =  [x~ | Y2_s] [M~ | L_G]^T.
    Three tensor-core contractions per layer with the B test inputs as the row dimension (~6 B d^2
    flops instead of the 2 d^3 of forming W_s), the same numbers for the same Z_s:
        Y1   = x~ L_A[:, :d_in]       (L_A^T upper: leading zero k-blocks skipped; the last column of
                                       x~ L_A is the constant L_A[d_in, d_in] and is filled, not computed)
        Y2_s = Y1 Z_s                 (Z_s^T comes straight from the Philox kernel in operand layout and
                                       lands in the right half of the K-concatenated buffer)
        out  = act([x~ | Y2_s] [M~ | L_G]^T)   (one GEMM over K = d_in' + d_out, bias through the ones
                                       column, ReLU and the next layer's bf16 operand in the epilogue)
    cat: (hi, lo, ldcat) buffer [S, B, ldcat] whose left part holds x~_s (ones column included).
    nxt: the next layer's cat buffer (the epilogue writes activations into its left part) or None."""
    x3 = prec == _lib.BK_PREC_BF16X3
    LA, LG = est.inv_state[layer]
    dinp, dout = LA.shape[0], LG.shape[0]
    d_in = dinp - 1
    dev = LA.device
    ops = est._implicit_operands(layer)
    kx = ops["kx"]
    c_hi, c_lo, ldcat = cat
    Sx = 1 if x_shared else S
    # 1. Y1 = x~ L_A
    ld1 = kx
    y1_hi = torch.empty(Sx, B, ld1, dtype=torch.bfloat16, device=dev)
    y1_lo = torch.empty_like(y1_hi) if x3 else None
    if ld1 != dinp:
        y1_hi[:, :, dinp:].zero_()
        if x3:
            y1_lo[:, :, dinp:].zero_()
    t_hi, t_lo, ldt = ops["LAT"]
    _gemm_raw(lib, _ptr(c_hi), _ptr(c_lo), ldcat, B * ldcat, _ptr(t_hi), _ptr(t_lo), ldt, 0,
              B, d_in, dinp, Sx, prec, _lib.GEMM_TRI_B_UPPER, 0, 0, 0,
              _ptr(y1_hi), _ptr(y1_lo), ld1, B * ld1, "bk_gemm_nt(x L_A)")
    dd = ops["la_dd"]
    dd_hi = dd.to(torch.bfloat16)
    y1_hi[:, :, d_in] = dd_hi
    if x3:
        y1_lo[:, :, d_in] = (dd - dd_hi.float()).to(torch.bfloat16)
    # 2. Y2_s = Y1 Z_s  ->  right half of the concatenated buffer
    ldz = kx
    pre = None if prefetched is None else prefetched.get(li)
    if pre is not None:
        zt_hi, zt_lo, ev = pre
        torch.cuda.current_stream().wait_event(ev)
    else:
        zt_hi = torch.empty(S, dout, ldz, dtype=torch.bfloat16, device=dev)
        zt_lo = torch.empty_like(zt_hi) if x3 else None
    if pre is not None:
        pass
    elif z is None:
        _lib.check(lib.bk_philox_normal(est.seed, sample0, li, dout, dinp, S, 0, 0, 0, zt_hi.data_ptr(),
                                        _lib.ptr(zt_lo), ldz, dout * ldz, _lib.stream_ptr()), "bk_philox_normal")
    else:
        zz = z.to(dev, torch.float32).contiguous()
        assert zz.shape == (S, dinp, dout), "noise must be [S, d_in', d_out]"
        if ldz != dinp:
            zt_hi.zero_()
            if x3:
                zt_lo.zero_()
        for s_ in range(S):
            _lib.check(lib.bk_transpose_split(zz[s_].data_ptr(), dout, dinp, dout, 1.0, 0, zt_hi[s_].data_ptr(),
                                              zt_lo[s_].data_ptr() if x3 else 0, ldz, _lib.stream_ptr()),
                       "bk_transpose_split")
    _gemm_raw(lib, _ptr(y1_hi), _ptr(y1_lo), ld1, 0 if x_shared else B * ld1,
              _ptr(zt_hi), _ptr(zt_lo), ldz, dout * ldz, B, dout, dinp, S, prec, 0, 0, 0, 0,
              _ptr(c_hi, kx), _ptr(c_lo, kx), ldcat, B * ldcat, "bk_gemm_nt(Y1 Z)")
    # 3. out_s = act([x~ | Y2_s] [M~ | L_G]^T)
    m_hi, m_lo, ldm = ops["MLG"]
    flags = _lib.GEMM_TRI_B | _lib.gemm_tri_koff(kx) | (_lib.GEMM_RELU if relu else 0)
    K = kx + dout
    if last:
        ldc = (dout + 3) // 4 * 4
        out = torch.empty(S, B, ldc, device=dev, dtype=torch.float32)
        _gemm_raw(lib, _ptr(c_hi), _ptr(c_lo), ldcat, B * ldcat, _ptr(m_hi), _ptr(m_lo), ldm, 0,
                  B, dout, K, S, prec, flags, out.data_ptr(), ldc, B * ldc, 0, 0, 0, 0,
                  "bk_gemm_nt([x|Y2][M|L_G]^T)")
        return out[:, :, :dout]
    n_hi, n_lo, ldn = nxt
    _gemm_raw(lib, _ptr(c_hi), _ptr(c_lo), ldcat, B * ldcat, _ptr(m_hi), _ptr(m_lo), ldm, 0,
              B, dout, K, S, prec, flags, 0, 0, 0, _ptr(n_hi), _ptr(n_lo), ldn, B * ldn,
              "bk_gemm_nt([x|Y2][M|L_G]^T)")
    # ones column of the next layer's x~ (hi = 1, lo = 0) and zero padding up to its kx
    kxn = _round8(dout + 1)
    n_hi[:, :, dout:kxn].zero_()
    n_hi[:, :, dout] = 1.0
    if n_lo is not None:
        n_lo[:, :, dout:kxn].zero_()
    return nxt


_noise_stream = {}


def _prefetch_noise(lib, est: KFAC, prog, layer_index, B, S, sample0, x3, implicit, dev):
    """Philox noise operands Z_s^T of every implicit Linear layer, generated on a side stream so that
    the (SIMT, shared-memory-free) generator runs underneath the tensor-core GEMMs of earlier layers.
    The noise does not depend on activations: element (o, i) of sample s of layer l is a pure function
    of (seed, sample id, layer id, o, i).  Returns {layer index: (z_hi, z_lo, ready event)}."""
    main = torch.cuda.current_stream()
    side = _noise_stream.get(dev)
    if side is None:
        side = _noise_stream[dev] = torch.cuda.Stream(device=dev)
    out = {}
    side.wait_stream(main)
    for op in prog:
        if op.kind != "linear" or op.layer.bias is None:
            continue
        LA, LG = est.inv_state[op.layer]
        dinp, dout = LA.shape[0], LG.shape[0]
        use = implicit
        if use is None:
            use = 2 * B <= min(dinp, dout) and min(dinp - 1, dout) >= 700
        if not use:
            continue
        li = layer_index[op.layer]
        ldz = _round8(dinp)
        with torch.cuda.stream(side):
            z_hi = torch.empty(S, dout, ldz, dtype=torch.bfloat16, device=dev)
            z_lo = torch.empty_like(z_hi) if x3 else None
            _lib.check(lib.bk_philox_normal(est.seed, sample0, li, dout, dinp, S, 0, 0, 0, z_hi.data_ptr(),
                                            _lib.ptr(z_lo), ldz, dout * ldz, side.cuda_stream), "bk_philox_normal")
            ev = torch.cuda.Event()
            ev.record(side)
        z_hi.record_stream(main)
        if z_lo is not None:
            z_lo.record_stream(main)
        out[li] = (z_hi, z_lo, ev)
    return out


def _mc_logits_chunk(est: KFAC, x: Tensor, S: int, sample0: int, prog, noise, implicit: Optional[bool]):
    lib = _lib.load()
    prec = gemm_precision(est.precision)
    x3 = prec == _lib.BK_PREC_BF16X3
    st = _lib.stream_ptr()
    dev = x.device
    layer_index = {l: i for i, l in est._selected_layers()}
    cur = x.float().contiguous()       # fp32 activation, [B, ...] (shared) or [S, B, ...]
    shared = True
    staged = None                      # bf16 operand of the next Linear (hi, lo, ld)
    staged_ones = False                # does `staged` carry the trailing ones column?
    cat = None                         # K-concatenated buffer of the next implicit Linear layer
    B = x.shape[0]
    prefetched = _prefetch_noise(lib, est, prog, layer_index, B, S, sample0, x3, implicit, dev) if noise is None else None
    for oi, op in enumerate(prog):
        if op.kind == "flatten":
            if staged is None:
                cur = cur.reshape(cur.shape[0], -1) if shared else cur.reshape(S, B, -1)
            continue
        layer = op.layer
        li = layer_index[layer]
        z = None if noise is None else noise[li]
        has_bias = layer.bias is not None
        LA, LG = est.inv_state[layer]
        d_out = LG.shape[0]
        d_in = LA.shape[0] - int(has_bias)
        last = all(o.kind == "flatten" for o in prog[oi + 1:])
        use_implicit = implicit
        if use_implicit is None:   # forming W_s costs ~2 d^3 per sample, the implicit form ~6 B d^2
            use_implicit = 2 * B <= min(d_in + 1, d_out) and min(d_in, d_out) >= 700   # measured crossover
        if op.kind == "linear" and use_implicit and has_bias:
            kx = _round8(d_in + 1)
            if cat is None:
                # first implicit layer: stage x~ once (fp32 activations, shared or per sample)
                cat = _new_cat(S, B, kx, d_out, x3, dev)
                if staged is not None:   # bf16 operand left by a materialised Linear layer
                    s_hi, s_lo, s_ld = staged
                    rows = s_hi.numel() // (B * s_ld)
                    cat[0][:, :, :d_in].copy_(s_hi.view(rows, B, s_ld)[:, :, :d_in].expand(S, B, d_in))
                    if x3:
                        cat[1][:, :, :d_in].copy_(s_lo.view(rows, B, s_ld)[:, :, :d_in].expand(S, B, d_in))
                    # ones column (hi = 1, lo = 0) and the zero padding up to kx: all inside the K range
                    # of the Y1 and output GEMMs (_new_cat only clears the columns beyond kx + d_out)
                    cat[0][:, :, d_in:kx].zero_()
                    if x3:
                        cat[1][:, :, d_in:kx].zero_()
                    cat[0][:, :, d_in] = 1.0
                    shared = shared and rows == 1
                else:
                    flat = cur.reshape(-1, d_in)
                    hi, lo, ld = _stage_activation(lib, flat, x3, True)
                    rows = flat.shape[0] // B
                    cat[0][:, :, :ld].copy_(hi.view(rows, B, ld).expand(S, B, ld))
                    if x3:
                        cat[1][:, :, :ld].copy_(lo.view(rows, B, ld).expand(S, B, ld))
            # the epilogue of this layer writes straight into the next implicit layer's buffer
            nxt = None
            if not last:
                nl = next(o.layer for o in prog[oi + 1:] if o.kind != "flatten")
                nxt = _new_cat(S, B, _round8(d_out + 1), est.inv_state[nl][1].shape[0], x3, dev)
            res = _implicit_linear(lib, est, layer, li, cat, shared, B, S, sample0, z, op.relu, last, prec,
                                   nxt, prefetched)
            if last:
                return res
            # does the next layer also run implicitly?  otherwise hand over a plain staged activation
            cat, shared = res, False
            staged, staged_ones = (cat[0], cat[1] if x3 else cat[0], cat[2]), True
            continue
        cat = None
        smp = est.sample_batch(layer, S, z=z, sample0=sample0)        # [S, d_out, d_in']
        mean_w = layer.weight.detach().float().reshape(d_out, d_in).contiguous()
        mean_b = layer.bias.detach().float().contiguous() if has_bias else None
        b_s = torch.zeros(S, d_out, device=dev, dtype=torch.float32)
        if op.kind == "conv":
            w_s = torch.empty(S, d_out, d_in, device=dev, dtype=torch.float32)
            _lib.check(lib.bk_sample_to_weights(smp.data_ptr(), mean_w.data_ptr(), _lib.ptr(mean_b), d_out,
                                                d_in, int(has_bias), S, w_s.data_ptr(), 0, 0, 0,
                                                b_s.data_ptr(), st), "bk_sample_to_weights")
            n, c, h, w = (cur.shape if shared else cur.shape[1:])
            kh, kw = layer.kernel_size
            sh, sw = layer.stride
            ph, pw = layer.padding
            oh = (h + 2 * ph - kh) // sh + 1
            ow = (w + 2 * pw - kw) // sw + 1
            qh, qw = (oh // 2, ow // 2) if op.pool else (oh, ow)
            out = torch.empty(S, n, d_out, qh, qw, device=dev, dtype=torch.float32)
            _lib.check(lib.bk_conv2d_relu_pool(cur.data_ptr(), 0 if shared else n * c * h * w, w_s.data_ptr(),
                                               b_s.data_ptr(), out.data_ptr(), S, n, c, h, w, d_out, kh, kw,
                                               sh, sw, ph, pw, int(op.relu), int(op.pool), st),
                       "bk_conv2d_relu_pool")
            cur, shared = out, False
            continue
        # ---- linear, materialised weights
        ldw = _round8(d_in)
        w_hi = torch.zeros(S, d_out, ldw, dtype=torch.bfloat16, device=dev)
        w_lo = torch.zeros_like(w_hi) if x3 else None
        _lib.check(lib.bk_sample_to_weights(smp.data_ptr(), mean_w.data_ptr(), _lib.ptr(mean_b), d_out, d_in,
                                            int(has_bias), S, 0, w_hi.data_ptr(), _lib.ptr(w_lo), ldw,
                                            b_s.data_ptr(), st), "bk_sample_to_weights")
        if staged is None:
            flat = cur.reshape(-1, d_in)                              # [B, d_in] or [S*B, d_in]
            hi, lo, ld = stage_operand(flat)
            staged = (hi, lo if x3 else hi, ld)
            staged_ones = False
        res = _linear_forward(lib, staged, shared, B, d_in, w_hi, w_lo, ldw, b_s, S, d_out, op.relu, last,
                              prec)
        if last:
            return res
        staged, shared, staged_ones = res, False, False
    raise RuntimeError("forward program must end with a Linear layer")


def mc_logits(est: KFAC, x: Tensor, n_samples: int, sample0: int = 0,
              program: Optional[Sequence[Op]] = None, noise: Optional[Sequence[Tensor]] = None,
              implicit: Optional[bool] = None, max_chunk_bytes: int = 6 << 30) -> Tensor:
    """Outputs of the network under `n_samples` posterior weight samples: [S, B, C] fp32.

    noise: optional per-layer external noise, noise[layer_index] = [S, d_in', d_out] (parity mode).
    implicit: None = choose per layer (see _implicit_linear), True / False = force.
    Samples are processed in chunks whose noise operands fit `max_chunk_bytes`."""
    prog = list(program) if program is not None else program_for(est.model)
    x3 = gemm_precision(est.precision) == _lib.BK_PREC_BF16X3
    per_sample = max(int(est.inv_state[op.layer][0].shape[0]) * int(est.inv_state[op.layer][1].shape[0])
                     for op in prog if op.layer is not None) * (4 if x3 else 2) * 3
    chunk = max(1, min(n_samples, max_chunk_bytes // max(per_sample, 1)))
    outs = []
    for s0 in range(0, n_samples, chunk):
        sc = min(chunk, n_samples - s0)
        nz = None if noise is None else [None if t is None else t[s0:s0 + sc] for t in noise]
        outs.append(_mc_logits_chunk(est, x, sc, sample0 + s0, prog, nz, implicit))
    return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)


@_lib.nvtx_range("mc_moments")
def mc_moments(est: KFAC, x: Tensor, n_samples: int, sample0: int = 0, mode: str = "classification",
               program=None, noise=None) -> Tuple[Tensor, Tensor]:
    """(E_s[p], E_s[p^2]) over `n_samples` samples starting at global sample id `sample0`;
    p = softmax(logits) (classification) or the raw output (regression).  mode "regression_centred"
    returns (E_s[y], E_s[(y - E_s[y])^2]): the variance directly, from a two-pass kernel (E[y^2] - E[y]^2
    in fp32 cancels catastrophically when |mean| >> std, as on the reference's y = x^3 task)."""
    lib = _lib.load()
    logits = mc_logits(est, x, n_samples, sample0, program, noise).contiguous()
    S, B, Cn = logits.shape
    mean = torch.empty(B, Cn, device=x.device, dtype=torch.float32)
    meansq = torch.empty_like(mean)
    kmode = {"classification": 0, "regression": 1, "regression_centred": 2}[mode]
    _lib.check(lib.bk_predictive_moments(logits.data_ptr(), S, B, Cn, kmode,
                                         mean.data_ptr(), meansq.data_ptr(), _lib.stream_ptr()),
               "bk_predictive_moments")
    return mean, meansq


def mc_predict(est: KFAC, x: Tensor, n_samples: int = 30, mode: str = "classification", program=None,
               noise=None, sample0: int = 0):
    """classification: mean softmax [B, C].  regression: (mean [B], std [B]) with ddof = 0."""
    if mode == "classification":
        return mc_moments(est, x, n_samples, sample0, mode, program, noise)[0]
    mean, var = mc_moments(est, x, n_samples, sample0, "regression_centred", program, noise)
    return mean.squeeze(1), var.sqrt().squeeze(1)


# ------------------------------------------------------------------------------------------------
def kron_quadform(V: Tensor, Q: Tensor, H: Tensor, *, precision: int = _lib.BK_PREC_BF16X3,
                  triangular: bool = False, out: Optional[Tensor] = None, accumulate: bool = False,
                  staged=None) -> Tensor:
    """|<V_b, Q V_b H^T>| for a batch of V_b [d_in', d_out]  ==  |J_b (Q (x) H) J_b^T| with
    V_b = J_b.view(d_in', d_out) — the Kronecker product is never formed.
    `triangular=True`: Q and H are lower-triangular (Cholesky factors, reference quirk Q1) and the
    zero k-blocks are skipped.  `staged` = ((q_hi, q_lo, ldq), (h_hi, h_lo, ldh)): bf16 operands of Q and H
    staged once by the caller (they are constants of an inverted estimator)."""
    lib = _lib.load()
    st = _lib.stream_ptr()
    x3 = precision == _lib.BK_PREC_BF16X3
    Bn, dinp, dout = V.shape
    dev = V.device
    V = V.float().contiguous()
    if staged is None:
        staged = (stage_operand(Q, lower_only=triangular), stage_operand(H, lower_only=triangular))
    (q_hi, q_lo, ldq), (h_hi, h_lo, ldh) = staged
    ldv = _round8(dinp)
    # all V_b^T in ONE launch: rows padded to ldv, [Bn * ldv, dout] -> [dout, Bn * ldv] = Bn adjacent column slabs,
    # i.e. a [dout, Bn, ldv] buffer, permuted to the batched K-major operand [Bn, dout, ldv] by a strided copy
    if ldv != dinp:
        Vp = torch.zeros(Bn, ldv, dout, device=dev, dtype=torch.float32)
        Vp[:, :dinp].copy_(V)
    else:
        Vp = V
    t_hi = torch.empty(dout, Bn * ldv, dtype=torch.bfloat16, device=dev)
    t_lo = torch.empty_like(t_hi)
    _lib.check(lib.bk_transpose_split(Vp.data_ptr(), dout, Bn * ldv, dout, 1.0, 0, t_hi.data_ptr(),
                                      t_lo.data_ptr(), Bn * ldv, st), "bk_transpose_split")
    vt_hi = t_hi.view(dout, Bn, ldv).permute(1, 0, 2).contiguous()
    vt_lo = t_lo.view(dout, Bn, ldv).permute(1, 0, 2).contiguous()
    ldu = _round8(dout)
    u_hi = torch.zeros(Bn, dinp, ldu, dtype=torch.bfloat16, device=dev)
    u_lo = torch.zeros_like(u_hi)
    # U_b = Q V_b
    _lib.check(lib.bk_gemm_nt(q_hi.data_ptr(), q_lo.data_ptr() if x3 else 0, ldq, 0,
                              vt_hi.data_ptr(), vt_lo.data_ptr() if x3 else 0, ldv, dout * ldv,
                              dinp, dout, dinp, Bn, precision, _lib.GEMM_TRI_A if triangular else 0, 1.0, 0.0,
                              0, 0, 0, 0, 0, u_hi.data_ptr(), u_lo.data_ptr() if x3 else 0, ldu, dinp * ldu,
                              st), "bk_gemm_nt(QV)")
    # W_b = U_b H^T
    Wm = torch.empty(Bn, dinp, dout, device=dev, dtype=torch.float32)
    _lib.check(lib.bk_gemm_nt(u_hi.data_ptr(), u_lo.data_ptr() if x3 else 0, ldu, dinp * ldu,
                              h_hi.data_ptr(), h_lo.data_ptr() if x3 else 0, ldh, 0,
                              dinp, dout, dout, Bn, precision, _lib.GEMM_TRI_B if triangular else 0, 1.0, 0.0,
                              Wm.data_ptr(), dout, dinp * dout, 0, 0, 0, 0, 0, 0, st), "bk_gemm_nt(UH^T)")
    if out is None:
        out = torch.zeros(Bn, device=dev, dtype=torch.float32)
        accumulate = False
    _lib.check(lib.bk_frob_dot(out.data_ptr(), V.data_ptr(), dinp * dout, Wm.data_ptr(), dinp * dout,
                               dinp * dout, Bn, 1, int(accumulate), st), "bk_frob_dot")
    return out


def _layer_jacobian(out: Tensor, layer: Module, grad_outputs: Optional[Tensor]) -> Tensor:
    """J_i = cat(flatten(d out / d p) for p in layer.parameters()) (classification_ll_block.py:128-130).
    Autograd is host plumbing exactly as in the reference script."""
    g = [torch.flatten(torch.autograd.grad(out, [p], grad_outputs=grad_outputs, retain_graph=True,
                                           allow_unused=True)[0]) for p in layer.parameters()]
    return torch.cat(g, dim=0)


def _layers_jacobians(out: Tensor, layers, grad_outputs: Optional[Tensor]) -> list:
    """The J_i of several layers from ONE backward pass (the reference runs one `autograd.grad` per parameter,
    classification_ll_block.py:128-130: same values, 2 x len(layers) backward passes)."""
    params = [p for l in layers for p in l.parameters()]
    grads = torch.autograd.grad(out, params, grad_outputs=grad_outputs, retain_graph=True, allow_unused=True)
    res, k = [], 0
    for l in layers:
        n = len(list(l.parameters()))
        res.append(torch.cat([torch.flatten(g) for g in grads[k:k + n]], dim=0).detach())
        k += n
    return res


def _layers_jacobians_per_output(preds: Tensor, layers) -> list:
    """Per-test-point Jacobians J[j] = d preds[j] / d params of every layer, [P, n_params_of_layer] each:
    one batched backward (`is_grads_batched`) instead of the script's P x 2 x len(layers) backward passes
    (regression_ll_block.py:131-134); falls back to the per-point loop where autograd cannot batch."""
    P = preds.shape[0]
    params = [p for l in layers for p in l.parameters()]
    try:
        eye = torch.eye(P, device=preds.device, dtype=preds.dtype).reshape((P,) + tuple(preds.shape))
        grads = torch.autograd.grad(preds, params, grad_outputs=eye, retain_graph=True, allow_unused=True,
                                    is_grads_batched=True)
        res, k = [], 0
        for l in layers:
            n = len(list(l.parameters()))
            res.append(torch.cat([g.reshape(P, -1) for g in grads[k:k + n]], dim=1).detach())
            k += n
        return res
    except RuntimeError:
        return [torch.stack([_layer_jacobian(preds[j], l, torch.ones_like(preds[j])).detach() for j in range(P)])
                for l in layers]


def argmax_grad_outputs(pred_mean: Tensor) -> Tensor:
    """grad_outputs[:, idx] = 1 with a vector idx (reference quirk Q3, classification_ll_block.py:119-121)."""
    idx = torch.argmax(pred_mean.detach(), dim=1)
    go = torch.zeros_like(pred_mean)
    go[:, idx] = 1
    return go


@_lib.nvtx_range("linearised_kfac_classification")
def linearised_kfac_classification(est: KFAC, x: Tensor) -> Tuple[Tensor, float, float]:
    """One test batch of the sampling-free KFAC predictive: (pred_mean [B, C], pred_std, entropy).
    classification_ll_block.py:114-135, including its quirks (Cholesky factors used as Q_i / H_i, flat
    Jacobian reinterpreted row-major, gradient summed over the batch)."""
    pred_mean = torch.softmax(est.model(x), dim=1)
    go = argmax_grad_outputs(pred_mean)
    total = torch.zeros(1, device=x.device, dtype=torch.float32)
    prec = gemm_precision(est.precision)
    layers = [layer for layer in list(est.model.modules())[1:] if layer in est.state]
    for layer, J_i in zip(layers, _layers_jacobians(pred_mean, layers, go)):
        Q_i, H_i = est.inv_state[layer]
        V = J_i.reshape(1, Q_i.shape[0], H_i.shape[0])
        kron_quadform(V, Q_i, H_i, precision=prec, triangular=True, out=total, accumulate=True,
                      staged=est._staged_factors(layer))
    pred_std = float(total.item())
    entropy = 0.5 * np.log2(2 * np.e * np.pi * pred_std)
    return pred_mean.detach(), pred_std, float(entropy)


@_lib.nvtx_range("linearised_kfac_regression")
def linearised_kfac_regression(est: KFAC, x_test: Tensor, tau: float, N: float, sigma: float) -> Tensor:
    """Predictive std per test point: sqrt(sum_layers |J (q_inv (x) h_inv) J^T|) + sigma with
    q_inv = (N (A + tau I))^-1, h_inv = (N (G + tau I))^-1 taken from `state`.
    regression_ll_block.py:120-140.  The inverses are computed ONCE (the script recomputes them per
    test point) as L L^T from the batched Cholesky kernel (R = N F + N tau I is SPD, so the
    pseudo-inverse of the script is the inverse)."""
    lib = _lib.load()
    layers = [l for l in list(est.model.modules())[1:] if l in est.state]
    dims = [(est.state[l][0].shape[0], est.state[l][1].shape[0]) for l in layers]
    preds = est.model(x_test)
    P = preds.shape[0]
    total = torch.zeros(P, device=x_test.device, dtype=torch.float32)
    jacs = _layers_jacobians_per_output(preds, layers)
    if all(max(a, b) <= _lib.BK_SMALL64_MAX_DIM and a * b <= _lib.BK_SMALL64_MAX_ELEMS for a, b in dims):
        # every factor fits one CTA: fp64 inverse + fp64 quadratic form (bk_small64.cu).  cond(N (F + tau I))
        # reaches 1e5..5e6 on this problem; fp32 products resolve the result to ~1e-2 only.
        invs = spd_inverse_f64([f for l in layers for f in est.state[l]], float(N) * tau, float(N))
        st = _lib.stream_ptr()
        for i, (l, J) in enumerate(zip(layers, jacs)):
            dinp, dout = dims[i]
            V = J.float().contiguous()
            _lib.check(lib.bk_kron_quadform_f64(V.data_ptr(), dinp * dout, P, dinp, dout, invs[2 * i].data_ptr(),
                                                invs[2 * i + 1].data_ptr(), total.data_ptr(), 1, st),
                       "bk_kron_quadform_f64")
        return total.sqrt() + sigma
    factors, adds, mults = [], [], []
    for l in layers:
        factors += list(est.state[l])
        adds += [(N * tau) ** 2] * 2
        mults += [float(N) ** 2] * 2
    chol = invert_factors(factors, adds, mults, est._ws)
    invs = [inverse_from_chol(Lc) for Lc in chol]
    prec = gemm_precision(est.precision)
    for i, (l, J) in enumerate(zip(layers, jacs)):
        q_inv, h_inv = invs[2 * i], invs[2 * i + 1]
        V = J.reshape(P, q_inv.shape[0], h_inv.shape[0])
        kron_quadform(V, q_inv, h_inv, precision=prec, out=total, accumulate=True)
    return total.sqrt() + sigma


def spd_inverse_f64(factors: Sequence[Tensor], add: float, multiply: float) -> List[Tensor]:
    """(multiply * sym(F) + add * I)^-1 in fp64 for small fp32 factors (d <= BK_SMALL64_MAX_DIM), one CTA each.
    Raises RuntimeError naming the first factor that is not positive definite."""
    import ctypes as C
    lib = _lib.load()
    outs: List[Tensor] = []
    st = _lib.stream_ptr()
    fs = [f.float() if f.stride(-1) == 1 else f.float().contiguous() for f in factors]
    status = torch.zeros(1, dtype=torch.int32, device=fs[0].device)
    for g0 in range(0, len(fs), _lib.BK_SMALL64_MAX_BATCH):
        grp = fs[g0:g0 + _lib.BK_SMALL64_MAX_BATCH]
        n = len(grp)
        res = [torch.empty(f.shape[0], f.shape[0], dtype=torch.float64, device=f.device) for f in grp]
        _lib.check(lib.bk_spd_inverse_f64((C.c_void_p * n)(*[f.data_ptr() for f in grp]),
                                          (C.c_longlong * n)(*[f.stride(0) for f in grp]),
                                          (C.c_int * n)(*[f.shape[0] for f in grp]),
                                          (C.c_double * n)(*[add] * n), (C.c_double * n)(*[multiply] * n),
                                          (C.c_void_p * n)(*[r.data_ptr() for r in res]), n,
                                          status.data_ptr(), st), "bk_spd_inverse_f64")
        code = int(status.item())
        if code:
            raise RuntimeError(f"factor {g0 + code // 65536} is not positive definite after damping "
                               f"(add={add}, multiply={multiply})")
        outs += res
    return outs


def inverse_from_chol(Lc: Tensor, precision: int = _lib.BK_PREC_BF16X3) -> Tensor:
    """R^-1 = L L^T from the Cholesky factor of the inverse (one triangular-aware tensor-core GEMM)."""
    lib = _lib.load()
    d = Lc.shape[0]
    hi, lo, ld = stage_operand(Lc, lower_only=True)
    out = torch.empty(d, d, device=Lc.device, dtype=torch.float32)
    x3 = precision == _lib.BK_PREC_BF16X3
    _lib.check(lib.bk_gemm_nt(hi.data_ptr(), lo.data_ptr() if x3 else 0, ld, 0,
                              hi.data_ptr(), lo.data_ptr() if x3 else 0, ld, 0,
                              d, d, d, 1, precision, _lib.GEMM_TRI_A | _lib.GEMM_TRI_B, 1.0, 0.0,
                              out.data_ptr(), d, 0, 0, 0, 0, 0, 0, 0, _lib.stream_ptr()), "bk_gemm_nt(LL^T)")
    return out


def diag_flat_inverse(est: Diagonal) -> Tensor:
    """h = cat(flatten(inv_state[layer])) in model.modules() order (classification_ll_diagonal.py:108-113)."""
    return torch.cat([torch.flatten(est.inv_state[l]) for l in list(est.model.modules())[1:]
                      if l in est.state], dim=0)


def linearised_diag(est: Diagonal, J: Tensor, h: Optional[Tensor] = None) -> Tensor:
    """var_b = sum_j J[b, j]^2 h_j for Jacobian rows J [B, P] over net.parameters() order
    (classification_ll_diagonal.py:127-131; regression_ll_diagonal.py:135-139)."""
    lib = _lib.load()
    h = diag_flat_inverse(est) if h is None else h
    J = J.float().contiguous()
    out = torch.empty(J.shape[0], device=J.device, dtype=torch.float32)
    _lib.check(lib.bk_diag_quadform(out.data_ptr(), J.data_ptr(), J.stride(0), h.data_ptr(), h.numel(),
                                    J.shape[0], _lib.stream_ptr()), "bk_diag_quadform")
    return out


def params_jacobian(out: Tensor, model: Module, grad_outputs: Optional[Tensor]) -> Tensor:
    """Flat Jacobian row over net.parameters() (classification_ll_diagonal.py:127-130)."""
    g = [torch.flatten(torch.autograd.grad(out, [p], grad_outputs=grad_outputs, retain_graph=True,
                                           allow_unused=True)[0]) for p in model.parameters()]
    return torch.cat(g, dim=0).unsqueeze(0)
