"""Bring-up check of the kernel library on a real B200 (development tool, not a test).

Compares every kernel family with plain torch on the GPU and prints one line per case; exits non-zero
if any case is out of tolerance.  Run under `timeout` through gpurun.
"""
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib  # noqa: E402

import os  # noqa: E402

L = _lib.load(strict=False)
_lib.require_device()
if os.environ.get("BK_CTA_GROUP"):
    L.bk_set_cta_group(int(os.environ["BK_CTA_GROUP"]))
dev = torch.device("cuda:0")
FAIL = []


def relerr(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item()


def report(name, err, tol):
    ok = err <= tol
    print(f"{'ok  ' if ok else 'FAIL'} {name:58s} err={err:.3e} tol={tol:.1e}", flush=True)
    if not ok:
        FAIL.append(name)


def split(x):
    hi = x.to(torch.bfloat16)
    lo = (x - hi.float()).to(torch.bfloat16)
    return hi.contiguous(), lo.contiguous()


def pad8(x):
    """Pad the last dim to a multiple of 8 and return (storage, ld)."""
    k = x.shape[-1]
    ld = (k + 7) // 8 * 8
    if ld == k:
        return x.contiguous(), ld
    out = torch.zeros(*x.shape[:-1], ld, dtype=x.dtype, device=x.device)
    out[..., :k] = x
    return out, ld


def gemm(A, B, prec, flags=0, alpha=1.0, beta=0.0, C=None, bias=None, want_bf16=False):
    """A [b?, M, K], B [b?, N, K] fp32 -> fp32 C via bk_gemm_nt."""
    batched = A.dim() == 3 or B.dim() == 3
    batch = (A.shape[0] if A.dim() == 3 else B.shape[0]) if batched else 1
    M, K = A.shape[-2:]
    N = B.shape[-2]
    Ah, Al = split(A)
    Bh, Bl = split(B)
    Ah, lda = pad8(Ah); Al, _ = pad8(Al); Bh, ldb = pad8(Bh); Bl, _ = pad8(Bl)
    sa = M * lda if A.dim() == 3 else 0
    sb = N * ldb if B.dim() == 3 else 0
    out = C if C is not None else torch.zeros(batch, M, N, device=dev)
    Oh = torch.zeros(batch, M, N, dtype=torch.bfloat16, device=dev) if want_bf16 else None
    Ol = torch.zeros_like(Oh) if want_bf16 else None
    rc = L.bk_gemm_nt(Ah.data_ptr(), Al.data_ptr() if prec == 3 else 0, lda, sa,
                      Bh.data_ptr(), Bl.data_ptr() if prec == 3 else 0, ldb, sb,
                      M, N, K, batch, prec, flags, alpha, beta,
                      out.data_ptr(), N, M * N,
                      _lib.ptr(bias), (bias.shape[-1] if bias is not None and bias.dim() == 2 else 0),
                      _lib.ptr(Oh), _lib.ptr(Ol), N, M * N, _lib.stream_ptr())
    _lib.check(rc, "bk_gemm_nt")
    torch.cuda.synchronize()
    return out, Oh, Ol


def case(fn):
    try:
        fn()
    except Exception:  # keep going: one broken family must not hide the others
        traceback.print_exc()
        FAIL.append(fn.__name__)
        try:
            torch.cuda.synchronize()
        except Exception:
            print("CUDA context is dead; stopping", flush=True)
            summary()
            sys.exit(2)


def summary():
    print("FAILED:" if FAIL else "ALL OK", FAIL, flush=True)


# ----------------------------------------------------------------------------------------------
def t_gemm_basic():
    g = torch.Generator(device="cpu").manual_seed(1)
    for (M, N, K) in [(128, 256, 64), (128, 256, 256), (256, 256, 64), (256, 256, 512), (200, 300, 100),
                      (1000, 520, 777), (64, 8, 40), (513, 1025, 130)]:
        A = torch.randn(M, K, generator=g).to(dev)
        B = torch.randn(N, K, generator=g).to(dev)
        ref_bf = (A.to(torch.bfloat16).double() @ B.to(torch.bfloat16).double().T)
        ref = A.double() @ B.double().T
        out, _, _ = gemm(A, B, 1)
        report(f"gemm bf16 {M}x{N}x{K} (vs bf16-rounded inputs)", relerr(out[0], ref_bf), 2e-5)
        out, _, _ = gemm(A, B, 3)
        report(f"gemm bf16x3 {M}x{N}x{K} (vs fp64)", relerr(out[0], ref), 3e-5)


def t_gemm_epilogue():
    g = torch.Generator(device="cpu").manual_seed(2)
    S, M, N, K = 3, 150, 270, 200
    A = torch.randn(S, M, K, generator=g).to(dev)
    B = torch.randn(S, N, K, generator=g).to(dev)
    bias = torch.randn(S, N, generator=g).to(dev)
    C0 = torch.randn(S, M, N, generator=g).to(dev)
    ref = torch.relu(0.5 * torch.einsum("smk,snk->smn", A.double(), B.double()) + 2.0 * C0.double()
                     + bias.double()[:, None, :])
    out, Oh, Ol = gemm(A, B, 3, flags=_lib.GEMM_RELU, alpha=0.5, beta=2.0, C=C0.clone(), bias=bias,
                       want_bf16=True)
    report("gemm batched alpha/beta/bias/relu fp32 out", relerr(out, ref), 3e-5)
    report("gemm batched split-bf16 out (hi+lo)", relerr(Oh.float() + Ol.float(), ref), 3e-5)
    # shared B operand across the batch
    out, _, _ = gemm(A, B[0], 3)
    report("gemm batched, B shared (stride 0)", relerr(out, torch.einsum("smk,nk->smn", A.double(), B[0].double())), 3e-5)


def t_gemm_tri():
    g = torch.Generator(device="cpu").manual_seed(3)
    M, N = 700, 600
    Lm = torch.tril(torch.randn(M, M, generator=g)).to(dev)
    Z = torch.randn(N, M, generator=g).to(dev)
    out, _, _ = gemm(Lm, Z, 3, flags=_lib.GEMM_TRI_A)
    report("gemm TRI_A (lower-triangular A)", relerr(out[0], Lm.double() @ Z.double().T), 3e-5)
    Lb = torch.tril(torch.randn(N, N, generator=g)).to(dev)
    X = torch.randn(M, N, generator=g).to(dev)
    out, _, _ = gemm(X, Lb, 3, flags=_lib.GEMM_TRI_B)
    report("gemm TRI_B (lower-triangular B)", relerr(out[0], X.double() @ Lb.double().T), 3e-5)


def syrk_call(x, has_bias, in_scale, alpha, beta, prec, state=None):
    n, d = x.shape
    dp = d + has_bias
    st = state if state is not None else torch.zeros(dp, dp, device=dev)
    wsb = L.bk_syrk_workspace_bytes(n, d, has_bias, prec)
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    rc = L.bk_syrk_accum(st.data_ptr(), dp, x.data_ptr(), x.stride(0), n, d, has_bias, in_scale,
                         alpha, beta, prec, ws.data_ptr(), wsb, _lib.stream_ptr())
    _lib.check(rc, "bk_syrk_accum")
    torch.cuda.synchronize()
    return st


def syrk_ref(x, has_bias, in_scale, alpha):
    xd = x.double() * in_scale
    if has_bias:
        xd = torch.cat([xd, torch.ones(x.shape[0], 1, dtype=torch.float64, device=x.device)], 1)
    return alpha * (xd.T @ xd)


def t_syrk():
    g = torch.Generator(device="cpu").manual_seed(4)
    for (n, d, hb) in [(256, 784, 1), (256, 1024, 0), (200, 300, 1), (250, 301, 1), (100, 513, 1), (77, 258, 0), (30, 30, 1), (30, 1, 1), (64, 159, 1),
                       (513, 161, 1), (1000, 80, 0)]:
        x = torch.relu(torch.randn(n, d, generator=g)).to(dev)
        ref = syrk_ref(x, hb, 1.0, 1.0 / n)
        for prec, tol in [(1, 2e-3), (3, 3e-5)]:
            st = syrk_call(x, hb, 1.0, 1.0 / n, 0.0, prec)
            report(f"syrk n={n} d={d} bias={hb} prec={prec}", relerr(st, ref), tol)
        # accumulate twice (state += ...), scaled input
        st = syrk_call(x, hb, 1.0, 1.0 / n, 0.0, 3)
        st = syrk_call(x, hb, 2.0, 1.0 / n, 1.0, 3, state=st)
        ref2 = ref + syrk_ref(x, hb, 2.0, 1.0 / n)
        report(f"syrk n={n} d={d} bias={hb} accumulate", relerr(st, ref2), 3e-5)
        report(f"syrk n={n} d={d} symmetric", (st - st.T).abs().max().item() / st.abs().max().item(), 1e-6)


def t_syrk_edges():
    """Boundary / ragged shapes: SIMT <-> tensor-core switch at d' = 176 | 177, rows that are not 16 B
    aligned (ones-row fallback), batch sizes that are not multiples of 8, a single sample, pitched
    state, grouped launch with mixed shapes."""
    import ctypes as C
    g = torch.Generator(device="cpu").manual_seed(14)
    for (n, d, hb) in [(37, 175, 1), (37, 176, 1), (37, 176, 0), (37, 177, 0), (1, 300, 1), (7, 193, 1),
                       (130, 255, 1), (130, 257, 0), (129, 512, 1), (1000, 201, 1)]:
        x = torch.randn(n, d, generator=g).to(dev)
        ref = syrk_ref(x, hb, 1.5, 1.0 / n)
        for prec, tol in [(0, 3e-6), (3, 3e-5)]:
            st = syrk_call(x, hb, 1.5, 1.0 / n, 0.0, prec)
            report(f"syrk edge n={n} d={d} bias={hb} prec={prec}", relerr(st, ref), tol)
    # pitched state view (ld % 4 == 0 -> TMA-reduce epilogue) must equal the dense one
    n, d = 96, 300
    x = torch.randn(n, d, generator=g).to(dev)
    buf = torch.full((d + 1, 304), float("nan"), device=dev)
    st = buf[:, :d + 1]
    wsb = L.bk_syrk_workspace_bytes(n, d, 1, 3)
    ws = torch.empty(max(wsb, 256), dtype=torch.uint8, device=dev)
    for beta in (0.0, 1.0):
        _lib.check(L.bk_syrk_accum(st.data_ptr(), 304, x.data_ptr(), d, n, d, 1, 1.0, 1.0 / n, beta, 3,
                                   ws.data_ptr(), wsb, _lib.stream_ptr()), "syrk pitched")
    torch.cuda.synchronize()
    report("syrk pitched state, beta 0 then 1 (TMA-reduce epilogue)", relerr(st, 2 * syrk_ref(x, 1, 1.0, 1.0 / n)), 3e-5)
    report("syrk pitched state leaves the padding untouched", float(torch.isnan(buf[:, d + 1:]).all().item() == 0), 0.0)
    # grouped call with mixed shapes (wide + small + unaligned)
    shapes = [(64, 400, 1), (64, 256, 0), (64, 20, 1), (50, 301, 1), (64, 1024, 1)]
    xs = [torch.randn(n_, d_, generator=g).to(dev) for n_, d_, _ in shapes]
    cnt = len(shapes)
    ns = (C.c_int * cnt)(*[s_[0] for s_ in shapes]); ds = (C.c_int * cnt)(*[s_[1] for s_ in shapes])
    hbs = (C.c_int * cnt)(*[s_[2] for s_ in shapes])
    nb = L.bk_syrk_grouped_workspace_bytes(ns, ds, hbs, cnt, 3)
    ws = torch.empty(max(nb, 256), dtype=torch.uint8, device=dev)
    # flags: 0 = mirrored epilogue, staging overlapped; 1 = lower-only (+ bk_sym_finalize); 3 = lower-only, no overlap
    for flags in (0, 1, 3, 2):
        sts = []
        for (n_, d_, hb_) in shapes:
            dp = d_ + hb_
            pitch = (dp + 3) // 4 * 4 if dp > 176 else dp
            sts.append(torch.zeros(dp, pitch, device=dev)[:, :dp])
        for beta in (0.0, 1.0):
            rc = L.bk_syrk_accum_grouped((C.c_void_p * cnt)(*[t.data_ptr() for t in sts]),
                                         (C.c_longlong * cnt)(*[t.stride(0) for t in sts]),
                                         (C.c_void_p * cnt)(*[t.data_ptr() for t in xs]), None,
                                         (C.c_longlong * cnt)(*[t.stride(0) for t in xs]), ns, ds, hbs,
                                         (C.c_float * cnt)(*[1.0] * cnt), (C.c_float * cnt)(*[1.0 / s_[0] for s_ in shapes]),
                                         (C.c_float * cnt)(*[beta] * cnt), cnt, 3, flags, ws.data_ptr(), nb,
                                         _lib.stream_ptr())
            _lib.check(rc, "bk_syrk_accum_grouped")
        scale = 1.0
        if flags & 1:
            scale = 0.5
            _lib.check(L.bk_sym_finalize((C.c_void_p * cnt)(*[t.data_ptr() for t in sts]),
                                         (C.c_longlong * cnt)(*[t.stride(0) for t in sts]),
                                         (C.c_int * cnt)(*[t.shape[0] for t in sts]), cnt, scale, _lib.stream_ptr()),
                       "bk_sym_finalize")
        torch.cuda.synchronize()
        for (n_, d_, hb_), x_, st_ in zip(shapes, xs, sts):
            report(f"grouped syrk flags={flags} n={n_} d={d_} bias={hb_} (beta 0 then 1)",
                   relerr(st_, 2 * scale * syrk_ref(x_, hb_, 1.0, 1.0 / n_)), 3e-5)
            report(f"grouped syrk flags={flags} d={d_} symmetric", (st_ - st_.t()).abs().max().item(), 0.0)


def t_chol_small():
    g = torch.Generator(device="cpu").manual_seed(8)
    fs = []
    for d in (1, 5, 10, 26, 64, 65, 81, 126, 161, 300, 785, 1025):
        x = torch.relu(torch.randn(256, d, generator=g)).to(dev)
        F_ = (x.T @ x / 256)
        F_ = F_ + 0.01 * torch.randn(d, d, generator=g).to(dev) * 1e-3  # slightly non-symmetric input
        fs.append(F_.contiguous())
    for (add, mult) in [(0.04, 200.0), (1.0, 200.0)]:
        rc, outs = chol_inv(fs, [add] * len(fs), [mult] * len(fs))
        report(f"chol_inv batched rc==0 add={add}", float(rc), 0.0)
        for F_, Lo in zip(fs, outs):
            Lref, R = chol_ref(F_, add, mult)
            d = F_.shape[0]
            report(f"chol_inv d={d} add={add} L vs reference", relerr(Lo, Lref), 1e-3)
            report(f"chol_inv d={d} upper triangle zero", torch.triu(Lo, 1).abs().max().item() if d > 1 else 0.0, 0.0)
    bad = torch.eye(70, device=dev); bad[40, 40] = -5.0
    rc, _ = chol_inv([fs[3], bad], [0.0, 0.0], [1.0, 1.0])
    report("chol_inv reports non-SPD factor index (expect 2)", abs(rc - 2), 0.0)


def t_syrk_wide():
    g = torch.Generator(device="cpu").manual_seed(5)
    n, d = 4096, 4096
    x = torch.randn(n, d, generator=g).to(dev)
    ref = syrk_ref(x, 1, 1.0, 1.0 / n)
    st = syrk_call(x, 1, 1.0, 1.0 / n, 0.0, 1)
    report("syrk 4096x4096 bias bf16, N(0,1) inputs not bf16-representable", relerr(st, ref), 2e-3)
    report("syrk 4096 A[-1,-1]==1", abs(st[-1, -1].item() - 1.0), 1e-6)
    st = syrk_call(x, 1, 1.0, 1.0 / n, 0.0, 3)
    report("syrk 4096x4096 bias bf16x3", relerr(st, ref), 3e-5)
    # timing: staged operand, SYRK only
    ldt = n
    hi = torch.empty(d + 1, ldt, dtype=torch.bfloat16, device=dev)
    lo = torch.empty_like(hi)
    L.bk_transpose_split(x.data_ptr(), d, n, d, 1.0, 1, hi.data_ptr(), lo.data_ptr(), ldt, _lib.stream_ptr())
    st = torch.zeros(d + 1, d + 1, device=dev)
    for dd in (d + 1, d):
        for cg in (1, 2):
            L.bk_set_cta_group(cg)
            for prec in (1, 3):
                for _ in range(3):
                    L.bk_syrk_accum_staged(st.data_ptr(), d + 1, hi.data_ptr(), lo.data_ptr(), ldt, n, dd, 1.0 / n, 1.0, prec, _lib.stream_ptr())
                e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
                e0.record()
                for _ in range(10):
                    L.bk_syrk_accum_staged(st.data_ptr(), d + 1, hi.data_ptr(), lo.data_ptr(), ldt, n, dd, 1.0 / n, 1.0, prec, _lib.stream_ptr())
                e1.record(); torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                flops = dd * (dd + 1) * n  # SYRK count: one FMA (2 flop) per lower-triangle entry per sample
                print(f"time syrk_staged d={dd} cta_group={cg} prec={prec}: {ms*1e3:.1f} us  {flops/ms/1e9:.1f} TFLOP/s (algorithmic)", flush=True)
    L.bk_set_cta_group(int(os.environ.get("BK_CTA_GROUP", "0")))
    ws = torch.empty(L.bk_syrk_workspace_bytes(n, d, 1, 1), dtype=torch.uint8, device=dev)
    for _ in range(3):
        L.bk_syrk_accum(st.data_ptr(), d + 1, x.data_ptr(), d, n, d, 1, 1.0, 1.0 / n, 1.0, 1, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        L.bk_syrk_accum(st.data_ptr(), d + 1, x.data_ptr(), d, n, d, 1, 1.0, 1.0 / n, 1.0, 1, ws.data_ptr(), ws.numel(), _lib.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    print(f"time bk_syrk_accum (stage + syrk + bias border) d=4096+1 bf16: {e0.elapsed_time(e1)*100:.1f} us", flush=True)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        L.bk_transpose_split(x.data_ptr(), d, n, d, 1.0, 1, hi.data_ptr(), 0, ldt, _lib.stream_ptr())
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"time transpose_split 4096x4096 (hi only): {ms*1e3:.1f} us  {(n*d*6)/ms/1e6:.0f} GB/s", flush=True)
    # plain torch bf16 matmul for context
    xb = x.to(torch.bfloat16)
    for _ in range(3):
        xb.T @ xb
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(10):
        xb.T @ xb
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"time torch bf16 x.T@x 4096^3 (full GEMM): {ms*1e3:.1f} us  {2*n*d*d/ms/1e9:.1f} TFLOP/s", flush=True)


def t_conv():
    g = torch.Generator(device="cpu").manual_seed(6)
    for (n, c, h, w, k, pad, stride, o) in [(16, 1, 28, 28, 5, 0, 1, 5), (16, 5, 12, 12, 5, 0, 1, 10),
                                            (8, 6, 14, 14, 5, 0, 1, 16), (4, 3, 13, 13, 3, 1, 2, 6)]:
        x = torch.rand(n, c, h, w, generator=g).to(dev)
        u = F.unfold(x, k, padding=pad, stride=stride)  # [n, c*k*k, L]
        u = u.permute(1, 0, 2).reshape(u.shape[1], -1).double()
        u = torch.cat([u, torch.ones_like(u[:1])], 0)
        ref = u @ u.T / u.shape[1]
        dp = c * k * k + 1
        st = torch.zeros(dp, dp, device=dev)
        rc = L.bk_conv_a_accum(st.data_ptr(), dp, x.data_ptr(), n, c, h, w, k, k, pad, pad, stride, stride,
                               1, 1.0 / u.shape[1], 0.0, _lib.stream_ptr())
        _lib.check(rc, "bk_conv_a_accum"); torch.cuda.synchronize()
        report(f"conv A n={n} c={c} {h}x{w} k={k} pad={pad} s={stride}", relerr(st, ref), 2e-5)
        oh = (h + 2 * pad - k) // stride + 1
        gr = torch.randn(n, o, oh, oh, generator=g).to(dev)
        gg = (gr.double() * n).permute(1, 0, 2, 3).reshape(o, -1)
        refg = gg @ gg.T / gg.shape[1]
        st = torch.zeros(o, o, device=dev)
        rc = L.bk_conv_g_accum(st.data_ptr(), o, gr.data_ptr(), n, o, oh * oh, float(n), 1.0 / gg.shape[1], 0.0,
                               _lib.stream_ptr())
        _lib.check(rc, "bk_conv_g_accum"); torch.cuda.synchronize()
        report(f"conv G n={n} o={o} hw={oh*oh}", relerr(st, refg), 2e-5)


def t_diag():
    g = torch.Generator(device="cpu").manual_seed(7)
    do, di = 37, 53
    wg = torch.randn(do, di, generator=g).to(dev)
    bg = torch.randn(do, generator=g).to(dev)
    st = torch.rand(do, di + 1, generator=g).to(dev)
    ref = st.double() + torch.cat([wg, bg[:, None]], 1).double() ** 2 * 32
    _lib.check(L.bk_diag_accum(st.data_ptr(), wg.data_ptr(), bg.data_ptr(), do, di, 32.0, 1.0, _lib.stream_ptr()), "diag_accum")
    report("diag accum", relerr(st, ref), 1e-6)
    inv = torch.empty_like(st)
    _lib.check(L.bk_diag_invert(inv.data_ptr(), st.data_ptr(), st.numel(), 0.04, 200.0, _lib.stream_ptr()), "diag_invert")
    report("diag invert", relerr(inv, torch.reciprocal(200.0 * st.double() + 0.04).sqrt()), 1e-6)
    z = torch.randn(3, do, di + 1, generator=g).to(dev)
    out = torch.empty_like(z)
    _lib.check(L.bk_diag_sample(out.data_ptr(), inv.data_ptr(), inv.numel(), 3, 0, 0, 0, z.data_ptr(), _lib.stream_ptr()), "diag_sample")
    report("diag sample (external z)", relerr(out, z * inv), 1e-6)
    big = torch.ones(1 << 20, device=dev)
    outb = torch.empty(2, 1 << 20, device=dev)
    _lib.check(L.bk_diag_sample(outb.data_ptr(), big.data_ptr(), big.numel(), 2, 1234, 0, 7, 0, _lib.stream_ptr()), "diag_sample")
    torch.cuda.synchronize()
    report("philox normal mean", abs(outb.mean().item()), 5e-3)
    report("philox normal var", abs(outb.var().item() - 1.0), 5e-3)
    report("philox samples differ", float((outb[0] == outb[1]).float().mean().item()), 1e-3)
    J = torch.randn(5, st.numel(), generator=g).to(dev)
    q = torch.empty(5, device=dev)
    _lib.check(L.bk_diag_quadform(q.data_ptr(), J.data_ptr(), J.stride(0), inv.data_ptr(), inv.numel(), 5, _lib.stream_ptr()), "diag_quadform")
    report("diag quadform", relerr(q, (J.double() ** 2 * inv.double().flatten()).sum(1)), 1e-6)


def t_philox():
    zf = torch.empty(2, 100, 64, device=dev)
    zh = torch.empty(2, 100, 64, dtype=torch.bfloat16, device=dev)
    zl = torch.empty_like(zh)
    _lib.check(L.bk_philox_normal(42, 5, 1, 100, 64, 2, zf.data_ptr(), 64, 6400, zh.data_ptr(), zl.data_ptr(), 64, 6400, _lib.stream_ptr()), "philox")
    torch.cuda.synchronize()
    report("philox hi+lo == fp32", relerr(zh.float() + zl.float(), zf), 1e-5)
    z2 = torch.empty(1, 100, 64, device=dev)
    _lib.check(L.bk_philox_normal(42, 6, 1, 100, 64, 1, z2.data_ptr(), 64, 6400, 0, 0, 0, 0, _lib.stream_ptr()), "philox")
    torch.cuda.synchronize()
    report("philox sample id addressing (sample0+s)", (z2[0] - zf[1]).abs().max().item(), 0.0)


def chol_inv(factors, add, mult):
    import ctypes as C
    n = len(factors)
    outs = [torch.empty_like(f) for f in factors]
    dims = (C.c_int * n)(*[f.shape[0] for f in factors])
    fp = (C.c_void_p * n)(*[f.data_ptr() for f in factors])
    op = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    av = (C.c_float * n)(*add)
    mv = (C.c_float * n)(*mult)
    wsb = L.bk_chol_inv_workspace_bytes(dims, n)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    rc = L.bk_damp_chol_inv_batched(fp, op, dims, av, mv, n, ws.data_ptr(), wsb, _lib.stream_ptr())
    return rc, outs


def chol_ref(F, add, mult):
    Fd = F.double()
    R = mult ** 0.5 * Fd + add ** 0.5 * torch.eye(F.shape[0], dtype=torch.float64, device=F.device)
    R = (R + R.T) / 2
    return torch.linalg.cholesky(torch.linalg.inv(R)), R


def t_chol():
    g = torch.Generator(device="cpu").manual_seed(8)
    fs, refs = [], []
    for d in (1, 5, 10, 26, 64, 81, 126, 161, 300, 785, 1025):
        x = torch.relu(torch.randn(256, d, generator=g)).to(dev)
        F_ = (x.T @ x / 256)
        F_ = F_ + 0.01 * torch.randn(d, d, generator=g).to(dev) * 1e-3  # slightly non-symmetric input
        fs.append(F_.contiguous())
    for (add, mult) in [(0.04, 200.0), (1.0, 200.0)]:
        rc, outs = chol_inv(fs, [add] * len(fs), [mult] * len(fs))
        report(f"chol_inv batched rc==0 add={add}", float(rc), 0.0)
        for F_, Lo in zip(fs, outs):
            Lref, R = chol_ref(F_, add, mult)
            d = F_.shape[0]
            report(f"chol_inv d={d} add={add} L vs reference", relerr(Lo, Lref), 1e-3)
            report(f"chol_inv d={d} add={add} L L^T R == I", relerr(Lo.double() @ Lo.double().T @ R, torch.eye(d, dtype=torch.float64, device=dev)), 1e-3)
            report(f"chol_inv d={d} upper triangle zero", torch.triu(Lo, 1).abs().max().item() if d > 1 else 0.0, 0.0)
    bad = torch.eye(70, device=dev); bad[40, 40] = -5.0
    rc, _ = chol_inv([fs[3], bad], [0.0, 0.0], [1.0, 1.0])
    report("chol_inv reports non-SPD factor index (expect 2)", abs(rc - 2), 0.0)
    # wide: 4097 (timing)
    x = torch.relu(torch.randn(4096, 4096, generator=g)).to(dev)
    xa = torch.cat([x, torch.ones(4096, 1, device=dev)], 1)
    F_ = (xa.T @ xa / 4096).contiguous()
    rc, outs = chol_inv([F_], [1.0], [200.0])
    Lref, R = chol_ref(F_, 1.0, 200.0)
    report("chol_inv d=4097 (1,200) L vs reference", relerr(outs[0], Lref), 1e-3)
    report("chol_inv d=4097 inverse L L^T vs inv(R)", relerr(outs[0].double() @ outs[0].double().T, torch.linalg.inv(R)), 1e-3)
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    rc, outs = chol_inv([F_], [1.0], [200.0])
    e1.record(); torch.cuda.synchronize()
    print(f"time chol_inv d=4097: {e0.elapsed_time(e1):.2f} ms", flush=True)
    F2 = [F_, F_.clone(), F_[:4096, :4096].contiguous()]
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    rc, outs = chol_inv(F2, [1.0] * 3, [200.0] * 3)
    e1.record(); torch.cuda.synchronize()
    print(f"time chol_inv 3 factors (4097,4097,4096) batched: {e0.elapsed_time(e1):.2f} ms", flush=True)
    t = time.time(); Lc = torch.linalg.cholesky(torch.linalg.inv(R.float())); torch.cuda.synchronize()
    t = time.time(); Lc = torch.linalg.cholesky(torch.linalg.inv(R.float())); torch.cuda.synchronize()
    print(f"time torch inverse+cholesky fp32 d=4097 on GPU (cuSOLVER): {(time.time()-t)*1e3:.2f} ms", flush=True)


if __name__ == "__main__":
    print(L.bk_version().decode(), torch.cuda.get_device_name(0), flush=True)
    t0 = time.time()
    only = sys.argv[1:]
    allfn = (t_diag, t_philox, t_conv, t_gemm_basic, t_gemm_epilogue, t_gemm_tri, t_syrk, t_syrk_edges, t_syrk_wide, t_chol_small, t_chol)
    for fn in [f for f in allfn if not only or f.__name__ in only]:
        print(f"--- {fn.__name__}", flush=True)
        case(fn)
    print(f"elapsed {time.time()-t0:.1f}s")
    summary()
    sys.exit(1 if FAIL else 0)
