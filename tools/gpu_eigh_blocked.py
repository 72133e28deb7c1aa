"""Bring-up / timing of the tensor-core block-Jacobi eigensolver (bk_eigh_blocked.cu) against fp64
torch.linalg.eigh: eigenvalue error (of lambda_max), reconstruction, orthogonality, time; the element-wise
streamed path (bk_set_eigh_mode(1)) and cuSOLVER beside it."""
import sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.utilities import eigh_factors
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
lib = _lib.load()
def psd(d, n):
    x = torch.relu(torch.randn(n, d, generator=g))
    return (x.t() @ x / n).to(dev)
sizes = [int(a) for a in sys.argv[1:]] or [165, 300, 785, 1025, 2049, 4097]
for d in sizes:
    for n in (max(64, d // 4), 2 * d):          # rank-deficient and full-rank factors
        m = psd(d, n)
        S = (0.5 * (m + m.t())).double()
        wref = torch.linalg.eigvalsh(S)
        for mode in (0, 1):
            if mode == 1 and d > 1100:
                continue
            lib.bk_set_eigh_mode(mode)
            try:
                eigh_factors([m])
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                (w,), (v,) = eigh_factors([m])
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) * 1e3
                vd = v.double()
                err = (w.double() - wref).abs().max().item() / wref.abs().max().item()
                rec = ((vd * w.double()) @ vd.t() - S).norm().item() / S.norm().item()
                orth = (vd.t() @ vd - torch.eye(d, device=dev, dtype=torch.float64)).norm().item() / d ** 0.5
                print(f"d={d} n={n} mode={mode}: {ms:8.1f} ms  eval err {err:.1e}  recon {rec:.1e}  orth {orth:.1e}", flush=True)
            except Exception as e:
                print(f"d={d} n={n} mode={mode}: FAILED {type(e).__name__}: {e}", flush=True)
        lib.bk_set_eigh_mode(0)
        torch.linalg.eigh(m)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        torch.linalg.eigh(m)
        torch.cuda.synchronize()
        print(f"d={d} n={n} cuSOLVER syevd fp32: {(time.perf_counter() - t0) * 1e3:8.1f} ms", flush=True)
