"""Bring-up of bk_eigh_batched through ctypes: results after a fixed number of sweeps even if not converged."""
import ctypes as C, sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator().manual_seed(3)
def psd(d, n):
    x = torch.relu(torch.randn(n, d, generator=g))
    return (x.t() @ x / n).to(dev)
def solve(m, sweeps):
    d = m.shape[0]
    w = torch.empty(d, device=dev); v = torch.empty(d, d, device=dev)
    dims = (C.c_int * 1)(d); lds = (C.c_longlong * 1)(m.stride(0))
    fp = (C.c_void_p * 1)(m.data_ptr()); wp = (C.c_void_p * 1)(w.data_ptr()); vp = (C.c_void_p * 1)(v.data_ptr())
    nb = lib.bk_eigh_workspace_bytes(dims, 1)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    rc = lib.bk_eigh_batched(fp, lds, wp, vp, dims, 1, 0.5, sweeps, ws.data_ptr(), nb, 0)
    torch.cuda.synchronize()
    return rc, w, v, (time.perf_counter() - t0) * 1e3
import os
lib.bk_set_eigh_mode(int(os.environ.get('BK_EIGH_MODE', '0')))
lib.bk_set_eigh_pair_width(int(os.environ.get('BK_EIGH_PAIR', '0')))
d = int(sys.argv[1]); n = int(sys.argv[2])
m = psd(d, n)
S = (0.5 * (m + m.t())).double()
wref = torch.linalg.eigvalsh(S)
for sweeps in [int(a) for a in sys.argv[3:]]:
    rc, w, v, ms = solve(m, sweeps)
    vd = v.double()
    D = vd.t() @ S @ vd
    off = (D - torch.diag(torch.diag(D))).norm().item() / S.norm().item()
    err = (w.double() - wref).abs().max().item() / wref.abs().max().item()
    rec = ((vd * w.double()) @ vd.t() - S).norm().item() / S.norm().item()
    orth = (vd.t() @ vd - torch.eye(d, device=dev, dtype=torch.float64)).norm().item() / d ** 0.5
    print(f"d={d} n={n} sweeps={sweeps}: rc={rc} {ms:8.1f} ms  off {off:.2e}  eval err {err:.1e}  recon {rec:.1e}  orth {orth:.1e}", flush=True)
