"""Development timing of the batched Jacobi eigensolver."""
import sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200.utilities import eigh_factors
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
def psd(d, n):
    x = torch.relu(torch.randn(n, d, generator=g))
    return (x.t() @ x / n).to(dev)
for dims in ([26, 126, 161, 81, 5, 10, 80, 10], [785, 1025, 1025, 1024, 1024, 10], [2049], [4097]):
    mats = [psd(d, 512) for d in dims]
    eigh_factors(mats)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    vals, vecs = eigh_factors(mats)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3
    ref = torch.linalg.eigvalsh(0.5 * (mats[1 if len(mats) > 1 else 0] + mats[1 if len(mats) > 1 else 0].t()).double())
    err = (vals[1 if len(mats) > 1 else 0].double() - ref).abs().max().item() / ref.abs().max().item()
    t0 = time.perf_counter()
    for m in mats: torch.linalg.eigh(m)
    torch.cuda.synchronize()
    tt = (time.perf_counter() - t0) * 1e3
    print(f"dims={dims}: bk_eigh_batched {ms:.1f} ms (eigenvalue err {err:.1e}); torch.linalg.eigh (cuSOLVER, one by one) {tt:.1f} ms", flush=True)
