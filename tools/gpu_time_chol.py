"""Development timing: damped inversion of 4097-wide factors (one, and the 8 factors of the wide MLP), with the
step sequence replayed from a CUDA graph (default) and enqueued kernel by kernel (bk_set_chol_graph(0)).  Single-shot
samples, no best-of: the spread is part of the result.  Also checks that both routes return identical bits."""
import sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import _Workspace, invert_factors
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(8)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
x = torch.relu(torch.randn(4096, d - 1, generator=g)).to(dev)
xa = torch.cat([x, torch.ones(4096, 1, device=dev)], 1)
F_ = (xa.T @ xa / 4096).contiguous()
ws = _Workspace()


def sample(fs, n=6):
    out = []
    for _ in range(n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        res = invert_factors(fs, [1.0] * len(fs), [200.0] * len(fs), ws)
        e1.record(); torch.cuda.synchronize()
        out.append(round(e0.elapsed_time(e1), 2))
    return out, res


for dims in ([d], [d, d - 1] * 3 + [d, 10]):
    fs = [F_[:k, :k].contiguous() for k in dims]
    ref = None
    for graph in (1, 0, 1):
        L.bk_set_chol_graph(graph)
        invert_factors(fs, [1.0] * len(fs), [200.0] * len(fs), ws)      # warm-up / capture
        ms, res = sample(fs)
        if ref is None:
            ref = res
        same = all(torch.equal(a, b) for a, b in zip(ref, res))
        print(f"chol_inv dims={dims[0]}x{len(dims)} graph={graph}: ms {ms}  median {sorted(ms)[len(ms)//2]:.2f}  "
              f"identical to first run: {same}", flush=True)
L.bk_set_chol_graph(1)
