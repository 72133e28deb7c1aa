"""Development timing: sweep of bk_set_chol_far_sms (SMs left to the background outer updates of the inversion)."""
import sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import _Workspace, invert_factors
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(8)
d = 4097
x = torch.relu(torch.randn(4096, d - 1, generator=g)).to(dev)
xa = torch.cat([x, torch.ones(4096, 1, device=dev)], 1)
F_ = (xa.T @ xa / 4096).contiguous()
ws = _Workspace()
for dims in ([d], [d, d - 1] * 3 + [d, 10], [1025, 1024, 1025, 1024, 1025, 10], [2049, 2048, 2049, 10]):
    fs = [F_[:k, :k].contiguous() for k in dims]
    outs = None
    for cap, look in ([(64, 0), (64, 1), (0, 1)] if dims[0] >= 2048 else [(64, 0), (64, 1)]):
        L.bk_set_chol_far_sms(cap)
        L.bk_set_chol_lookahead(look)
        for _ in range(2):
            outs = invert_factors(fs, [1.0] * len(fs), [200.0] * len(fs), ws)
        ms = []
        for _ in range(7):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            outs = invert_factors(fs, [1.0] * len(fs), [200.0] * len(fs), ws)
            e1.record(); torch.cuda.synchronize()
            ms.append(round(e0.elapsed_time(e1), 2))
        print(f"dims={dims[0]}x{len(dims)} far_sms={cap} lookahead={look}: median {sorted(ms)[3]:.2f} ms  min {min(ms):.2f}  max {max(ms):.2f}", flush=True)
L.bk_set_chol_far_sms(64)
L.bk_set_chol_lookahead(1)
