"""Development timing: one damped inversion of a 4097 x 4097 factor (and a batch of 8)."""
import sys, ctypes as C
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import invert_factors
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(8)
d = int(sys.argv[1]) if len(sys.argv) > 1 else 4097
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 1
x = torch.relu(torch.randn(4096, d - 1, generator=g)).to(dev)
xa = torch.cat([x, torch.ones(4096, 1, device=dev)], 1)
F_ = (xa.T @ xa / 4096).contiguous()
fs = [F_.clone() for _ in range(nb)]
invert_factors(fs, [1.0] * nb, [200.0] * nb)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
invert_factors(fs, [1.0] * nb, [200.0] * nb)
e1.record(); torch.cuda.synchronize()
print(f"chol_inv d={d} batch={nb}: {e0.elapsed_time(e1):.2f} ms")
