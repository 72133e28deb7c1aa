"""ncu target: one damped inversion of a 4097-wide factor, kernel by kernel (graph replay off)."""
import sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import _Workspace, invert_factors
L = _lib.load(); _lib.require_device()
L.bk_set_chol_graph(0)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(8)
d = 4097
x = torch.relu(torch.randn(4096, d - 1, generator=g)).to(dev)
xa = torch.cat([x, torch.ones(4096, 1, device=dev)], 1)
F_ = (xa.T @ xa / 4096).contiguous()
ws = _Workspace()
invert_factors([F_], [1.0], [200.0], ws)
torch.cuda.synchronize()
print("ok")
