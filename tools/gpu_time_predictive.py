"""Development timing of the posterior-predictive path on cfg5 shapes (MLP 4096x3 -> 10)."""
import sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.predictive import mc_moments
from bnn_kfac_b200.wrapper import MLP
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
W = [4096, 4096, 4096, 4096, 10]
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
torch.manual_seed(0)
model = MLP(W).to(dev)
est = KFAC(model, precision="bf16")
layers = [l for _, l in est._selected_layers()]
g = torch.Generator().manual_seed(1)
for l, (a, b) in zip(layers, zip(W[:-1], W[1:])):
    est.record[l] = [torch.randn(4096, a, generator=g).to(dev), (torch.randn(4096, b, generator=g) / 4096).to(dev)]
est.update(4096)
est.invert(1.0, 200.0)
x = torch.randn(B, W[0], device=dev)
for _ in range(2):
    mc_moments(est, x, S, sample0=0)
torch.cuda.synchronize()
c0 = L.bk_launch_count()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
mc_moments(est, x, S, sample0=0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"S={S} B={B}: {ms:.2f} ms -> {S/ms*1e3:.0f} weight samples/s, {S*B/ms*1e3:.0f} (samples x inputs)/s, launches={L.bk_launch_count()-c0}")
