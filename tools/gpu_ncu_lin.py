"""ncu / timing target: linearised KFAC predictive of a batch (classification_ll_block.py:114-141) on BaseNet_15k."""
import sys
import time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.predictive import linearised_kfac_classification
from bnn_kfac_b200.wrapper import BaseNet_15k, LeNet5
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = (LeNet5() if (sys.argv[1:] or ["basenet"])[0] == "lenet" else BaseNet_15k()).to(dev)
model.weight_init_uniform(0.05)
x = torch.rand(256, 1, 28, 28, device=dev)
y = torch.randint(0, 10, (256,), device=dev)
est = KFAC(model, precision="bf16x3")
loss = torch.nn.functional.cross_entropy(model(x), y)
model.zero_grad(); loss.backward()
est.update(256)
est.invert(1e2, 1e4)
for _ in range(3):
    linearised_kfac_classification(est, x)
torch.cuda.synchronize()
c0 = L.bk_launch_count()
t0 = time.perf_counter()
for _ in range(5):
    linearised_kfac_classification(est, x)
torch.cuda.synchronize()
print(f"linearised_kfac_classification batch 256: {(time.perf_counter() - t0) * 200:.3f} ms wall, {(L.bk_launch_count() - c0) / 5:.0f} library launches", flush=True)
