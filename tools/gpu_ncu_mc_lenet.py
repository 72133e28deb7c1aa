"""ncu / timing target: MC predictive of LeNet-5 and BaseNet_15k (BASELINE config 4: batch 256, S = 100)."""
import sys
import time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.predictive import mc_predict
from bnn_kfac_b200.wrapper import BaseNet_15k, LeNet5
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
which = sys.argv[1] if len(sys.argv) > 1 else "lenet"
torch.manual_seed(0)
model = (LeNet5() if which == "lenet" else BaseNet_15k()).to(dev)
model.weight_init_uniform(0.05)
x = torch.rand(256, 1, 28, 28, device=dev)
y = torch.randint(0, 10, (256,), device=dev)
est = KFAC(model, precision="bf16x3")
loss = torch.nn.functional.cross_entropy(model(x), y)
model.zero_grad(); loss.backward()
est.update(256)
est.invert(1e2, 1e4)
for _ in range(3):
    mc_predict(est, x, 100)
torch.cuda.synchronize()
c0 = L.bk_launch_count()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(5):
    mc_predict(est, x, 100)
e1.record(); torch.cuda.synchronize()
print(f"{which}: mc_predict S=100 B=256: {e0.elapsed_time(e1) / 5:.3f} ms gpu, {(time.perf_counter() - t0) * 200:.3f} ms wall, "
      f"{(L.bk_launch_count() - c0) / 5:.0f} launches", flush=True)
