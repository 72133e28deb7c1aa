"""ncu / timing target: MC predictive of the cfg1 MLP (784-1024-1024-10, batch 256, S = 30), per precision and path."""
import sys
import time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.predictive import mc_logits
from bnn_kfac_b200.wrapper import MLP
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = MLP([784, 1024, 1024, 10]).to(dev)
model.weight_init_uniform(0.05)
x = torch.rand(256, 1, 28, 28, device=dev)
y = torch.randint(0, 10, (256,), device=dev)
for prec in (sys.argv[1:] or ["bf16x3", "bf16"]):
    est = KFAC(model, precision=prec)
    loss = torch.nn.functional.cross_entropy(model(x), y)
    model.zero_grad(); loss.backward()
    est.update(256)
    est.invert(1e2, 1e4)
    for imp in (None, True, False):
        for _ in range(3):
            mc_logits(est, x, 30, implicit=imp)
        torch.cuda.synchronize()
        c0 = L.bk_launch_count()
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(10):
            mc_logits(est, x, 30, implicit=imp)
        e1.record(); torch.cuda.synchronize()
        print(f"{prec} implicit={imp}: {e0.elapsed_time(e1) / 10:.3f} ms gpu, {(time.perf_counter() - t0) * 100:.3f} ms wall, "
              f"{(L.bk_launch_count() - c0) / 10:.0f} launches", flush=True)
    for h in est.hooks:
        h.remove()
