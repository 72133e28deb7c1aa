"""Development timing: implicit (weight-free) vs materialised MC forward across sizes."""
import sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.predictive import mc_logits
from bnn_kfac_b200.wrapper import MLP
dev = torch.device("cuda:0")
def ev(fn, reps=5):
    fn(); fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3
for widths, S in [([784, 1024, 1024, 10], 30), ([2048, 2048, 2048, 10], 16), ([512, 512, 512, 10], 30)]:
    torch.manual_seed(0)
    model = MLP(widths).to(dev); model.weight_init_uniform(0.05)
    est = KFAC(model, precision="bf16x3")
    x = torch.rand(256, widths[0], device=dev)
    loss = torch.nn.functional.cross_entropy(model(x), torch.randint(0, 10, (256,), device=dev))
    model.zero_grad(); loss.backward(); est.update(256); est.invert(1e2, 1e4)
    for B in (16, 64, 256, 1024):
        xt = torch.rand(B, widths[0], device=dev)
        ti = ev(lambda: mc_logits(est, xt, S, implicit=True))
        tm = ev(lambda: mc_logits(est, xt, S, implicit=False))
        print(f"widths={widths} S={S} B={B}: implicit {ti:.2f} ms, materialised {tm:.2f} ms", flush=True)
