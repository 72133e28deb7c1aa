"""Bring-up probe: posterior-predictive timing in the bench's own sequence (update -> fwd/bwd step -> invert -> mc_moments)."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.predictive import mc_moments
from bnn_kfac_b200.wrapper import MLP
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
W, BATCH = bench.WIDTHS, bench.BATCH
torch.manual_seed(0)
model = MLP(W).to(dev)
est = KFAC(model, precision="bf16")
layers = [l for _, l in est._selected_layers()]
synth = bench.synth_batch(torch.Generator().manual_seed(1234), BATCH, W)
res = [(a.to(dev), g.to(dev)) for a, g in synth]


def ev(fn, reps=4):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def upd():
    for layer, (a, g) in zip(layers, res):
        est.record[layer] = [a, g]
    est.update(BATCH)


x = torch.randn(1024, W[0], device=dev)
for tag in sys.argv[1:] or ["plain"]:
    for _ in range(5):
        upd()
    if tag == "fullstep":
        xb = torch.randn(BATCH, W[0], device=dev); yb = torch.randint(0, 10, (BATCH,), device=dev)
        for _ in range(3):
            loss = torch.nn.functional.cross_entropy(model(xb), yb); model.zero_grad(); loss.backward(); est.update(BATCH)
    if tag == "nograph":
        L.bk_set_chol_graph(0)
    est.invert(1.0, 200.0)
    t_inv = ev(lambda: est.invert(1.0, 200.0), reps=1)
    for _ in range(2):
        mc_moments(est, x, 16, sample0=0)
    c0 = L.bk_launch_count()
    t = min(ev(lambda: mc_moments(est, x, 16, sample0=0)) for _ in range(3))
    print(f"{tag}: invert {t_inv:.2f} ms, mc_moments {t:.2f} ms, launches/call {(L.bk_launch_count() - c0) / 12:.1f}", flush=True)
