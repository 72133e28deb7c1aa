"""A/B of the factor-update step (cfg5, device-resident fp32 containers, precision bf16): overlapped staging with the
conventional tile grid (default), with the persistent staging kernel (BK_SYRK_STAGE_PERSISTENT: measured slower), and without overlap."""
import sys
import torch
sys.path.insert(0, ".")
import bench
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.wrapper import MLP
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
W, BATCH = bench.WIDTHS, bench.BATCH
model = MLP(W).to(dev)
est = KFAC(model, precision="bf16")
layers = [l for _, l in est._selected_layers()]
synth = bench.synth_batch(torch.Generator().manual_seed(1234), BATCH, W)
res = [(a.to(dev), g.to(dev)) for a, g in synth]


def step():
    for layer, (a, g) in zip(layers, res):
        est.record[layer] = [a, g]
    est.update(BATCH)


def ev(reps=20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for name, extra in (("tile-grid staging under the SYRK", 0), ("persistent staging under the SYRK", _lib.SYRK_STAGE_PERSISTENT),
                    ("no overlap", _lib.SYRK_NO_OVERLAP), ("tile-grid staging under the SYRK", 0)):
    est._syrk_flags_extra = extra
    for _ in range(5):
        step()
    ms = sorted(ev() for _ in range(5))
    print(f"{name:40s}: {ms[0]:.4f} ms/step (min of 5 x 20), median {ms[2]:.4f}", flush=True)
