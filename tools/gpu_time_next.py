"""Development timing of the SURVEY 8(f) rows f3 / f4 on the GPU box: calibration metrics (achieved GB/s of the
row kernel) and the INF pipeline on the cfg1 MLP (784-1024-1024-10, rank 100)."""
import sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import utilities as U
from bnn_kfac_b200.curvatures import EFB, INF, KFAC, Diagonal
from bnn_kfac_b200.wrapper import MLP
dev = torch.device("cuda:0")
torch.manual_seed(0)

def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

for rows, classes in ((1 << 22, 10), (1 << 17, 1000)):
    p = torch.softmax(torch.randn(rows, classes, device=dev), 1)
    lab = torch.randint(0, classes, (rows,), device=dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    U.calibration_rows(p, lab); torch.cuda.synchronize()
    e0.record()
    for _ in range(10): r = U.calibration_rows(p, lab)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    gb = (rows * classes * 4 + rows * 8 + rows * 20) / 1e9
    print(f"calibration_rows [{rows} x {classes}]: {ms:.3f} ms/call incl. the 4-double D2H = {gb / ms * 1e3:.0f} GB/s; "
          f"ECE(10 bins) {timed(lambda: U.expected_calibration_error(p, lab)):.2f} ms, "
          f"calibration_curve(20) {timed(lambda: U.calibration_curve(p, lab)):.2f} ms", flush=True)

model = MLP([784, 1024, 1024, 10]).to(dev)
kf, dg = KFAC(model), Diagonal(model)
def step(est_list):
    x = torch.rand(256, 784, device=dev)
    out = model(x)
    y = torch.distributions.Categorical(logits=out).sample()
    loss = torch.nn.functional.cross_entropy(out, y); model.zero_grad(); loss.backward()
    for e in est_list: e.update(256)
for _ in range(4): step([kf, dg])
t0 = time.perf_counter(); efb = EFB(model, kf.state); torch.cuda.synchronize()
print(f"EFB.__init__ (eigenvectors of 6 factors, d <= 1025): {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
for _ in range(4): step([efb])
t0 = time.perf_counter(); inf = INF(model, dg.state, kf.state, efb.state); torch.cuda.synchronize()
print(f"INF.__init__: {(time.perf_counter() - t0) * 1e3:.1f} ms", flush=True)
for rank in (100, 400):
    inf.state, inf.inv_state = dict(), dict()
    print(f"INF rank={rank}: update {timed(lambda: inf.update(rank=rank), 3):.2f} ms", end="", flush=True)
    t0 = time.perf_counter(); inf.invert(0.04, 200.0); torch.cuda.synchronize()
    layers = list(inf.inv_state.keys())
    rs = [inf.inv_state[l][3].shape[0] for l in layers]
    print(f", invert {(time.perf_counter() - t0) * 1e3:.1f} ms (r = {rs}), sample_and_replace "
          f"{timed(inf.sample_and_replace, 5):.2f} ms", flush=True)
