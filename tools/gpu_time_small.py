"""Development timing of the small-shape configs (cfg1 MLP 784-1024-1024-10, cfg4 BaseNet_15k / LeNet-5)."""
import sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import KFAC, Diagonal
from bnn_kfac_b200.predictive import mc_predict, linearised_kfac_classification
from bnn_kfac_b200.wrapper import MLP, BaseNet_15k, LeNet5
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")

def ev(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, (time.perf_counter() - t0) / reps * 1e3

for name, ctor, shape in [("cfg1 MLP 784-1024-1024-10", lambda: MLP([784, 1024, 1024, 10]), (256, 1, 28, 28)),
                          ("cfg4 BaseNet_15k", BaseNet_15k, (256, 1, 28, 28)),
                          ("cfg4 LeNet-5", LeNet5, (256, 1, 28, 28))]:
    torch.manual_seed(0)
    model = ctor().to(dev)
    model.weight_init_uniform(0.05)
    est = KFAC(model, precision="bf16x3")
    x = torch.rand(*shape, device=dev)
    y = torch.randint(0, 10, (shape[0],), device=dev)
    def fwdbwd():
        loss = torch.nn.functional.cross_entropy(model(x), y)
        model.zero_grad(); loss.backward()
    fwdbwd()
    c0 = L.bk_launch_count()
    est.update(shape[0]); torch.cuda.synchronize()
    nl = L.bk_launch_count() - c0
    t_upd = ev(lambda: est.update(shape[0]))
    t_fb = ev(fwdbwd)
    t_inv = ev(lambda: est.invert(1e2, 1e4), reps=5, warm=2)
    S = 100 if "cfg4" in name else 30
    t_mc = ev(lambda: mc_predict(est, x, S), reps=10, warm=3)
    print(f"{name}: update {t_upd[0]*1e3:.0f} us gpu / {t_upd[1]*1e3:.0f} us wall ({nl} launches) -> {shape[0]/t_upd[1]*1e3:.0f} samples/s; "
          f"model fwd+bwd {t_fb[1]*1e3:.0f} us; invert {t_inv[1]:.2f} ms; mc_predict S={S} B={shape[0]}: {t_mc[1]:.2f} ms "
          f"-> {S*shape[0]/t_mc[1]*1e3:.0f} (samples x inputs)/s", flush=True)
    if "cfg4" in name:
        t_lin = ev(lambda: linearised_kfac_classification(est, x), reps=5, warm=1)
        print(f"   linearised KFAC predictive (batch 256): {t_lin[1]:.2f} ms", flush=True)

# cfg2: sampling-free linearised predictive of the 1-D regression net (regression_ll_block.py:84-140)
from bnn_kfac_b200.predictive import linearised_kfac_regression
class RegNet(torch.nn.Module):
    def __init__(self, n_hid=50):
        super().__init__()
        self.fc1 = torch.nn.Linear(1, n_hid); self.fc2 = torch.nn.Linear(n_hid, n_hid); self.fc3 = torch.nn.Linear(n_hid, 1)
    def forward(self, x):
        return self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))
torch.manual_seed(2)
net = RegNet().to(dev)
est = KFAC(net)
xs = torch.sort(torch.rand(30, 1, device=dev) * 8 - 4, 0).values
ys = xs ** 3 + 3 * torch.rand(30, 1, device=dev)
loss = torch.nn.functional.mse_loss(net(xs), ys); net.zero_grad(); loss.backward(); est.update(30)
xt = torch.linspace(-6, 6, 100, device=dev).reshape(-1, 1)
t_reg = ev(lambda: linearised_kfac_regression(est, xt, 0.01, 30, 3.0), reps=5, warm=1)
print(f"cfg2 Net(1,1,50): linearised KFAC regression predictive, 100 test points: {t_reg[1]:.2f} ms", flush=True)
