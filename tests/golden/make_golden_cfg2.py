"""BASELINE config 2 fixtures: the reference's sampling-free regression predictive, run by THE REFERENCE ITSELF in
fp64 (TianmingQiu/BNN_KFAC mounted at /root/reference; build container only):

    python tests/golden/make_golden_cfg2.py        ->  tests/golden/reference_golden_cfg2.npz

For n_hid in {30, 50} (the script's 30 and BASELINE.json's "50-unit hidden layers"): the data, the network and the
loop of sampling_free/regression/regression_ll_block.py:84-140 executed line for line on the reference's own
`KFAC` (models/curvatures.py) with every tensor in float64 — 40 optimisation steps instead of 10 000 keep the
fixture fast and still exercise `state +=` — and the predictive std of 25 test points on linspace(-6, 6).
Also recorded: cond(N (F + tau I)) of every factor (1e5 .. 5e6: why fp32 cannot meet 1e-3 here) and the factors of
ONE further update at the final parameters (checks `KFAC.update` on this net without the inversion).
"""
import sys
import types
import warnings
from pathlib import Path

import numpy as np
import torch

REF = "/root/reference"
OUT = Path(__file__).resolve().parent

for name in ("matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.path.insert(0, REF)
warnings.filterwarnings("ignore")

from models.curvatures import KFAC  # noqa: E402
from models.utilities import kron  # noqa: E402


class RegNet(torch.nn.Module):
    """Net(input_dim=1, output_dim=1, n_hid) of regression_ll_block.py:23-34."""

    def __init__(self, n_hid):
        super().__init__()
        self.fc1 = torch.nn.Linear(1, n_hid)
        self.fc2 = torch.nn.Linear(n_hid, n_hid)
        self.fc3 = torch.nn.Linear(n_hid, 1)

    def forward(self, x):
        return self.fc3(torch.relu(self.fc2(torch.relu(self.fc1(x)))))


def npy(t):
    return t.detach().cpu().numpy()


def run(n_hid, store):
    torch.manual_seed(2)                                            # :84
    N, sigma, tau = 30, 3, 0.01                                     # :88-90
    x = torch.FloatTensor(30, 1).uniform_(-4, 4).sort(dim=0).values  # :91
    y = x.pow(3) + sigma * torch.rand(x.size())                     # :92
    x, y = x.double(), y.double()
    net = RegNet(n_hid)
    for layer in net.modules():                                     # weight_init_uniform(0.2), :97
        if isinstance(layer, torch.nn.Linear):
            torch.nn.init.uniform_(layer.weight, -0.2, 0.2)
            layer.bias.data.fill_(0)
    net = net.double()
    opt = torch.optim.SGD(net.parameters(), lr=1e-3)                # :99
    kf = KFAC(net)                                                  # :102
    for _ in range(40):                                             # :104-110 (10 000 there)
        loss = torch.nn.functional.mse_loss(net(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        kf.update(batch_size=1)
    p = f"reg{n_hid}_"
    store[p + "x"], store[p + "y"] = npy(x), npy(y)
    store.update({p + "param_" + k: npy(v) for k, v in net.state_dict().items()})
    layers = [l for l in list(kf.model.modules())[1:] if l in kf.state]
    conds = []
    for li, layer in enumerate(layers):
        store[p + f"state_{li}_A"] = npy(kf.state[layer][0])
        store[p + f"state_{li}_G"] = npy(kf.state[layer][1])
        for f in kf.state[layer]:
            conds.append(torch.linalg.cond(N * (f + tau * torch.eye(f.shape[0], dtype=f.dtype))).item())
    store[p + "conds"] = np.array(conds)
    x_ = torch.unsqueeze(torch.linspace(-6, 6, 25), dim=1).double()  # :115 (100 points there)
    store[p + "xtest"] = npy(x_)
    stds = []
    for x_j in x_:                                                  # :120-140
        pred_j = net(x_j)
        std_j = 0
        for layer in list(kf.model.modules())[1:]:
            gl = []
            if layer in kf.state:
                q_i, h_i = kf.state[layer]
                dq = torch.diag(tau * torch.ones(q_i.shape[0], dtype=q_i.dtype))
                dh = torch.diag(tau * torch.ones(h_i.shape[0], dtype=h_i.dtype))
                q_inv = torch.pinverse(N * (q_i + dq))
                h_inv = torch.pinverse(N * (h_i + dh))
                for prm in layer.parameters():
                    gl.append(torch.flatten(torch.autograd.grad(pred_j, [prm], retain_graph=True)[0]))
                J_i = torch.cat(gl, dim=0).unsqueeze(0)
                std_j += torch.abs(J_i @ kron(q_inv, h_inv) @ J_i.t()).item()
        stds.append(std_j ** 0.5 + sigma)
    store[p + "pred_std"] = np.array(stds)
    store[p + "pred_mean"] = npy(net(x_)).squeeze(1)
    # one more update on a fresh estimator at the final parameters
    kf1 = KFAC(net)
    loss = torch.nn.functional.mse_loss(net(x), y)
    net.zero_grad()
    loss.backward()
    kf1.update(batch_size=1)
    for li, layer in enumerate(layers):
        store[p + f"step_{li}_A"] = npy(kf1.state[layer][0])
        store[p + f"step_{li}_G"] = npy(kf1.state[layer][1])
    for h in kf.hooks + kf1.hooks:
        h.remove()


if __name__ == "__main__":
    store = {}
    for n_hid in (30, 50):
        run(n_hid, store)
        print(n_hid, "cond max %.2e" % store[f"reg{n_hid}_conds"].max(), "std-3:", store[f"reg{n_hid}_pred_std"][:4] - 3)
    np.savez_compressed(OUT / "reference_golden_cfg2.npz", **store)
    print("wrote", OUT / "reference_golden_cfg2.npz")
