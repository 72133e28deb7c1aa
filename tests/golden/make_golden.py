"""Generates tests/golden/*.npz by RUNNING THE REFERENCE ITSELF (TianmingQiu/BNN_KFAC, mounted at
/root/reference) on seeded synthetic inputs, on the CPU, in fp64 and fp32.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference imports matplotlib at module import (models/utilities.py:19) and matplotlib is not
installed here, so two empty stub modules are placed in sys.modules first; nothing else of the
reference is altered.  Pieces that only exist inside the reference's scripts (the sampling-free
predictive loops) are executed here line-for-line from the cited script lines on top of the reference's
own `KFAC` / `Diagonal` objects and its own helper functions.
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

from models.curvatures import KFAC, Diagonal  # noqa: E402
from models.utilities import kron  # noqa: E402
from models.wrapper import BaseNet_750  # noqa: E402


class MLP(torch.nn.Module):
    def __init__(self, d_in=20, d_h=16, d_out=5):
        super().__init__()
        self.fc1 = torch.nn.Linear(d_in, d_h)
        self.fc2 = torch.nn.Linear(d_h, d_out)

    def forward(self, x):
        return self.fc2(torch.relu(self.fc1(x)))


class RegNet(torch.nn.Module):
    """Net(input_dim=1, output_dim=1, n_hid) of sampling_free/regression/regression_ll_block.py:23-34."""

    def __init__(self, n_hid=30):
        super().__init__()
        self.fc1 = torch.nn.Linear(1, n_hid)
        self.fc2 = torch.nn.Linear(n_hid, n_hid)
        self.fc3 = torch.nn.Linear(n_hid, 1)

    def forward(self, x):
        x = torch.relu(self.fc1(x))
        x = torch.relu(self.fc2(x))
        return self.fc3(x)


def npy(t):
    return t.detach().cpu().numpy().copy()  # copy: later in-place updates must not alias the fixture


def layers_of(est):
    return [m for m in est.model.modules() if m.__class__.__name__ in ("Linear", "Conv2d")]


def run_kfac(model, batches, labels, dtype, add, multiply, n_samples, tag, store):
    """Reference KFAC: update over batches, invert, draw samples with recorded noise."""
    model = model.to(dtype)
    est = KFAC(model)
    crit = torch.nn.CrossEntropyLoss()
    for i, (x, y) in enumerate(zip(batches, labels)):
        logits = model(x.to(dtype))
        loss = crit(logits, y)
        model.zero_grad()
        loss.backward()
        est.update(batch_size=x.shape[0])
    est.invert(add, multiply)
    lys = layers_of(est)
    for li, layer in enumerate(lys):
        store[f"{tag}_state_{li}_A"] = npy(est.state[layer][0])
        store[f"{tag}_state_{li}_G"] = npy(est.state[layer][1])
        store[f"{tag}_inv_{li}_A"] = npy(est.inv_state[layer][0])
        store[f"{tag}_inv_{li}_G"] = npy(est.inv_state[layer][1])
    # samples: replay torch.randn with a known seed to capture z (curvatures.py:404)
    for s in range(n_samples):
        for li, layer in enumerate(lys):
            torch.manual_seed(1000 + 10 * s + li)
            smp = est.sample(layer)
            torch.manual_seed(1000 + 10 * s + li)
            first, second = est.inv_state[layer]
            z = torch.randn(first.size(0), second.size(0), dtype=first.dtype)
            store[f"{tag}_z_{s}_{li}"] = npy(z)
            store[f"{tag}_sample_{s}_{li}"] = npy(smp)
    return est


def main():
    torch.set_num_threads(4)
    store = {}
    g = torch.Generator().manual_seed(1234)

    # ---------------------------------------------------------------- kron doctest vector
    a = torch.tensor([[1, 2], [3, 4]])
    b = torch.tensor([[0, 5], [6, 7]])
    store["kron_a"], store["kron_b"], store["kron_out"] = npy(a), npy(b), npy(kron(a, b))

    # ---------------------------------------------------------------- MLP, KFAC + Diagonal
    torch.manual_seed(7)
    mlp = MLP()
    store.update({f"mlp_param_{k}": npy(v) for k, v in mlp.state_dict().items()})
    xs = [torch.rand(12, 20, generator=g) for _ in range(2)]
    ys = [torch.randint(0, 5, (12,), generator=g) for _ in range(2)]
    for i in range(2):
        store[f"mlp_x_{i}"], store[f"mlp_y_{i}"] = npy(xs[i]), npy(ys[i])
    for dtype, tag in ((torch.float64, "mlp64"), (torch.float32, "mlp32")):
        m = MLP()
        m.load_state_dict(mlp.state_dict())
        est = run_kfac(m, xs, ys, dtype, 0.04, 200.0, 2, tag, store)
        if dtype == torch.float64:
            # sample_and_replace with known noise -> perturbed weights + MC softmax mean
            xt = torch.rand(6, 20, generator=g).to(dtype)
            store["mlp_xtest"] = npy(xt)
            mean = 0
            for s in range(3):
                torch.manual_seed(2000 + s)
                est.sample_and_replace()
                torch.manual_seed(2000 + s)
                for li, layer in enumerate(layers_of(est)):
                    first, second = est.inv_state[layer]
                    store[f"mlp64_sar_z_{s}_{li}"] = npy(torch.randn(first.size(0), second.size(0), dtype=dtype))
                for li, layer in enumerate(layers_of(est)):
                    store[f"mlp64_sar_w_{s}_{li}"] = npy(layer.weight.data)
                    store[f"mlp64_sar_b_{s}_{li}"] = npy(layer.bias.data)
                with torch.no_grad():
                    mean = mean + torch.softmax(est.model(xt), dim=1)
            store["mlp64_mc_mean"] = npy(mean / 3)
            est.model.load_state_dict(est.model_state)
            # per-layer list damping (curvatures.py:374-376)
            est.inv_state = {}
            est.invert([1.0, 0.5], [200.0, 100.0])
            for li, layer in enumerate(layers_of(est)):
                store[f"mlp64_listinv_{li}_A"] = npy(est.inv_state[layer][0])
                store[f"mlp64_listinv_{li}_G"] = npy(est.inv_state[layer][1])

    m = MLP().double()
    m.load_state_dict(mlp.state_dict())
    diag = Diagonal(m)
    crit = torch.nn.CrossEntropyLoss()
    for x, y in zip(xs, ys):
        loss = crit(m(x.double()), y)
        m.zero_grad()
        loss.backward()
        diag.update(batch_size=x.shape[0])
    diag.invert(0.04, 200.0)
    for li, layer in enumerate(layers_of(diag)):
        store[f"diag64_state_{li}"] = npy(diag.state[layer])
        store[f"diag64_inv_{li}"] = npy(diag.inv_state[layer])
        torch.manual_seed(3000 + li)
        smp = diag.sample(layer)
        torch.manual_seed(3000 + li)
        z = diag.inv_state[layer].new(diag.inv_state[layer].size()).normal_()
        store[f"diag64_z_{li}"], store[f"diag64_sample_{li}"] = npy(z), npy(smp)
    # linearised diagonal predictive (classification_ll_diagonal.py:104-131)
    h = torch.cat([torch.flatten(diag.inv_state[l]) for l in layers_of(diag)], dim=0)
    H_inv = torch.diag(h)
    xt = torch.tensor(store["mlp_xtest"])
    pred_mean = torch.softmax(m(xt), dim=1)
    grad_outputs = torch.zeros_like(pred_mean)
    idx = np.argmax(pred_mean.detach().numpy(), axis=1)
    grad_outputs[:, idx] = 1
    gl = [torch.flatten(torch.autograd.grad(pred_mean, [p], grad_outputs=grad_outputs, retain_graph=True)[0])
          for p in m.parameters()]
    J = torch.cat(gl, dim=0).unsqueeze(0)
    store["diag64_lin_J"] = npy(J)
    store["diag64_lin_var"] = np.array(torch.abs(J * H_inv * J).sum().item())

    # ---------------------------------------------------------------- conv net (reference BaseNet_750)
    torch.manual_seed(11)
    cnn = BaseNet_750()
    cnn.weight_init_uniform(0.2)
    store.update({f"cnn_param_{k}": npy(v) for k, v in cnn.state_dict().items()})
    cx = [torch.rand(6, 1, 28, 28, generator=g) for _ in range(2)]
    cy = [torch.randint(0, 10, (6,), generator=g) for _ in range(2)]
    for i in range(2):
        store[f"cnn_x_{i}"], store[f"cnn_y_{i}"] = npy(cx[i]), npy(cy[i])
    c64 = BaseNet_750()
    c64.load_state_dict(cnn.state_dict())
    est = run_kfac(c64, cx, cy, torch.float64, 0.04, 200.0, 1, "cnn64", store)

    # sampling-free classification loop on top (classification_ll_block.py:114-135), batch of 4
    xt = torch.rand(4, 1, 28, 28, generator=g).double()
    store["cnn_xtest"] = npy(xt)
    pred_mean = torch.softmax(est.model(xt), dim=1)
    pred_std = 0
    idx = np.argmax(pred_mean.detach().numpy(), axis=1)
    grad_outputs = torch.zeros_like(pred_mean)
    grad_outputs[:, idx] = 1
    for li, layer in enumerate(list(est.model.modules())[1:]):
        gl = []
        if layer in est.state:
            Q_i = est.inv_state[layer][0]
            H_i = est.inv_state[layer][1]
            for p in layer.parameters():
                gl.append(torch.flatten(torch.autograd.grad(pred_mean, [p], grad_outputs=grad_outputs,
                                                            retain_graph=True)[0]))
            J_i = torch.cat(gl, dim=0).unsqueeze(0)
            Hk = torch.kron(Q_i, H_i)
            term = torch.abs(J_i @ Hk @ J_i.t()).item()
            store[f"cnn64_lin_J_{li}"] = npy(J_i)
            store[f"cnn64_lin_term_{li}"] = np.array(term)
            pred_std += term
    store["cnn64_lin_pred_mean"] = npy(pred_mean)
    store["cnn64_lin_pred_std"] = np.array(pred_std)
    store["cnn64_lin_entropy"] = np.array(0.5 * np.log2(2 * np.e * np.pi * pred_std))

    # ---------------------------------------------------------------- regression (regression_ll_block.py)
    torch.manual_seed(2)
    N, sigma, tau = 30, 3, 0.01
    x = torch.FloatTensor(30, 1).uniform_(-4, 4).sort(dim=0).values
    y = x.pow(3) + sigma * torch.rand(x.size())
    reg = RegNet(30)
    for layer in reg.modules():
        if isinstance(layer, torch.nn.Linear):
            torch.nn.init.uniform_(layer.weight, -0.2, 0.2)
            layer.bias.data.fill_(0)
    opt = torch.optim.SGD(reg.parameters(), lr=1e-3)
    kf = KFAC(reg)
    for t in range(40):  # the script runs 10 000 steps; 40 keep the fixture fast and exercise `+=`
        loss = torch.nn.functional.mse_loss(reg(x), y)
        opt.zero_grad()
        loss.backward()
        opt.step()
        kf.update(batch_size=1)
    store["reg_x"], store["reg_y"] = npy(x), npy(y)
    store.update({f"reg_param_{k}": npy(v) for k, v in reg.state_dict().items()})
    lys = layers_of(kf)
    for li, layer in enumerate(lys):
        store[f"reg_state_{li}_A"] = npy(kf.state[layer][0])
        store[f"reg_state_{li}_G"] = npy(kf.state[layer][1])
    x_ = torch.unsqueeze(torch.linspace(-6, 6, 7), dim=1)
    store["reg_xtest"] = npy(x_)
    stds = []
    for x_j in x_:
        pred_j = reg(x_j)
        std_j = 0
        for layer in list(kf.model.modules())[1:]:
            gl = []
            if layer in kf.state:
                q_i, h_i = kf.state[layer]
                q_inv = torch.pinverse(N * (q_i + torch.diag(tau * torch.ones(q_i.shape[0]))))
                h_inv = torch.pinverse(N * (h_i + torch.diag(tau * torch.ones(h_i.shape[0]))))
                for p in layer.parameters():
                    gl.append(torch.flatten(torch.autograd.grad(pred_j, [p], retain_graph=True)[0]))
                J_i = torch.cat(gl, dim=0).unsqueeze(0)
                H_inv = kron(q_inv, h_inv)
                std_j += torch.abs(J_i @ H_inv @ J_i.t()).item()
        stds.append(std_j ** 0.5 + sigma)
    store["reg_pred_std"] = np.array(stds)
    store["reg_pred_mean"] = npy(reg(x_)).squeeze(1)

    np.savez_compressed(OUT / "reference_golden.npz", **store)
    size = (OUT / "reference_golden.npz").stat().st_size
    print(f"wrote {len(store)} arrays, {size/1024:.1f} KiB")


if __name__ == "__main__":
    main()
