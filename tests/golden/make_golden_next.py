"""Golden fixtures for the "next" rows of SURVEY.md §8(f): BlockDiagonal and EFB, produced by RUNNING
THE REFERENCE's own classes (models/curvatures.py:210-275, 408-473) on the seeded MLP of
make_golden.py, fp64, CPU.

    python tests/golden/make_golden_next.py        # build container only (needs /root/reference)

Two shims, both outside the reference's code: empty matplotlib modules (as in make_golden.py) and
`torch.symeig`, which the reference calls (models/utilities.py:155-157) but torch >= 1.13 no longer
has; it is provided here as a thin wrapper over torch.linalg.eigh (same ascending order, same
eigenvectors-as-columns layout)."""
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
sys.path.insert(0, str(OUT))
warnings.filterwarnings("ignore")


def _symeig(x, eigenvectors=False, upper=True):
    w, v = torch.linalg.eigh(x, UPLO="U" if upper else "L")
    return w, (v if eigenvectors else torch.empty(0, dtype=x.dtype))


torch.symeig = _symeig

from models.curvatures import EFB, KFAC, BlockDiagonal  # noqa: E402
from make_golden import MLP, layers_of, npy  # noqa: E402


def main():
    torch.set_num_threads(4)
    gold = dict(np.load(OUT / "reference_golden.npz"))
    store = {}
    dtype = torch.float64
    xs = [torch.tensor(gold[f"mlp_x_{i}"]).to(dtype) for i in range(2)]
    ys = [torch.tensor(gold[f"mlp_y_{i}"]) for i in range(2)]
    crit = torch.nn.CrossEntropyLoss()

    def fresh():
        m = MLP().to(dtype)
        m.load_state_dict({k: torch.tensor(gold[f"mlp_param_{k}"]).to(dtype) for k in m.state_dict().keys()})
        return m

    # ---------------------------------------------------------------- BlockDiagonal
    m = fresh()
    bd = BlockDiagonal(m)
    for x, y in zip(xs, ys):
        loss = crit(m(x), y)
        m.zero_grad()
        loss.backward()
        bd.update(batch_size=x.shape[0])
    for add, mult, tag in ((0.5, 10.0, "a"), (0.0, 1.0, "pinv")):
        bd.inv_state = dict()
        bd.invert(add, mult)
        for li, layer in enumerate(layers_of(bd)):
            store[f"bd_state_{li}"] = npy(bd.state[layer])
            store[f"bd_inv_{tag}_{li}"] = npy(bd.inv_state[layer])
            if tag == "a":
                torch.manual_seed(3000 + li)
                smp = bd.sample(layer)
                torch.manual_seed(3000 + li)
                z = bd.inv_state[layer].new(bd.inv_state[layer].shape[0]).normal_()
                store[f"bd_z_{li}"], store[f"bd_sample_{li}"] = npy(z), npy(smp)

    # ---------------------------------------------------------------- EFB on top of KFAC factors
    m = fresh()
    kf = KFAC(m)
    for x, y in zip(xs, ys):
        loss = crit(m(x), y)
        m.zero_grad()
        loss.backward()
        kf.update(batch_size=x.shape[0])
    factors = {layer: kf.state[layer] for layer in layers_of(kf)}
    efb = EFB(m, factors)
    for x, y in zip(xs, ys):
        loss = crit(m(x), y)
        m.zero_grad()
        loss.backward()
        efb.update(batch_size=x.shape[0])
    efb.invert(0.04, 200.0)
    for li, layer in enumerate(layers_of(efb)):
        store[f"efb_UA_{li}"] = npy(efb.eigvecs[layer][0])
        store[f"efb_UG_{li}"] = npy(efb.eigvecs[layer][1])
        store[f"efb_state_{li}"] = npy(efb.state[layer])
        store[f"efb_diags_{li}"] = npy(efb.diags[layer])
        store[f"efb_inv_{li}"] = npy(efb.inv_state[layer])
        torch.manual_seed(4000 + li)
        smp = efb.sample(layer)
        torch.manual_seed(4000 + li)
        first, second = efb.eigvecs[layer]
        z = torch.randn(first.size(0), second.size(0), dtype=dtype)
        store[f"efb_z_{li}"], store[f"efb_sample_{li}"] = npy(z), npy(smp)
    np.savez_compressed(OUT / "reference_golden_next.npz", **store)
    print(f"wrote {len(store)} arrays, {(OUT / 'reference_golden_next.npz').stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
