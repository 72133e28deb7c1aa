"""Golden fixtures for SURVEY.md §8(f) rows f4 (INF) and f3 (calibration metrics), produced by RUNNING THE
REFERENCE's own code (models/curvatures.py:476-682, models/utilities.py:178-366) on the seeded MLP of
make_golden.py, fp64, CPU.

    python tests/golden/make_golden_inf.py        # build container only (needs /root/reference)

Shims, all outside the reference's code: empty matplotlib modules and `torch.symeig` (as in
make_golden_next.py), and ONE method: `INF._dim_reduction` (models/curvatures.py:615-658) indexes a tensor
with a Python list of 0-dim tensors (`lambda_vec[[idx - 1 for idx in idx_top_lm]]`), which torch >= 2 rejects
("too many indices").  It is replaced by the same statements with the list entries converted to Python ints;
nothing else of INF (update, _diagonal_accumulator, invert, pre_sampler, sampler, sample) is touched."""
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

from models import utilities as U  # noqa: E402
from models.curvatures import EFB, INF, KFAC, Diagonal  # noqa: E402
from make_golden import MLP, layers_of, npy  # noqa: E402


def _dim_reduction_int_indices(frst_eigvecs, scnd_eigvecs, lambda_vec, rank):
    if rank >= lambda_vec.shape[0]:
        return frst_eigvecs, scnd_eigvecs, lambda_vec
    m = scnd_eigvecs.shape[1]
    idx_total = torch.argsort(-torch.abs(lambda_vec)) + 1
    idx_top_l = idx_total[0:rank]
    idx_left, idx_right = [], []
    for z in range(rank):
        i = int((idx_top_l[z] - 1.) / m + 1.)
        j = idx_top_l[z] - (m * (i - 1))
        idx_left.append(i)
        idx_right.append(j)
    idx_left = torch.unique(torch.tensor(idx_left))
    idx_right = torch.unique(torch.tensor(idx_right))
    idx_top_lm = [int(m * (idx_left[i] - 1) + idx_right[j]) for i in range(len(idx_left)) for j in range(len(idx_right))]
    return (frst_eigvecs[:, [int(i) - 1 for i in idx_left]], scnd_eigvecs[:, [int(j) - 1 for j in idx_right]],
            lambda_vec[[i - 1 for i in idx_top_lm]])


INF._dim_reduction = staticmethod(_dim_reduction_int_indices)


def main():
    torch.set_num_threads(4)
    gold = dict(np.load(OUT / "reference_golden.npz"))
    store = {}
    dtype = torch.float64
    xs = [torch.tensor(gold[f"mlp_x_{i}"]).to(dtype) for i in range(2)]
    ys = [torch.tensor(gold[f"mlp_y_{i}"]) for i in range(2)]
    crit = torch.nn.CrossEntropyLoss()
    m = MLP().to(dtype)
    m.load_state_dict({k: torch.tensor(gold[f"mlp_param_{k}"]).to(dtype) for k in m.state_dict().keys()})

    kf, dg = KFAC(m), Diagonal(m)
    for x, y in zip(xs, ys):
        loss = crit(m(x), y)
        m.zero_grad()
        loss.backward()
        kf.update(batch_size=x.shape[0])
        dg.update(batch_size=x.shape[0])
    factors = {layer: kf.state[layer] for layer in layers_of(kf)}
    efb = EFB(m, factors)
    for x, y in zip(xs, ys):
        loss = crit(m(x), y)
        m.zero_grad()
        loss.backward()
        efb.update(batch_size=x.shape[0])
    for li, layer in enumerate(layers_of(efb)):
        store[f"inf_UA_{li}"] = npy(efb.eigvecs[layer][0])
        store[f"inf_UG_{li}"] = npy(efb.eigvecs[layer][1])
        store[f"inf_lambdas_{li}"] = npy(efb.state[layer])
        store[f"inf_diags_{li}"] = npy(dg.state[layer])

    for rank in (10, 30):
        inf = INF(m, dg.state, factors, efb.state)
        inf.update(rank=rank)
        for li, layer in enumerate(layers_of(inf)):
            for name, t in zip(("lrA", "lrG", "lrlam", "corr"), inf.state[layer]):
                store[f"inf_r{rank}_{name}_{li}"] = npy(t)       # before invert() clamps corr in place
        inf.invert(0.04, 200.0)
        for li, layer in enumerate(layers_of(inf)):
            a, b, c, p = inf.inv_state[layer]
            store[f"inf_r{rank}_ric_{li}"] = npy(c)
            store[f"inf_r{rank}_P_{li}"] = npy(p)
            torch.manual_seed(5000 + 10 * rank + li)
            smp = inf.sample(layer)
            torch.manual_seed(5000 + 10 * rank + li)
            z = torch.randn(a.shape[0] * b.shape[0], dtype=dtype)
            store[f"inf_r{rank}_z_{li}"], store[f"inf_r{rank}_sample_{li}"] = npy(z), npy(smp)

    # ---------------------------------------------------------------- calibration metrics (f3)
    g = torch.Generator().manual_seed(77)
    logits = 2.5 * torch.randn(600, 10, generator=g)
    labels = torch.randint(0, 10, (600,), generator=g)
    # make the predictions informative so accuracy / calibration are non-trivial
    logits[torch.arange(600), labels] += 2.0 * torch.rand(600, generator=g)
    probs = torch.softmax(logits, 1).numpy()           # float32, what wrapper.eval hands the scripts
    probs[7] = np.eye(10, dtype=np.float32)[3]         # a one-hot row: confidence exactly 1, 0 log 0
    lab = labels.numpy()
    store["met_probs"], store["met_labels"] = probs, lab
    store["met_accuracy"] = np.float64(U.accuracy(probs, lab))
    store["met_confidence"] = np.float64(U.confidence(probs))
    store["met_confidence_rows"] = U.confidence(probs, mean=False)
    store["met_nll"] = np.float64(U.negative_log_likelihood(probs, lab))
    store["met_entropy_rows"] = U.predictive_entropy(probs)
    store["met_entropy_mean"] = np.float64(U.predictive_entropy(probs, mean=True))
    for bins in (10, 15):
        ece, ace, acc, conf = U.expected_calibration_error(probs, lab, bins=bins)
        store[f"met_ece{bins}"] = np.float64(ece)
        store[f"met_ece{bins}_ace"], store[f"met_ece{bins}_acc"], store[f"met_ece{bins}_conf"] = ace, acc, conf
    for bins in (20, 7):
        ece, xs_, ys_, zs_ = U.calibration_curve(probs, lab, bins=bins)
        store[f"met_curve{bins}"] = np.float64(ece)
        store[f"met_curve{bins}_x"], store[f"met_curve{bins}_y"], store[f"met_curve{bins}_z"] = xs_, ys_, zs_
    d1 = np.abs(np.random.default_rng(5).normal(size=4000)).astype(np.float32) * 0.3
    d2 = np.abs(np.random.default_rng(6).normal(size=3000)).astype(np.float32) * 0.5
    store["met_kl_d1"], store["met_kl_d2"] = d1, d2
    store["met_kl"] = np.float64(U.binned_kl_distance(d1, d2))

    np.savez_compressed(OUT / "reference_golden_inf.npz", **store)
    print(f"wrote {len(store)} arrays, {(OUT / 'reference_golden_inf.npz').stat().st_size / 1024:.1f} KiB")


if __name__ == "__main__":
    main()
