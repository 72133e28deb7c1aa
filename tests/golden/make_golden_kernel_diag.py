"""Golden fixtures for the kernel-block-diagonal helpers (second half), produced by
RUNNING THE REFERENCE's own functions `sampling_free/utils.py:42-211` (generate_diag, generate_H,
generate_kernel_diag_748 / _141 / generate_kernel_diag) on seeded Fisher-like matrices, fp64, CPU.

    python tests/golden/make_golden_kernel_diag.py      # build container only (needs /root/reference)

Inputs are stored as the stacked gradient rows G [n, P] (H = G^T G / n is rebuilt by the tests in fp64); the
outputs (mostly zeros) are stored as float32, compressed.  The only shim is the pair of empty matplotlib modules."""
import importlib.util
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
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
warnings.filterwarnings("ignore")

spec = importlib.util.spec_from_file_location("ref_sf_utils", REF + "/sampling_free/utils.py")
U = importlib.util.module_from_spec(spec)
spec.loader.exec_module(U)

CASES = {  # name -> (P, n gradient rows, tau, n scale, n_hid)
    "kd748": (748, 96, 0.01, 1.0, None),
    "kd141": (141, 64, 0.01, 30.0, None),
    "kdreg5": (46, 40, 0.01, 30.0, 5),       # Net(1, 1, 5): 5 + 5 + 25 + 5 + 5 + 1 parameters
    "kdreg30": (1021, 72, 0.01, 30.0, 30),  # regression_ll_kernel.py:134 (overlapping first-layer blocks)
}


def main():
    store = {}
    for name, (P, n, tau, scale, n_hid) in CASES.items():
        gen = torch.Generator().manual_seed(sum(map(ord, name)))
        G = (0.3 * torch.randn(n, P, generator=gen, dtype=torch.float64)).float().double()
        H = G.t() @ G / n
        store[f"{name}_G"] = G.numpy().astype(np.float32)
        store[f"{name}_meta"] = np.array([P, n, tau, scale, -1 if n_hid is None else n_hid], dtype=np.float64)
        Hc = H.clone()
        if name == "kd748":
            res, inv = U.generate_kernel_diag_748(Hc, tau)
        elif name == "kd141":
            res, inv = U.generate_kernel_diag_141(Hc, tau, scale)
        else:
            res, inv = U.generate_kernel_diag(Hc, tau, scale, n_hid)
        assert torch.allclose(Hc, H + tau * torch.eye(P, dtype=torch.float64))   # the in-place side effect
        store[f"{name}_res"] = res.numpy().astype(np.float32)
        store[f"{name}_inv"] = inv.numpy().astype(np.float32)
        if name == "kd141":
            d_res, d_inv = U.generate_diag(H.clone(), tau)
            store[f"{name}_diag_res"] = d_res.numpy()
            store[f"{name}_diag_inv"] = d_inv.numpy()
            h_reg, h_inv = U.generate_H(H.clone(), tau)
            store[f"{name}_H_reg"] = h_reg.numpy()
            store[f"{name}_H_inv"] = h_inv.numpy()
            store[f"{name}_dominance"] = np.array(U.calculate_dominance(H.clone()))
    np.savez_compressed(OUT / "reference_golden_kernel_diag.npz", **store)
    print("wrote", OUT / "reference_golden_kernel_diag.npz", {k: v.shape for k, v in store.items()})


if __name__ == "__main__":
    main()
