"""Development diagnostic: per-layer factor / inverse errors of cfg1 vs the fp64 oracle."""
import sys
import torch
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import kfac_oracle as O
from bnn_kfac_b200.curvatures import KFAC, invert_factors
from bnn_kfac_b200.wrapper import MLP as WMLP
from test_gpu_parity import _oracle_vs_gpu, _layers

dev = torch.device("cuda:0")
def relerr(a, b):
    return ((a.double().cpu() - b.double().cpu()).norm() / b.double().cpu().norm()).item()

g = torch.Generator().manual_seed(1234)
x = [torch.rand(256, 1, 28, 28, generator=g) for _ in range(2)]
for prec in ["fp32", "bf16x3", "bf16"]:
    for damping in [(0.04, 200.0), (1.0, 200.0)]:
        cm, gm, oest, gest = _oracle_vs_gpu(lambda: WMLP([784, 1024, 1024, 10]), x, dev, *damping, precision=prec)
        for li, (ol, gl) in enumerate(zip(oest.layers, _layers(gest))):
            for k in range(2):
                F = gest.state[gl][k]
                stage = O.kfac_invert_factor(F.double().cpu(), *damping)
                (L32,) = invert_factors([oest.state[ol][k].float().to(dev)], [damping[0]], [damping[1]])
                R = damping[1] ** 0.5 * oest.state[ol][k] + damping[0] ** 0.5 * torch.eye(F.shape[0], dtype=torch.float64)
                cond = torch.linalg.cond(R).item()
                print(f"{prec:7s} {damping} layer{li} k{k} d={F.shape[0]:5d} cond={cond:.2e} factor={relerr(F, oest.state[ol][k]):.2e} "
                      f"e2e_inv={relerr(gest.inv_state[gl][k], oest.inv_state[ol][k]):.2e} "
                      f"chol_stage={relerr(gest.inv_state[gl][k], stage):.2e} "
                      f"oraclefactor_gpuchol={relerr(L32, oest.inv_state[ol][k]):.2e}", flush=True)
