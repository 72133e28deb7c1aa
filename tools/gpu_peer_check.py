"""Peer-memory exchange (csrc/bk_peer.cu) against the NCCL route, under torchrun: the reduce-scatter of the cfg5
factors to their owners and the return of the Cholesky factors must give the same tensors either way; times both
(barrier before every repetition, device events, max over ranks) and the whole invert_sharded."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from bnn_kfac_b200 import distributed as D
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.wrapper import MLP
widths = [int(v) for v in os.environ.get("BK_WIDTHS", "4096,4096,4096,4096,10").split(",")]
n = 4096 if widths[0] >= 2048 else 256
model = MLP(widths).to(dev)
est = KFAC(model, precision="bf16")
layers = [l for _, l in est._selected_layers()]
g = torch.Generator().manual_seed(1 + rank)
for l, (i, o) in zip(layers, zip(widths[:-1], widths[1:])):
    est.record[l] = [torch.randn(n, i, generator=g).to(dev), (torch.randn(n, o, generator=g) / n).to(dev)]
est.update(n)
factors = [f for _, v in est._raw_items() for f in v]
dims = [f.shape[0] for f in factors]
owners = D.plan_owners(dims, world)


def timed(fn, reps=7):
    ms = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(t.item())
    return out, min(ms), sorted(ms)[len(ms) // 2]


def route(peer):
    os.environ.pop("BK_NO_PEER", None)
    os.environ.pop("BK_PEER_PUSH", None)
    if peer == "push":
        os.environ["BK_PEER_PUSH"] = "1"
    elif not peer:
        os.environ["BK_NO_PEER"] = "1"


res = {}
for peer in (False, True, "push", False, True):
    route(peer)
    D.reduce_scatter_to_owners(est, owners)
    red, t_min, t_med = timed(lambda: D.reduce_scatter_to_owners(est, owners))
    mine = sorted(red)
    owned = {i: torch.tril(red[i]).contiguous() for i in mine}
    D.allgather_cholesky(owned, dims, owners, dev)
    gat, g_min, g_med = timed(lambda: D.allgather_cholesky(owned, dims, owners, dev))
    D.invert_sharded(est, 1.0, 200.0)
    _, i_min, i_med = timed(lambda: D.invert_sharded(est, 1.0, 200.0))
    res[peer] = (red, gat, [est.inv_state[l][k].clone() for l in layers for k in range(2)])
    if rank == 0:
        print(f"world={world} {'PEER(push)' if peer == 'push' else 'PEER' if peer else 'NCCL'}: reduce-scatter {t_min:.3f} / {t_med:.3f} ms, "
              f"return of the Cholesky factors {g_min:.3f} / {g_med:.3f} ms, invert_sharded {i_min:.2f} / {i_med:.2f} ms "
              f"(min / median)", flush=True)
bad = 0.0
for i in res[True][0]:
    a, b = res[True][0][i], res[False][0][i]
    bad = max(bad, ((a - b).norm() / b.norm()).item())
    assert torch.equal(a, a.T), "peer route: factor not symmetric"
gat_err = max(((a - b).abs().max()).item() for a, b in zip(res[True][1], res[False][1]))
inv_err = max(((a - b).norm() / b.norm()).item() for a, b in zip(res[True][2], res[False][2]))
ctx = D.PeerExchange.get(0, dev)
err = ctx.error() if ctx is not None else -1
t = torch.tensor([bad, gat_err, inv_err, float(err)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    vol = sum(d * (d + 1) // 2 for d in dims) * 4 * (world - 1) / world
    print(f"peer vs NCCL: reduced factors rel. diff {t[0].item():.2e}, gathered Cholesky factors max abs diff "
          f"{t[1].item():.2e}, inv_state rel. diff {t[2].item():.2e}, wait error word {int(t[3].item())}; "
          f"{vol / 1e6:.0f} MB arrive at / leave every rank per exchange", flush=True)
    ok = t[0].item() < 1e-6 and t[1].item() < 1e-5 and t[2].item() < 1e-4 and int(t[3].item()) == 0
    print("peer check ok" if ok else "peer check FAILED", flush=True)
dist.destroy_process_group()
