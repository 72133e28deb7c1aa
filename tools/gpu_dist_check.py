"""Multi-GPU check (run under torchrun, one rank per GPU): the sharded path must reproduce the
single-GPU result.  rank r accumulates factors on its own batch shard; invert_sharded reduces to
owners, inverts and broadcasts; mc_predict_sharded shards the posterior samples.  Rank 0 then replays
the whole thing alone on the full batch and compares."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bnn_kfac_b200 import distributed as D  # noqa: E402
from bnn_kfac_b200.curvatures import KFAC  # noqa: E402
from bnn_kfac_b200.predictive import mc_predict  # noqa: E402
from bnn_kfac_b200.wrapper import MLP  # noqa: E402


def relerr(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-300)).item()


def build(dev, widths, seed=0):
    torch.manual_seed(seed)
    m = MLP(widths)
    m.weight_init_uniform(0.05)
    return m.to(dev)


def run_check(rank, world, dev):
    """Runs inside an initialised NCCL process group.  Returns {"inverse_relerr", "predictive_relerr",
    "identical_on_all_ranks", "ok"} (the error figures are rank 0's; every rank gets the same `ok`)."""
    widths = [300, 520, 260, 10]
    n_per, steps = 64, 2
    g = torch.Generator().manual_seed(1234)
    X = torch.rand(steps, world * n_per, widths[0], generator=g)
    Y = torch.randint(0, 10, (steps, world * n_per), generator=g)
    xt = torch.rand(48, widths[0], generator=g).to(dev)

    model = build(dev, widths)
    est = KFAC(model, seed=11)
    for s in range(steps):
        xb = X[s, rank * n_per:(rank + 1) * n_per].to(dev)
        yb = Y[s, rank * n_per:(rank + 1) * n_per].to(dev)
        # per-shard mean loss; hooks rescale by the LOCAL batch size (models/curvatures.py:322-323)
        loss = torch.nn.functional.cross_entropy(model(xb), yb)
        model.zero_grad()
        loss.backward()
        est.update(n_per)
    D.invert_sharded(est, 1e2, 1e4)
    S = 2 * world + 1
    pred = D.mc_predict_sharded(est, xt, S)
    worst, e_pred = 0.0, 0.0
    if rank == 0:
        ref_model = build(dev, widths)
        ref = KFAC(ref_model, seed=11)
        for s in range(steps):
            loss = torch.nn.functional.cross_entropy(ref_model(X[s].to(dev)), Y[s].to(dev))
            ref_model.zero_grad()
            loss.backward()
            ref.update(world * n_per)
        ref.invert(1e2, 1e4)
        ref_pred = mc_predict(ref, xt, S)
        layers = [l for _, l in est._selected_layers()]
        rlayers = [l for _, l in ref._selected_layers()]
        for l, rl in zip(layers, rlayers):
            for k in range(2):
                worst = max(worst, relerr(est.inv_state[l][k], ref.inv_state[rl][k]))
        e_pred = relerr(pred, ref_pred)
        for h in ref.hooks:
            h.remove()
    for h in est.hooks:
        h.remove()
    # every rank must hold the same inverse factors and prediction
    chk = torch.stack([est.inv_state[l][k].double().sum() for _, l in est._selected_layers() for k in range(2)]
                      + [pred.double().sum()])
    lo, hi = chk.clone(), chk.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(lo, hi))
    flag = torch.tensor([1.0 if (worst < 1e-3 and e_pred < 1e-3 and same) else 0.0], device=dev)
    dist.broadcast(flag, src=0)
    return {"inverse_relerr": worst, "predictive_relerr": e_pred, "identical_on_all_ranks": same,
            "ok": bool(flag.item() > 0.5) and same, "tolerance": 1e-3,
            "what": "sharded accumulation + invert_sharded + mc_predict_sharded vs one GPU on the full batch "
                    "(MLP 300-520-260-10)"}


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    res = run_check(rank, world, dev)
    if rank == 0:
        print(f"world={world} inverse relerr (worst factor) {res['inverse_relerr']:.2e}  predictive relerr "
              f"{res['predictive_relerr']:.2e}  identical on all ranks: {res['identical_on_all_ranks']}", flush=True)
        if res["ok"]:
            print("dist check ok", flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if res["ok"] else 1)


if __name__ == "__main__":
    main()
