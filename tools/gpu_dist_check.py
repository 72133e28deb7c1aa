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


def _ev_ms(fn, dev, warm=True):
    """Device time of the second call of fn(), max over ranks, after a barrier (collectives inside must start
    together; the first call pays one-off costs: workspace growth, autograd's batched-gradient tracing)."""
    if warm:
        fn()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return out, t.item()


def run_check_rows(rank, world, dev, dense_p=3000):
    """SURVEY 8(e) rows 4-6 under NCCL against the single-GPU result, with a device-timed figure each:
    linearised predictive with the test inputs sharded, Diagonal update with the batch sharded (mean gradient
    all-reduced before squaring), dense Fisher with the gradient rows split (reduce-scatter by row block,
    dominance from the shards, bordered blocked Cholesky on the sharded rows)."""
    from bnn_kfac_b200 import dense, dense_sharded
    from bnn_kfac_b200.curvatures import Diagonal
    from bnn_kfac_b200.predictive import linearised_diag, linearised_kfac_regression
    out = {}
    # ---- e4: sampling-free regression predictive (regression_ll_block.py:120-140), test points sharded
    torch.manual_seed(2)
    net = MLP([1, 50, 50, 1]).to(dev)      # the toy regression net of BASELINE config 2
    est = KFAC(net, seed=3)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(30, 1, generator=g) * 8 - 4).to(dev)
    y = x ** 3
    loss = torch.nn.functional.mse_loss(net(x), y)
    net.zero_grad()
    loss.backward()
    est.update(30)
    xt = torch.linspace(-6, 6, 100, device=dev).view(-1, 1)
    got, ms = _ev_ms(lambda: D.linearised_kfac_regression_sharded(est, xt, 0.01, 30.0, 3.0), dev)
    ref = linearised_kfac_regression(est, xt, 0.01, 30.0, 3.0)
    out["linearised_regression"] = {"relerr": relerr(got, ref), "ms": ms, "test_points": 100}
    for h in est.hooks:
        h.remove()
    # ---- e6: Diagonal, batch sharded
    widths = [300, 520, 260, 10]
    n_per = 64
    gen = torch.Generator().manual_seed(99)
    X = torch.rand(world * n_per, widths[0], generator=gen).to(dev)
    Y = torch.randint(0, 10, (world * n_per,), generator=gen).to(dev)
    model = build(dev, widths)
    dg = Diagonal(model)
    loss = torch.nn.functional.cross_entropy(model(X[rank * n_per:(rank + 1) * n_per]), Y[rank * n_per:(rank + 1) * n_per])
    model.zero_grad()
    loss.backward()
    _, ms = _ev_ms(lambda: D.diagonal_update_sharded(dg, n_per), dev, warm=False)     # accumulates: once
    ref_model = build(dev, widths)
    rd = Diagonal(ref_model)
    loss = torch.nn.functional.cross_entropy(ref_model(X), Y)
    ref_model.zero_grad()
    loss.backward()
    rd.update(world * n_per)
    worst = max(relerr(a, b) for a, b in zip(dg.state.values(), rd.state.values()))
    dg.invert(1.0, 10.0)
    rd.invert(1.0, 10.0)
    P = sum(p.numel() for p in model.parameters())
    Jall = torch.randn(40, P, generator=torch.Generator().manual_seed(7)).to(dev)
    a, b = D.row_slice(40, world, rank)
    v = D.linearised_diag_sharded(dg, Jall[a:b].contiguous(), 40)
    worst = max(worst, relerr(v, linearised_diag(rd, Jall)))
    out["diagonal"] = {"relerr": worst, "update_ms": ms, "params": P}
    # ---- e5: dense Fisher, gradient rows split
    P, n, B, tau = dense_p, 64 * world, 8 * world, 0.04
    gen = torch.Generator().manual_seed(13)
    G = (0.3 * torch.randn(n, P, generator=gen)).to(dev)
    J = torch.randn(B, P, generator=gen).to(dev)
    ga, gb = D.row_slice(n, world, rank)
    ja, jb = D.row_slice(B, world, rank)
    sh, ms_acc = _ev_ms(lambda: dense_sharded.dense_fisher_sharded(G[ga:gb].contiguous(), n), dev)
    coords = [(i, min(i + 100, P)) for i in range(0, P, 100)]
    dom, ms_dom = _ev_ms(lambda: sh.dominance(coords, 1e-5), dev)
    var, ms_var = _ev_ms(lambda: sh.variance(J[ja:jb].contiguous(), tau, n_rows=B), dev)
    def single_gpu():
        H1 = dense.dense_fisher(G)
        return H1, dense.dense_variance(J, dense.dense_inverse(H1, tau))
    (H, var1), ms_single = _ev_ms(single_gpu, dev)
    dom1 = dense.dominance(H, coords, 1e-5)
    out["dense_fisher"] = {"relerr": relerr(var, var1), "dominance_err": max(abs(dom[0] - dom1[0]), abs(dom[1] - dom1[1])),
                           "accumulate_exchange_ms": ms_acc, "dominance_ms": ms_dom, "cholesky_variance_ms": ms_var,
                           "single_gpu_replicated_ms": ms_single,
                           "P": P, "gradients": n, "test_rows": B}
    ok = (out["linearised_regression"]["relerr"] < 1e-3 and out["diagonal"]["relerr"] < 1e-3
          and out["dense_fisher"]["relerr"] < 1e-3 and out["dense_fisher"]["dominance_err"] < 1e-5)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["ok"] = bool(flag.item() > 0.5)
    out["tolerance"] = 1e-3
    return out


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.backends.cuda.matmul.allow_tf32 = False
    res = run_check(rank, world, dev)
    rows = run_check_rows(rank, world, dev, dense_p=int(os.environ.get("BK_DENSE_P", "3000")))
    if rank == 0:
        print("rows", rows, flush=True)
    res["ok"] = res["ok"] and rows["ok"]
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
