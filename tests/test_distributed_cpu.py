"""World-size-2 gloo tests (CPU) of the multi-GPU host logic in bnn_kfac_b200/distributed.py:
ownership plan, deferred factor reduction, reduce-to-owner / broadcast choreography of the sharded
inversion, sample sharding of the MC predictive, row gather, mean-gradient all-reduce.
The arithmetic between the collectives is injected (oracle functions stand in for the CUDA library),
so these tests check exactly what a second GPU adds: who owns what, and what is exchanged when."""
import os
import socket
import sys
from pathlib import Path

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from bnn_kfac_b200 import distributed as D  # noqa: E402


def test_plan_owners_balances_cubic_cost():
    dims = [4097, 4096, 4097, 4096, 4097, 4096, 4097, 10]
    for world in (1, 2, 4, 8):
        owners = D.plan_owners(dims, world)
        assert len(owners) == len(dims) and all(0 <= o < world for o in owners)
        load = [sum(d ** 3 for d, o in zip(dims, owners) if o == r) for r in range(world)]
        assert max(load) <= sum(load) / world + max(dims) ** 3     # LPT bound
    assert D.plan_owners([5, 3], 4) == [0, 1]
    assert D.plan_owners([], 2) == []


def test_sample_slice_partitions_exactly():
    for n, w in [(100, 8), (30, 4), (7, 2), (8, 8), (5, 1)]:
        got = [D.sample_slice(n, w, r) for r in range(w)]
        assert got[0][0] == 0 and got[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(got, got[1:]))
        sizes = [b - a for a, b in got]
        assert max(sizes) - min(sizes) <= 1


class _FakeEst:
    """The attributes distributed.py touches on an estimator."""

    def __init__(self, state):
        self.state = state
        self.inv_state = {}
        self.invalidated = 0

    def _invalidate_caches(self):
        self.invalidated += 1


def _spd(d, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(3 * d, d, generator=g, dtype=torch.float64)
    return x.t() @ x / (3 * d)


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import kfac_oracle as O
        dims = [(6, 4), (9, 5), (3, 2)]
        layers = [torch.nn.Linear(a - 1, b) for a, b in dims]
        # per-rank partial factors (what each rank's batch shard produced)
        part = {l: [_spd(a, 10 * i + rank), _spd(b, 100 + 10 * i + rank)] for i, (l, (a, b)) in enumerate(zip(layers, dims))}
        full = {l: [sum(_spd(a, 10 * i + r) for r in range(world)) / world,
                    sum(_spd(b, 100 + 10 * i + r) for r in range(world)) / world]
                for i, (l, (a, b)) in enumerate(zip(layers, dims))}

        # 1) deferred all-reduce of the state == mean over ranks (one factor is a 16-byte pitched view,
        #    as curvatures._alloc_factor hands out for wide odd-sized factors)
        est = _FakeEst({l: [t.clone() for t in v] for l, v in part.items()})
        pitched = torch.full((6, 8), float("nan"), dtype=torch.float64)[:, :6]
        pitched.copy_(est.state[layers[0]][0])
        est.state[layers[0]][0] = pitched
        assert not pitched.is_contiguous()
        D.allreduce_state(est)
        for l in layers:
            for k in range(2):
                assert torch.allclose(est.state[l][k], full[l][k], atol=1e-12)

        # 2) sharded inversion: reduce to owners, invert owned, broadcast
        calls = []

        def inverter(fs, adds, mults):
            calls.append([f.shape[0] for f in fs])
            return [O.kfac_invert_factor(f, a, m) for f, a, m in zip(fs, adds, mults)]

        est = _FakeEst({l: [t.clone() for t in v] for l, v in part.items()})
        D.invert_sharded(est, [0.5, 1.0, 2.0], [10.0, 20.0, 30.0], inverter=inverter)
        adds, mults = [0.5, 1.0, 2.0], [10.0, 20.0, 30.0]
        for i, l in enumerate(layers):
            for k in range(2):
                ref = O.kfac_invert_factor(full[l][k], adds[i], mults[i])
                assert torch.allclose(est.inv_state[l][k], ref, atol=1e-10), (rank, i, k)
                # the local partial accumulator is untouched (update() can continue)
                assert torch.equal(est.state[l][k], part[l][k])
            assert isinstance(est.inv_state[l], tuple)
        owners = D.plan_owners([d for ab in dims for d in ab], world)
        mine = sorted(d for d, o in zip([d for ab in dims for d in ab], owners) if o == rank)
        assert sorted(sum(calls, [])) == mine            # each rank inverted exactly what it owns
        assert est.invalidated == 1

        # 3) MC predictive: samples sharded, result independent of the world size
        def moments_fn(est_, x, n, sample0=0, mode="classification", program=None):
            ids = torch.arange(sample0, sample0 + n, dtype=torch.float64)
            p = torch.softmax(x.double() * (1 + ids.view(-1, 1, 1) / 10), dim=-1)   # sample id -> value
            return p.mean(0), (p * p).mean(0)

        x = torch.linspace(-1, 1, 12).view(4, 3)
        S = 7
        got = D.mc_predict_sharded(None, x, S, group=None, moments_fn=moments_fn)
        ref = moments_fn(None, x, S)[0]
        assert torch.allclose(got, ref, atol=1e-12)
        # regression: per-rank (mean, centred second moment) combined with Chan's formula.  The outputs sit on
        # a mean of 200 with a spread of 1e-3 (y = x^3 at the edge of the reference's test range): the
        # E[y^2] - E[y]^2 form would lose this in fp32; the moments travel as float32 like on the device.
        def reg_vals(xx, sample0, n):
            ids = torch.arange(sample0, sample0 + n, dtype=torch.float64).view(-1, 1, 1)
            return 200.0 + xx.double() * 1e-3 * ids

        def reg_moments(e, xx, n, sample0=0, mode="regression_centred", program=None):
            assert mode == "regression_centred"
            v = reg_vals(xx, sample0, n)
            return v.mean(0).float(), v.var(0, unbiased=False).float()

        mean, std = D.mc_predict_sharded(None, x[:, :1], S, mode="regression", moments_fn=reg_moments)
        vals = reg_vals(x[:, :1], 0, S)
        assert torch.allclose(mean.double(), vals.mean(0).squeeze(1), rtol=1e-6)
        assert torch.allclose(std.double(), vals.std(0, unbiased=False).squeeze(1), rtol=2e-2, atol=1e-9)
        with pytest.raises(ValueError):
            D.mc_predict_sharded(None, x, 1, moments_fn=moments_fn)

        # 4) row gather of sharded linearised-predictive results
        n_rows = 5
        a, b = D.row_slice(n_rows, world, rank)
        local = torch.arange(a, b, dtype=torch.float32) * 2
        assert torch.equal(D.gather_rows(local, n_rows), torch.arange(n_rows, dtype=torch.float32) * 2)

        # 5) Diagonal: mean gradient all-reduced before squaring
        m = torch.nn.Linear(3, 2)
        for p in m.parameters():
            p.grad = torch.full_like(p, float(rank + 1))
        D.allreduce_mean_grads(m)
        for p in m.parameters():
            assert torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2))

        # 6) Diagonal update with the batch sharded: all-reduce of the mean gradient, global batch size
        class _Diag:
            def __init__(self, model):
                self.model, self.seen = model, None

            def update(self, batch_size):
                self.seen = (batch_size, [p.grad.clone() for p in self.model.parameters()])

        m = torch.nn.Linear(3, 2)
        for p in m.parameters():
            p.grad = torch.full_like(p, float(2 * rank + 1))
        dg = _Diag(m)
        D.diagonal_update_sharded(dg, 16)
        assert dg.seen[0] == 16 * world
        assert all(torch.allclose(g, torch.full_like(g, float(world))) for g in dg.seen[1])   # mean of 1, 3

        # 7) linearised predictive, test inputs / test batches sharded
        xt = torch.linspace(-6, 6, 7).view(7, 1)
        got = D.linearised_kfac_regression_sharded(None, xt, 0.01, 30.0, 3.0,
                                                   fn=lambda e, x, tau, N, sigma: (x[:, 0] ** 2 * N + sigma))
        assert torch.allclose(got, xt[:, 0] ** 2 * 30.0 + 3.0)
        batches = [torch.full((3 + i, 4), float(i)) for i in range(5)]
        fn = lambda e, x: (x[:, :2] + 1, float(x[0, 0]) * 2, float(x[0, 0]) - 1)   # noqa: E731
        pm, ps, pe = D.linearised_kfac_classification_sharded(None, batches, fn=fn)
        assert torch.equal(pm, torch.cat([b[:, :2] + 1 for b in batches]))
        assert ps.tolist() == [0.0, 2.0, 4.0, 6.0, 8.0] and pe.tolist() == [-1.0, 0.0, 1.0, 2.0, 3.0]
        a5, b5 = D.row_slice(5, world, rank)
        Jl = torch.arange(a5, b5, dtype=torch.float32).view(-1, 1).repeat(1, 3)
        got = D.linearised_diag_sharded(None, Jl, 5, fn=lambda e, J: (J * J).sum(1))
        assert torch.equal(got, 3 * torch.arange(5, dtype=torch.float32) ** 2)
        # more ranks than rows: empty shards still take part in the collectives
        got = D.linearised_kfac_regression_sharded(None, xt[:1], 0.01, 30.0, 3.0,
                                                   fn=lambda e, x, tau, N, sigma: x[:, 0] * 0 + 5.0)
        assert got.tolist() == [5.0]

        # 8) dense Fisher: gradient rows split, reduce-scatter by row block (block-cyclic), dominance from the
        #    shards, bordered blocked Cholesky on the sharded rows for |J (H + tau I)^-1 J^T|
        from bnn_kfac_b200 import dense_sharded as DS

        class TorchOps:
            """fp64 torch stand-ins for the four CUDA operations (test infrastructure only)."""

            def syrk(self, grads, normalise):
                return grads.t() @ grads / normalise

            def chol_trinv(self, blk, add, status):
                a = torch.tril(blk) + torch.tril(blk, -1).t() + add * torch.eye(blk.shape[0], dtype=blk.dtype)
                return torch.linalg.inv(torch.linalg.cholesky(a))

            def gemm_nt(self, a, b, out, alpha, beta):
                if a.shape[0] and b.shape[0]:
                    out.copy_(alpha * (a.clone() @ b.clone().t()) + beta * out)

            def rownorm2(self, y):
                return (y * y).sum(1)

            def dominance_rows(self, rows, row0, P, tau, coords):
                reg = rows[:, :P].clone()
                idx = torch.arange(rows.shape[0])
                reg[idx, row0 + idx] += tau
                blk = sum(reg[max(a, row0) - row0:max(min(b, row0 + rows.shape[0]) - row0, 0), a:b].abs().sum()
                          for a, b in coords)
                return torch.stack([reg[idx, row0 + idx].abs().sum(), reg.abs().sum(),
                                    torch.as_tensor(blk, dtype=reg.dtype)])

        for P, nb in ((23, 4), (16, 4), (5, 8)):
            gen = torch.Generator().manual_seed(77)
            G = torch.randn(12, P, generator=gen, dtype=torch.float64)
            J = torch.randn(5, P, generator=gen, dtype=torch.float64)
            ga, gb = D.row_slice(12, world, rank)
            ja, jb = D.row_slice(5, world, rank)
            sh = DS.dense_fisher_sharded(G[ga:gb], 12, nb=nb, ops=TorchOps())
            H = G.t() @ G / 12
            for l, g in enumerate(sh.mine):                     # each rank holds exactly its block rows of H
                n = min(nb, P - g * nb)
                assert torch.allclose(sh.rows[l * nb:l * nb + n, :P], H[g * nb:g * nb + n], atol=1e-12)
            coords = [(0, 3), (3, 5)] if P > 5 else [(0, 2)]
            d1, d2 = sh.dominance(coords, 1e-5)
            r1, r2 = O.dominance(H, coords, 1e-5)
            assert abs(d1 - r1) < 1e-12 and abs(d2 - r2) < 1e-12
            var = sh.variance(J[ja:jb], 0.04, n_rows=5)
            ref = torch.stack([torch.as_tensor(O.dense_variance(J[i:i + 1], O.dense_inverse(H, 0.04)), dtype=torch.float64)
                               for i in range(5)])
            assert torch.allclose(var, ref.to(var.dtype), rtol=1e-9), (P, nb, var, ref)
        Path(tmp, f"ok{rank}").write_text("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert (tmp_path / "ok0").exists() and (tmp_path / "ok1").exists()
