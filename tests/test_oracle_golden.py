"""Pins the CPU oracle (oracle/kfac_oracle.py) against outputs of the reference itself
(tests/golden/reference_golden.npz, produced by tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import ROOT, relerr
from models_for_tests import MLP, RegNet, load_params
from oracle import kfac_oracle as O
from bnn_kfac_b200.wrapper import BaseNet_750

TOL64 = 1e-10   # fp64 restatement vs fp64 reference run: same algorithm, same LAPACK
TOL32 = 2e-4    # fp32


def _run_kfac(model, xs, ys):
    est = O.OracleKFAC(model)
    for x, y in zip(xs, ys):
        O.fisher_backward(model, x, labels=y)
        est.update()
    return est


@pytest.mark.parametrize("dtype,tag,tol", [(torch.float64, "mlp64", TOL64), (torch.float32, "mlp32", TOL32)])
def test_mlp_factors_inverse_samples(golden, dtype, tag, tol):
    model = load_params(MLP(), golden, "mlp", dtype)
    xs = [torch.tensor(golden[f"mlp_x_{i}"]).to(dtype) for i in range(2)]
    ys = [torch.tensor(golden[f"mlp_y_{i}"]) for i in range(2)]
    est = _run_kfac(model, xs, ys)
    est.invert(0.04, 200.0)
    for li, layer in enumerate(est.layers):
        assert relerr(est.state[layer][0], golden[f"{tag}_state_{li}_A"]) < tol
        assert relerr(est.state[layer][1], golden[f"{tag}_state_{li}_G"]) < tol
        # the inverse is conditioned ~1e3: allow cond * eps in fp32
        itol = tol if dtype == torch.float64 else 5e-3
        assert relerr(est.inv_state[layer][0], golden[f"{tag}_inv_{li}_A"]) < itol
        assert relerr(est.inv_state[layer][1], golden[f"{tag}_inv_{li}_G"]) < itol
        for s in range(2):
            z = torch.tensor(golden[f"{tag}_z_{s}_{li}"])
            ref_inv = (torch.tensor(golden[f"{tag}_inv_{li}_A"]), torch.tensor(golden[f"{tag}_inv_{li}_G"]))
            assert relerr(O.kfac_sample(ref_inv[0], ref_inv[1], z), golden[f"{tag}_sample_{s}_{li}"]) < tol
        # known-answer: bias augmentation makes A[-1, -1] == number of updates (curvatures.py:346-349)
        assert abs(est.state[layer][0][-1, -1].item() - 2.0) < 1e-6


def test_mlp_sample_and_replace_and_mc_mean(golden):
    model = load_params(MLP(), golden, "mlp")
    xs = [torch.tensor(golden[f"mlp_x_{i}"]).double() for i in range(2)]
    ys = [torch.tensor(golden[f"mlp_y_{i}"]) for i in range(2)]
    est = _run_kfac(model, xs, ys)
    est.invert(0.04, 200.0)
    noise = [[torch.tensor(golden[f"mlp64_sar_z_{s}_{li}"]) for li in range(2)] for s in range(3)]
    for s in range(3):
        est.sample_and_replace(noise[s])
        for li, layer in enumerate(est.layers):
            assert relerr(layer.weight.data, golden[f"mlp64_sar_w_{s}_{li}"]) < TOL64
            assert relerr(layer.bias.data, golden[f"mlp64_sar_b_{s}_{li}"]) < TOL64
    mean = O.mc_predict_classification(model, est, torch.tensor(golden["mlp_xtest"]), noise)
    assert relerr(mean, golden["mlp64_mc_mean"]) < TOL64


def test_per_layer_damping_lists(golden):
    model = load_params(MLP(), golden, "mlp")
    xs = [torch.tensor(golden[f"mlp_x_{i}"]).double() for i in range(2)]
    ys = [torch.tensor(golden[f"mlp_y_{i}"]) for i in range(2)]
    est = _run_kfac(model, xs, ys)
    est.invert([1.0, 0.5], [200.0, 100.0])
    for li, layer in enumerate(est.layers):
        assert relerr(est.inv_state[layer][0], golden[f"mlp64_listinv_{li}_A"]) < TOL64
        assert relerr(est.inv_state[layer][1], golden[f"mlp64_listinv_{li}_G"]) < TOL64


def test_diagonal(golden):
    model = load_params(MLP(), golden, "mlp")
    est = O.OracleDiagonal(model)
    for i in range(2):
        x, y = torch.tensor(golden[f"mlp_x_{i}"]).double(), torch.tensor(golden[f"mlp_y_{i}"])
        O.fisher_backward(model, x, labels=y)
        est.update(batch_size=x.shape[0])
    est.invert(0.04, 200.0)
    for li, layer in enumerate(est.layers):
        assert relerr(est.state[layer], golden[f"diag64_state_{li}"]) < TOL64
        assert relerr(est.inv_state[layer], golden[f"diag64_inv_{li}"]) < TOL64
        z = torch.tensor(golden[f"diag64_z_{li}"])
        assert relerr(O.diag_sample(est.inv_state[layer], z), golden[f"diag64_sample_{li}"]) < TOL64
    h = O.diag_flat_inverse(est.layers, est.inv_state)
    xt = torch.tensor(golden["mlp_xtest"])
    pred = torch.softmax(model(xt), dim=1)
    J = O.params_jacobian_flat(pred, model, O.argmax_grad_outputs(pred))
    assert relerr(J.detach(), golden["diag64_lin_J"]) < TOL64
    assert abs(O.linearised_diag_variance(J.detach(), h) / float(golden["diag64_lin_var"]) - 1) < 1e-9


def test_conv_factors_and_linearised_predictive(golden):
    model = load_params(BaseNet_750(), golden, "cnn")
    xs = [torch.tensor(golden[f"cnn_x_{i}"]).double() for i in range(2)]
    ys = [torch.tensor(golden[f"cnn_y_{i}"]) for i in range(2)]
    est = _run_kfac(model, xs, ys)
    est.invert(0.04, 200.0)
    for li, layer in enumerate(est.layers):
        assert relerr(est.state[layer][0], golden[f"cnn64_state_{li}_A"]) < TOL64
        assert relerr(est.state[layer][1], golden[f"cnn64_state_{li}_G"]) < TOL64
        assert relerr(est.inv_state[layer][0], golden[f"cnn64_inv_{li}_A"]) < 1e-8
        assert relerr(est.inv_state[layer][1], golden[f"cnn64_inv_{li}_G"]) < 1e-8
        # upper triangle of the Cholesky factors is exactly zero
        assert torch.triu(est.inv_state[layer][0], 1).abs().max().item() == 0.0
    xt = torch.tensor(golden["cnn_xtest"])
    for use_kron in (True, False):
        pm, pstd, ent = O.linearised_classification_batch(model, est.layers, est.inv_state, xt, use_kron)
        assert relerr(pm, golden["cnn64_lin_pred_mean"]) < TOL64
        assert abs(pstd / float(golden["cnn64_lin_pred_std"]) - 1) < 1e-8
        assert abs(ent - float(golden["cnn64_lin_entropy"])) < 1e-8
    # per-layer terms, kron-free identity J (Q (x) H) J^T = <V, Q V H^T>  (quirk Q2 layout)
    mods = list(model.modules())[1:]
    for li, layer in enumerate(mods):
        key = f"cnn64_lin_J_{li}"
        if key in golden:
            J = torch.tensor(golden[key])
            Q, H = est.inv_state[layer]
            assert abs(O.linearised_kfac_variance(J, Q, H) / float(golden[f"cnn64_lin_term_{li}"]) - 1) < 1e-8


def test_regression_linearised(golden):
    model = load_params(RegNet(30), golden, "reg", torch.float32)
    layers = O.selected_layers(model)
    state = {l: (torch.tensor(golden[f"reg_state_{i}_A"]), torch.tensor(golden[f"reg_state_{i}_G"]))
             for i, l in enumerate(layers)}
    xt = torch.tensor(golden["reg_xtest"])
    for use_kron in (True, False):
        stds = [O.linearised_regression_point(model, layers, state, x_j, 0.01, 30, 3, use_kron) for x_j in xt]
        # fp32 pinverse of matrices conditioned ~1e4: compare the data-dependent part loosely
        np.testing.assert_allclose(np.array(stds) - 3, golden["reg_pred_std"] - 3, rtol=5e-2, atol=1e-4)
    assert relerr(model(xt).detach().squeeze(1), golden["reg_pred_mean"]) < 1e-6


@pytest.mark.parametrize("n_hid", [30, 50])
def test_regression_linearised_cfg2_fp64(n_hid):
    """BASELINE config 2 (n_hid 30 = the script, 50 = BASELINE.json): the oracle against the reference's own fp64
    run (tests/golden/make_golden_cfg2.py).  fp64 on both sides: agreement to rounding, both kron forms."""
    g = dict(np.load(ROOT / "tests" / "golden" / "reference_golden_cfg2.npz"))
    model = load_params(RegNet(n_hid), g, f"reg{n_hid}", torch.float64)
    layers = O.selected_layers(model)
    state = {l: (torch.tensor(g[f"reg{n_hid}_state_{i}_A"]), torch.tensor(g[f"reg{n_hid}_state_{i}_G"]))
             for i, l in enumerate(layers)}
    xt = torch.tensor(g[f"reg{n_hid}_xtest"])
    for use_kron in (True, False):
        stds = [O.linearised_regression_point(model, layers, state, x_j, 0.01, 30, 3, use_kron) for x_j in xt[::4]]
        np.testing.assert_allclose(np.array(stds) - 3, g[f"reg{n_hid}_pred_std"][::4] - 3, rtol=1e-7)
    # one update at the final parameters: oracle factors == the reference's
    x, y = torch.tensor(g[f"reg{n_hid}_x"]), torch.tensor(g[f"reg{n_hid}_y"])
    oest = O.OracleKFAC(model)
    loss = torch.nn.functional.mse_loss(model(x), y)
    model.zero_grad()
    loss.backward()
    oest.update()
    for i, l in enumerate(oest.layers):
        assert relerr(oest.state[l][0], g[f"reg{n_hid}_step_{i}_A"]) < 1e-12
        assert relerr(oest.state[l][1], g[f"reg{n_hid}_step_{i}_G"]) < 1e-12


def test_kron_doctest_vector(golden):
    """The reference's only executable check: models/utilities.py:400-407."""
    out = O.kron(torch.tensor(golden["kron_a"]), torch.tensor(golden["kron_b"]))
    assert np.array_equal(out.numpy(), golden["kron_out"])
    assert np.array_equal(out.numpy(), np.array([[0, 5, 0, 10], [6, 7, 12, 14], [0, 15, 0, 20], [18, 21, 24, 28]]))


def test_invert_identity_and_eigh():
    torch.manual_seed(0)
    x = torch.randn(40, 12, dtype=torch.float64)
    F_ = x.t() @ x / 40
    L = O.kfac_invert_factor(F_, 0.04, 200.0)
    R = 200.0 ** 0.5 * F_ + 0.04 ** 0.5 * torch.eye(12, dtype=torch.float64)
    assert relerr(L @ L.t() @ R, torch.eye(12)) < 1e-10
    wa, va, wg, vg = O.factor_eigenvectors(F_, F_)
    assert relerr(va @ torch.diag(wa) @ va.t(), 2 * F_) < 1e-10   # eigen-decomposition of F + F^T
    ev = O.factor_eigenvalues(F_, F_)
    assert ev.shape == (144,)


def test_dense_fisher_and_dominance():
    torch.manual_seed(1)
    coords = O.kernel_block_coords_basenet15k()
    assert coords[-1][1] == 15080 and len(coords) == 109   # hessian/utils.py:67-95
    g = torch.randn(8, 50, dtype=torch.float64)
    H = O.dense_fisher(g)
    assert relerr(H, sum(torch.outer(r, r) for r in g) / 8) < 1e-12
    dd, kd = O.dominance(H, [(0, 10), (10, 50)])
    assert 0 < dd <= kd <= 1
    J = torch.randn(1, 50, dtype=torch.float64)
    assert O.dense_variance(J, O.dense_inverse(H, 0.04)) > 0


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors)."""
    out = O.philox4x32_10(np.zeros((1, 4), dtype=np.uint32), (0, 0))
    assert [hex(v) for v in out[0]] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    out = O.philox4x32_10(np.full((1, 4), 0xFFFFFFFF, dtype=np.uint32), (0xFFFFFFFF, 0xFFFFFFFF))
    assert [hex(v) for v in out[0]] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    z = O.philox_normal(1234, 0, 7, 200000)
    assert abs(z.mean()) < 0.01 and abs(z.var() - 1) < 0.02


# ------------------------------------------------------------------ "next" rows: BlockDiagonal, EFB
@pytest.fixture(scope="module")
def golden_next():
    return dict(np.load(ROOT / "tests" / "golden" / "reference_golden_next.npz"))


def _mlp_with_grads(golden, i):
    model = load_params(MLP(), golden, "mlp", torch.float64)
    x = torch.tensor(golden[f"mlp_x_{i}"]).double()
    y = torch.tensor(golden[f"mlp_y_{i}"])
    loss = torch.nn.functional.cross_entropy(model(x), y)
    model.zero_grad()
    loss.backward()
    return model, x.shape[0]


def test_blockdiagonal_oracle_vs_reference(golden, golden_next):
    states = [None, None]
    for i in range(2):
        model, bs = _mlp_with_grads(golden, i)
        layers = O.selected_layers(model)
        states = [O.blockdiag_update(states[li], l, bs) for li, l in enumerate(layers)]
    for li, l in enumerate(layers):
        np.testing.assert_allclose(states[li].numpy(), golden_next[f"bd_state_{li}"], rtol=1e-12, atol=1e-14)
        inv = O.blockdiag_invert(states[li], 0.5, 10.0)
        np.testing.assert_allclose(inv.numpy(), golden_next[f"bd_inv_a_{li}"], rtol=1e-9, atol=1e-12)
        smp = O.blockdiag_sample(inv, l, torch.tensor(golden_next[f"bd_z_{li}"]))
        np.testing.assert_allclose(smp.numpy(), golden_next[f"bd_sample_{li}"], rtol=1e-9, atol=1e-12)
        pinv = O.blockdiag_invert(states[li], 0.0, 1.0)
        assert relerr(pinv.numpy(), golden_next[f"bd_inv_pinv_{li}"]) < 1e-6


def test_efb_oracle_vs_reference(golden, golden_next):
    eig = [(torch.tensor(golden_next[f"efb_UA_{li}"]), torch.tensor(golden_next[f"efb_UG_{li}"])) for li in range(2)]
    state, diags = [None, None], [None, None]
    for i in range(2):
        model, bs = _mlp_with_grads(golden, i)
        layers = O.selected_layers(model)
        for li, l in enumerate(layers):
            state[li], diags[li] = O.efb_update(state[li], diags[li], l, eig[li], bs)
    for li in range(2):
        np.testing.assert_allclose(state[li].numpy(), golden_next[f"efb_state_{li}"], rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose(diags[li].numpy(), golden_next[f"efb_diags_{li}"], rtol=1e-12, atol=1e-16)
        inv = O.efb_invert(state[li], 0.04, 200.0)
        np.testing.assert_allclose(inv.numpy(), golden_next[f"efb_inv_{li}"], rtol=1e-10)
        smp = O.efb_sample(eig[li], inv, torch.tensor(golden_next[f"efb_z_{li}"]))
        np.testing.assert_allclose(smp.numpy(), golden_next[f"efb_sample_{li}"], rtol=1e-9, atol=1e-12)
    # the oracle's own eigenvectors (linalg.eigh of F + F^T) span the same eigenspaces: lambdas agree
    A = torch.tensor(golden["mlp64_state_0_A"])
    G = torch.tensor(golden["mlp64_state_0_G"])
    _, va, _, vg = O.factor_eigenvectors(A, G)
    model, bs = _mlp_with_grads(golden, 0)
    l0 = O.selected_layers(model)[0]
    lam_own, _ = O.efb_update(None, None, l0, (va, vg), bs)
    lam_ref, _ = O.efb_update(None, None, l0, eig[0], bs)
    # squares of projections are invariant to eigenvector signs; degenerate eigenvalues (rank-deficient
    # factors) leave the basis of the null space free, so compare the sums over the free blocks: totals
    assert abs(lam_own.sum().item() / lam_ref.sum().item() - 1) < 1e-9


# ------------------------------------------------------------------ "next" rows: INF (f4), metrics (f3)
@pytest.fixture(scope="module")
def golden_inf():
    return dict(np.load(ROOT / "tests" / "golden" / "reference_golden_inf.npz"))


@pytest.mark.parametrize("rank", [10, 30])
def test_inf_oracle_vs_reference(golden_inf, rank):
    g = golden_inf
    for li in range(2):
        eig = (torch.tensor(g[f"inf_UA_{li}"]), torch.tensor(g[f"inf_UG_{li}"]))
        state = O.inf_update(eig, torch.tensor(g[f"inf_lambdas_{li}"]), torch.tensor(g[f"inf_diags_{li}"]), rank)
        for name, t in zip(("lrA", "lrG", "lrlam", "corr"), state):
            # the reference accumulates sif_diag in a float32 buffer (torch.zeros(n*m), curvatures.py:675)
            # whatever the factors' dtype: its correction carries that rounding (~1e-10 absolute here)
            atol = 1e-9 if name == "corr" else 1e-14   # one fp32 ulp of sif_diag
            np.testing.assert_allclose(t.numpy(), g[f"inf_r{rank}_{name}_{li}"], rtol=1e-9, atol=atol)
        # stage-wise from here: feed the reference's own correction so that a flipped fp32 rounding of
        # sif_diag is not amplified by the (ill-conditioned) pre-sampler
        state = state[:3] + (torch.tensor(g[f"inf_r{rank}_corr_{li}"]),)
        inv = O.inf_invert(state, 0.04, 200.0)
        np.testing.assert_allclose(inv[2].numpy(), g[f"inf_r{rank}_ric_{li}"], rtol=1e-10)
        assert relerr(inv[3].numpy(), g[f"inf_r{rank}_P_{li}"]) < 1e-7      # five LAPACK inverses, cond 1e5..1e7
        ref_inv = (inv[0], inv[1], inv[2], torch.tensor(g[f"inf_r{rank}_P_{li}"]))
        smp = O.inf_sample(ref_inv, torch.tensor(g[f"inf_r{rank}_z_{li}"]))
        np.testing.assert_allclose(smp.numpy(), g[f"inf_r{rank}_sample_{li}"], rtol=1e-9, atol=1e-11)


def test_inf_presampler_closed_form(golden_inf):
    """The identity the CUDA path uses: L_c = A^-T (I - B^-1) A^-1 with A = chol(vtv), B = chol(vtv + I)
    equals the reference's (C^-1 + vtv)^-1 (see bk_inf.cu)."""
    g = golden_inf
    li, rank = 1, 10
    a, b = torch.tensor(g[f"inf_r{rank}_lrA_{li}"]), torch.tensor(g[f"inf_r{rank}_lrG_{li}"])
    c = torch.tensor(g[f"inf_r{rank}_ric_{li}"])
    s = (200.0 * torch.tensor(g[f"inf_r{rank}_lrlam_{li}"])).sqrt()
    v = c.view(-1, 1) * O.kron(a, b) @ torch.diag(s)
    vtv = v.t() @ v
    eye = torch.eye(vtv.shape[0], dtype=vtv.dtype)
    ainv = torch.linalg.inv(torch.linalg.cholesky(vtv))
    binv = torch.linalg.inv(torch.linalg.cholesky(vtv + eye))
    p = torch.diag(s) @ ainv.t() @ (eye - binv) @ ainv @ torch.diag(s)
    assert relerr(p.numpy(), g[f"inf_r{rank}_P_{li}"]) < 1e-8


def test_metrics_oracle_vs_reference(golden_inf):
    g = golden_inf
    p, lab = g["met_probs"], g["met_labels"]
    rows = O.metric_rows(p, lab)
    assert abs(O.metric_accuracy(p, lab) - g["met_accuracy"]) < 1e-12
    assert abs(float(rows["conf"].mean()) - g["met_confidence"]) < 1e-7
    np.testing.assert_array_equal(rows["conf"], g["met_confidence_rows"])
    assert abs(O.metric_nll(p, lab) - g["met_nll"]) < 1e-6
    np.testing.assert_allclose(rows["entropy"], g["met_entropy_rows"], rtol=1e-5, atol=1e-7)
    for bins in (10, 15):
        ece, ace, acc, cf = O.metric_ece(p, lab, bins)
        assert abs(ece - g[f"met_ece{bins}"]) < 1e-7
        np.testing.assert_allclose(ace, g[f"met_ece{bins}_ace"], atol=1e-6)
        np.testing.assert_allclose(acc, g[f"met_ece{bins}_acc"], atol=1e-6)
        np.testing.assert_allclose(cf, g[f"met_ece{bins}_conf"], atol=1e-6)
    for bins in (20, 7):
        ece, xs, ys, zs = O.metric_calibration_curve(p, lab, bins)
        assert abs(ece - g[f"met_curve{bins}"]) < 1e-7
        np.testing.assert_allclose(xs, g[f"met_curve{bins}_x"], atol=1e-6)
        np.testing.assert_allclose(ys, g[f"met_curve{bins}_y"], atol=1e-6)
        np.testing.assert_allclose(zs, g[f"met_curve{bins}_z"], atol=1e-12)
    assert abs(O.metric_binned_kl(g["met_kl_d1"], g["met_kl_d2"]) - g["met_kl"]) < 1e-9


@pytest.mark.parametrize("rank", [10, 30, 1000])
def test_inf_dim_reduction_product_host_logic(golden_inf, rank):
    """The product's index bookkeeping (top-k / unique / gather in torch, `INF._dim_reduction`) selects the same
    eigen-directions as the reference (fixtures) and the oracle; it is plain torch and runs on CPU tensors."""
    from bnn_kfac_b200.curvatures import INF
    g = golden_inf
    for li in range(2):
        ua, ug = torch.tensor(g[f"inf_UA_{li}"]), torch.tensor(g[f"inf_UG_{li}"])
        lam = torch.tensor(g[f"inf_lambdas_{li}"]).t().contiguous().view(-1)
        got = INF._dim_reduction(ua, ug, lam, rank)
        want = O.inf_dim_reduction(ua, ug, lam, rank)
        for a, b in zip(got, want):
            assert a.shape == b.shape and torch.equal(a, b)
        if rank < lam.numel():
            np.testing.assert_array_equal(got[0].numpy(), g[f"inf_r{rank}_lrA_{li}"])
            np.testing.assert_array_equal(got[2].numpy(), g[f"inf_r{rank}_lrlam_{li}"])


def test_staged_reference_matches_oracle_port():
    """oracle/_ref (the unmodified reference modules staged by oracle/make_ref.py, what `bench.py --impl
    reference` times) and the oracle port compute the same factors from the same hook records."""
    from oracle import make_ref
    if not make_ref.available():
        if not make_ref.REF.exists():
            pytest.skip("oracle/_ref not staged and /root/reference absent")
        make_ref.make()
    ref = make_ref.load()
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(12, 9), torch.nn.ReLU(), torch.nn.Linear(9, 4)).double()
    est = ref.KFAC(model)
    x = torch.rand(16, 12, dtype=torch.float64)
    loss = torch.nn.functional.cross_entropy(model(x), torch.randint(0, 4, (16,)))
    loss.backward()
    recs = {l: [r[0].detach().clone(), r[1].detach().clone()] for l, r in est.record.items()}
    est.update(16)
    est.update(16)      # state +=
    for layer, (a, g) in recs.items():
        f1, f2 = O.kfac_linear_factors(a, g, True)
        assert torch.allclose(est.state[layer][0], 2 * f1, atol=1e-12)
        assert torch.allclose(est.state[layer][1], 2 * f2, atol=1e-12)


@pytest.mark.parametrize("name", ["kd748", "kd141", "kdreg5", "kdreg30"])
def test_kernel_diag_oracle_vs_reference(name):
    """oracle.kernel_diag / kernel_coords against outputs of the reference's generate_kernel_diag* functions
    (tests/golden/make_golden_kernel_diag.py)."""
    g = dict(np.load(ROOT / "tests" / "golden" / "reference_golden_kernel_diag.npz"))
    P, n, tau, scale, n_hid = g[f"{name}_meta"]
    P, n_hid = int(P), (None if n_hid < 0 else int(n_hid))
    G = torch.tensor(g[f"{name}_G"]).double()
    H = G.t() @ G / G.shape[0]
    H0 = H.clone()
    res, inv = O.kernel_diag(H, O.kernel_coords(P, n_hid), float(tau), float(scale))
    assert torch.equal(H, H0 + float(tau) * torch.eye(P, dtype=torch.float64))     # in-place side effect
    assert relerr(res, g[f"{name}_res"]) < 1e-6
    assert relerr(inv, g[f"{name}_inv"]) < 1e-5
    if name == "kd141":
        d_res, d_inv = O.diag_approximation(H0, float(tau))
        assert relerr(d_res, g["kd141_diag_res"]) < 1e-12 and relerr(d_inv, g["kd141_diag_inv"]) < 1e-12
        assert relerr(O.dense_inverse(H0, float(tau)), g["kd141_H_inv"]) < 1e-6
        assert abs(O.dominance(H0, [], 1e-5)[0] - float(g["kd141_dominance"])) < 1e-12
