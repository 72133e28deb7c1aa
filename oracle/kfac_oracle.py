"""CPU ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.

A restatement, in plain CPU torch / numpy, of the algorithm of the reference's Kronecker-factored
Laplace hot path (TianmingQiu/BNN_KFAC).  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module, and only as the checker
or as the timed CPU baseline.  Nothing under `bnn_kfac_b200/` imports it.

Parity pinning: the reference ships no golden vectors (only a `kron` doctest, models/utilities.py:
400-407).  This oracle is pinned against OUTPUTS OF THE REFERENCE ITSELF, executed in the build
container by `tests/golden/make_golden.py` (which imports /root/reference with a matplotlib stub) and
committed as `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks every function below against
those fixtures.  Pieces of the reference that no longer run on a current torch (`torch.symeig`) or
only exist inside scripts are restated from the cited lines and say so.

Every function cites the reference lines it follows (paths relative to the reference repo root).
The arithmetic library is whatever the reference calls: torch (ATen: MKL GEMM, LAPACK getrf/getri/
potrf/gesdd) — not vendored; versions in this image: torch 2.11, numpy 2.3.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor


# =============================================================================== KFAC factors
def kfac_linear_factors(a: Tensor, g_scaled: Tensor, has_bias: bool = True) -> Tuple[Tensor, Tensor]:
    """First / second Kronecker factor of a Linear layer for one mini-batch.

    a: layer input [N, d_in]; g_scaled: grad_output * N [N, d_out] (models/curvatures.py:322-323).
    models/curvatures.py:345-349 (A, with the appended row of ones) and :355-356 (G)."""
    fwd = a.t()
    if has_bias:
        fwd = torch.cat([fwd, torch.ones_like(fwd[:1])], dim=0)
    first = fwd @ fwd.t() / float(fwd.shape[1])
    bwd = g_scaled.t()
    second = bwd @ bwd.t() / float(bwd.shape[1])
    return first, second


def kfac_conv_factors(x: Tensor, g_scaled: Tensor, kernel_size, padding, stride,
                      has_bias: bool = True) -> Tuple[Tensor, Tensor]:
    """Conv2d factors: x [N, C, H, W], g_scaled = grad_output * N [N, O, H', W'].

    models/curvatures.py:341-343 (unfold -> [C*kh*kw, N*L]), :346-349, :353, :356.  Both factors are
    normalised by their own column count (N*L and N*H'*W')."""
    u = F.unfold(x, kernel_size, padding=padding, stride=stride)
    fwd = u.permute(1, 0, 2).contiguous().view(u.shape[1], -1)
    if has_bias:
        fwd = torch.cat([fwd, torch.ones_like(fwd[:1])], dim=0)
    first = fwd @ fwd.t() / float(fwd.shape[1])
    bwd = g_scaled.permute(1, 0, 2, 3).contiguous().view(g_scaled.shape[1], -1)
    second = bwd @ bwd.t() / float(bwd.shape[1])
    return first, second


def kfac_invert_factor(factor: Tensor, add: float, multiply: float) -> Tensor:
    """chol_lower(inverse(sym(sqrt(s) F + sqrt(n) I))).  models/curvatures.py:381-392."""
    d = factor.shape[0]
    reg = multiply ** 0.5 * factor + torch.diag(factor.new_full((d,), add ** 0.5))
    reg = (reg + reg.t()) / 2.0
    return torch.linalg.cholesky(torch.linalg.inv(reg))


def kfac_invert(state: Sequence[Tuple[Tensor, Tensor]], add, multiply) -> List[Tuple[Tensor, Tensor]]:
    """Per-layer damping values as scalars or sequences.  models/curvatures.py:373-398."""
    out = []
    for index, (first, second) in enumerate(state):
        if not isinstance(add, (float, int)) and not isinstance(multiply, (float, int)):
            assert len(add) == len(multiply) == len(state)
            n, s = add[index], multiply[index]
        else:
            n, s = float(add), float(multiply)
        out.append((kfac_invert_factor(first, n, s), kfac_invert_factor(second, n, s)))
    return out


def kfac_sample(chol_first: Tensor, chol_second: Tensor, z: Tensor) -> Tensor:
    """(L_A z L_G^T)^T with z [d_in', d_out] -> [d_out, d_in'].  models/curvatures.py:403-405."""
    return (chol_first @ z @ chol_second.t()).t()


def replace(sample: Tensor, weight: Tensor, bias: Optional[Tensor]) -> Tuple[Tensor, Optional[Tensor]]:
    """Returns (weight + sample[:, :-1], bias + sample[:, -1]).  models/curvatures.py:67-82."""
    if bias is not None:
        new_b = bias + sample[:, -1].contiguous().view(*bias.shape)
        sample = sample[:, :-1]
    else:
        new_b = None
    return weight + sample.contiguous().view(*weight.shape), new_b


# =============================================================================== Diagonal
def diag_update(wgrad: Tensor, bgrad: Optional[Tensor], batch_size: int,
                prev: Optional[Tensor] = None) -> Tensor:
    """state (+)= [W.grad | b.grad]^2 * batch_size.  models/curvatures.py:165-172."""
    grads = wgrad.contiguous().view(wgrad.shape[0], -1)
    if bgrad is not None:
        grads = torch.cat([grads, bgrad.unsqueeze(dim=1)], dim=1)
    grads = grads ** 2 * batch_size
    return grads if prev is None else prev + grads


def diag_invert(state: Tensor, add: float, multiply: float) -> Tensor:
    """reciprocal(s * state + n).sqrt().  models/curvatures.py:202."""
    return torch.reciprocal(multiply * state + add).sqrt()


def diag_sample(inv_state: Tensor, z: Tensor) -> Tensor:
    """z * inv_state.  models/curvatures.py:207."""
    return z * inv_state


# =============================================================================== model-level drivers
def selected_layers(model: torch.nn.Module, layer_types=('Linear', 'Conv2d')) -> List[torch.nn.Module]:
    """Layers in `model.modules()` order, matched by class-name string (models/curvatures.py:121,311)."""
    return [m for m in model.modules() if m.__class__.__name__ in layer_types]


class OracleKFAC:
    """Hook-driven KFAC on a CPU model: same observable behaviour as the reference class
    (models/curvatures.py:295-405), composed from the functions above."""

    def __init__(self, model: torch.nn.Module, averaging: str = "sum", decay: float = 0.95):
        """averaging="sum": the reference (curvatures.py:359-363).  averaging="ema" is NOT in the reference
        (BASELINE.json's north_star asks for a "fused running-average update"): state_1 = F_1,
        state_t = decay * state_{t-1} + (1 - decay) * F_t — the definition the CUDA path is tested against."""
        self.model = model
        self.averaging, self.decay = averaging, decay
        self.layers = selected_layers(model)
        self.record: Dict[torch.nn.Module, list] = {m: [None, None] for m in self.layers}
        self.state: Dict[torch.nn.Module, list] = {}
        self.inv_state: Dict[torch.nn.Module, tuple] = {}
        self.map_state = {k: v.clone() for k, v in model.state_dict().items()}
        self.hooks = []
        for m in self.layers:
            self.hooks.append(m.register_forward_pre_hook(self._fwd))
            self.hooks.append(m.register_full_backward_hook(self._bwd))

    def _fwd(self, module, inp):  # curvatures.py:319-320
        self.record[module][0] = inp[0]

    def _bwd(self, module, grad_input, grad_output):  # curvatures.py:322-323
        self.record[module][1] = grad_output[0] * grad_output[0].size(0)

    def update(self):  # curvatures.py:325-363 (batch_size is unused there)
        for m in self.layers:
            a, g = self.record[m]
            a, g = a.detach(), g.detach()
            if m.__class__.__name__ == 'Conv2d':
                f1, f2 = kfac_conv_factors(a, g, m.kernel_size, m.padding, m.stride, m.bias is not None)
            else:
                f1, f2 = kfac_linear_factors(a, g, m.bias is not None)
            if m in self.state and self.averaging == "ema":
                self.state[m][0] = self.decay * self.state[m][0] + (1 - self.decay) * f1
                self.state[m][1] = self.decay * self.state[m][1] + (1 - self.decay) * f2
            elif m in self.state:
                self.state[m][0] += f1
                self.state[m][1] += f2
            else:
                self.state[m] = [f1, f2]

    def invert(self, add=0., multiply=1.):
        inv = kfac_invert([tuple(v) for v in self.state.values()], add, multiply)
        for m, pair in zip(self.state.keys(), inv):
            self.inv_state[m] = pair

    def sample(self, layer, z: Optional[Tensor] = None) -> Tensor:
        first, second = self.inv_state[layer]
        if z is None:  # curvatures.py:404
            z = torch.randn(first.size(0), second.size(0), dtype=first.dtype)
        return kfac_sample(first, second, z)

    def sample_and_replace(self, zs: Optional[Sequence[Tensor]] = None):
        """curvatures.py:117-129; zs[i] is the noise of the i-th selected layer."""
        self.model.load_state_dict(self.map_state)
        for i, m in enumerate(self.layers):
            s = self.sample(m, None if zs is None else zs[i])
            w, b = replace(s, m.weight.data, None if m.bias is None else m.bias.data)
            m.weight.data.copy_(w)
            if b is not None:
                m.bias.data.copy_(b)

    def remove_hooks(self):
        for h in self.hooks:
            h.remove()


class OracleDiagonal:
    """models/curvatures.py:146-207."""

    def __init__(self, model: torch.nn.Module):
        self.model = model
        self.layers = selected_layers(model)
        self.state: Dict[torch.nn.Module, Tensor] = {}
        self.inv_state: Dict[torch.nn.Module, Tensor] = {}

    def update(self, batch_size: int):
        for m in self.layers:
            bg = None if m.bias is None else m.bias.grad
            self.state[m] = diag_update(m.weight.grad, bg, batch_size, self.state.get(m))

    def invert(self, add=0., multiply=1.):
        for index, (m, v) in enumerate(self.state.items()):
            if isinstance(add, (list, tuple)) and isinstance(multiply, (list, tuple)):
                n, s = add[index], multiply[index]
            else:
                n, s = add, multiply
            self.inv_state[m] = diag_invert(v, n, s)


def fisher_backward(model: torch.nn.Module, x: Tensor, labels: Optional[Tensor] = None,
                    generator: Optional[torch.Generator] = None) -> Tensor:
    """One forward/backward with labels sampled from the model's own predictive (true-Fisher MC-1
    estimate): sampling_free/classification/classification_ll_block.py:93-100.  Returns the labels."""
    logits = model(x)
    if labels is None:
        probs = torch.softmax(logits.detach(), dim=1)
        labels = torch.multinomial(probs, 1, generator=generator).squeeze(1)
    loss = F.cross_entropy(logits, labels)
    model.zero_grad()
    loss.backward()
    return labels


# =============================================================================== predictive loops
def mc_predict_classification(model, oracle: OracleKFAC, x: Tensor,
                              noise: Sequence[Sequence[Tensor]]) -> Tensor:
    """mean over samples of softmax(model_s(x)); noise[s][layer] = z.
    sampling/classification_sampling.py:74-79 with models/wrapper.py:35-44."""
    mean = 0
    with torch.no_grad():
        for zs in noise:
            oracle.sample_and_replace(zs)
            model.eval()
            mean = mean + torch.softmax(model(x), dim=1)
    model.load_state_dict(oracle.map_state)
    return mean / len(noise)


def mc_predict_regression(model, oracle: OracleKFAC, x: Tensor,
                          noise: Sequence[Sequence[Tensor]]) -> Tuple[np.ndarray, np.ndarray]:
    """Per-input mean and std (ddof = 0) over samples.  sampling/regression_sampling.py:81-88."""
    preds = []
    with torch.no_grad():
        for zs in noise:
            oracle.sample_and_replace(zs)
            preds.append(model(x).numpy().squeeze(1))
    model.load_state_dict(oracle.map_state)
    pred = np.array(preds).T
    return pred.mean(axis=1), pred.std(axis=1)


def kron(a: Tensor, b: Tensor) -> Tensor:
    """models/utilities.py:387-409 (einsum Kronecker product); same as sampling_free/utils.py:279-290."""
    return torch.einsum("ab,cd->acbd", a, b).contiguous().view(a.size(0) * b.size(0),
                                                               a.size(1) * b.size(1))


def layer_jacobian_flat(out: Tensor, layer: torch.nn.Module, grad_outputs: Tensor) -> Tensor:
    """J_i = cat(flatten(d out / d p) for p in layer.parameters()) as a row vector.
    sampling_free/classification/classification_ll_block.py:128-130, sampling_free/utils.py:221-226."""
    g = []
    for p in layer.parameters():
        g.append(torch.flatten(torch.autograd.grad(out, [p], grad_outputs=grad_outputs,
                                                   retain_graph=True, allow_unused=True)[0]))
    return torch.cat(g, dim=0).unsqueeze(0)


def argmax_grad_outputs(pred_mean: Tensor) -> Tensor:
    """grad_outputs[:, idx] = 1 with a VECTOR idx: every row gets a one in every column that is the
    arg-max of any row (quirk Q3).  classification_ll_block.py:119-121."""
    idx = np.argmax(pred_mean.detach().numpy(), axis=1)
    grad_outputs = torch.zeros_like(pred_mean)
    grad_outputs[:, idx] = 1
    return grad_outputs


def linearised_kfac_variance_kron(J: Tensor, Q: Tensor, H: Tensor) -> float:
    """|J (Q (x) H) J^T| with the Kronecker product materialised, exactly as the script does.
    classification_ll_block.py:130-132 (torch.kron) / regression_ll_block.py:136-139."""
    return torch.abs(J @ torch.kron(Q, H) @ J.t()).item()


def linearised_kfac_variance(J: Tensor, Q: Tensor, H: Tensor) -> float:
    """Kron-free restatement: with V = J.view(d_in', d_out) (row-major reinterpretation of the flat
    [W.flatten(), b] vector, quirk Q2),  J (Q (x) H) J^T = <V, Q V H^T>_F.  Identical to
    `linearised_kfac_variance_kron` (checked in tests/test_oracle_golden.py)."""
    V = J.reshape(Q.shape[0], H.shape[0])
    return torch.abs((V * (Q @ V @ H.t())).sum()).item()


def linearised_classification_batch(model, layers: Sequence[torch.nn.Module],
                                    inv_state: Dict[torch.nn.Module, Tuple[Tensor, Tensor]],
                                    x: Tensor, use_kron: bool = False) -> Tuple[Tensor, float, float]:
    """One test batch of the sampling-free classification loop: returns (pred_mean, pred_std, entropy).
    classification_ll_block.py:114-135.  Uses inv_state (the Cholesky factors) as Q_i, H_i (quirk Q1)."""
    pred_mean = torch.softmax(model(x), dim=1)
    grad_outputs = argmax_grad_outputs(pred_mean)
    pred_std = 0.0
    for layer in layers:
        Q_i, H_i = inv_state[layer]
        J_i = layer_jacobian_flat(pred_mean, layer, grad_outputs)
        f = linearised_kfac_variance_kron if use_kron else linearised_kfac_variance
        pred_std += f(J_i.detach(), Q_i, H_i)
    entropy = 0.5 * np.log2(2 * np.e * np.pi * pred_std)
    return pred_mean.detach(), pred_std, float(entropy)


def linearised_regression_point(model, layers: Sequence[torch.nn.Module],
                                state: Dict[torch.nn.Module, Sequence[Tensor]], x_j: Tensor,
                                tau: float, N: float, sigma: float, use_kron: bool = False) -> float:
    """Predictive std of one test point of the regression script: q_inv = pinv(N (A + tau I)),
    h_inv = pinv(N (G + tau I)), std = sqrt(sum_layers |J (q_inv (x) h_inv) J^T|) + sigma.
    sampling_free/regression/regression_ll_block.py:120-140."""
    pred_j = model(x_j)
    std_j = 0.0
    for layer in layers:
        q_i, h_i = state[layer]
        # torch.pinverse == linalg.pinv(rcond=1e-15): NO singular value is truncated (the default rtol
        # of linalg.pinv would drop directions of these cond ~1e5 matrices in fp32 and change the result)
        q_inv = torch.pinverse(N * (q_i + tau * torch.eye(q_i.shape[0], dtype=q_i.dtype)))
        h_inv = torch.pinverse(N * (h_i + tau * torch.eye(h_i.shape[0], dtype=h_i.dtype)))
        J_i = layer_jacobian_flat(pred_j, layer, torch.ones_like(pred_j)).detach()
        f = linearised_kfac_variance_kron if use_kron else linearised_kfac_variance
        std_j += f(J_i, q_inv, h_inv)
    return std_j ** 0.5 + sigma


def diag_flat_inverse(layers: Sequence[torch.nn.Module], inv_state: Dict[torch.nn.Module, Tensor]) -> Tensor:
    """h = cat(flatten(inv_state[layer])) in model.modules() order.
    sampling_free/classification/classification_ll_diagonal.py:108-113."""
    return torch.cat([torch.flatten(inv_state[l]) for l in layers], dim=0)


def params_jacobian_flat(out: Tensor, model: torch.nn.Module, grad_outputs: Tensor) -> Tensor:
    """J over net.parameters() order ([W1, b1, W2, b2, ...]).  classification_ll_diagonal.py:127-130."""
    g = []
    for p in model.parameters():
        g.append(torch.flatten(torch.autograd.grad(out, [p], grad_outputs=grad_outputs,
                                                   retain_graph=True, allow_unused=True)[0]))
    return torch.cat(g, dim=0).unsqueeze(0)


def linearised_diag_variance(J: Tensor, h: Tensor) -> float:
    """|J * diag(h) * J|.sum() == sum_j J_j^2 h_j (the script broadcasts against a P x P diagonal
    matrix).  classification_ll_diagonal.py:131; regression_ll_diagonal.py:139."""
    return torch.abs(J.flatten() ** 2 * h.flatten()).sum().item()


# =============================================================================== dense Fisher
def flat_gradient(model: torch.nn.Module) -> Tensor:
    """cat over modules()[1:] of cat(flatten(p.grad)) — the parameter order of the dense scripts.
    hessian/classification_ll_dense_kernel_diag.py:79-84."""
    g = []
    for layer in list(model.modules())[1:]:
        for p in layer.parameters():
            g.append(torch.flatten(p.grad.data))
    return torch.cat(g, dim=0)


def dense_fisher(grads: Tensor) -> Tensor:
    """H = sum_b g_b g_b^T / n_batches for stacked flat gradients [n, P].
    hessian/classification_ll_dense_kernel_diag.py:85-89."""
    return grads.t() @ grads / grads.shape[0]


def kernel_block_coords_basenet15k() -> List[Tuple[int, int]]:
    """Per-kernel diagonal block ranges of BaseNet_15k's flat parameter vector.  hessian/utils.py:67-95."""
    coords, curr = [], 0
    for count, size, bias in ((5, 25, 5), (10, 125, 10), (80, 160, 80), (10, 80, 10)):
        for _ in range(count):
            coords.append((curr, curr + size))
            curr += size
        coords.append((curr, curr + bias))
        curr += bias
    return coords


def dominance(H: Tensor, coords: Sequence[Tuple[int, int]], tau: float = 1e-5) -> Tuple[float, float]:
    """(sum|diag| / sum|all|, sum|kernel blocks| / sum|all|) of H + tau I.  hessian/utils.py:4-23."""
    reg = H + tau * torch.eye(H.shape[0], dtype=H.dtype)
    sum_diag = torch.diag(reg).abs().sum().item()
    sum_all = reg.abs().sum().item()
    sum_block = 0.0
    for (a, b) in coords:
        sum_block += reg[a:b, a:b].abs().sum().item()
    return sum_diag / sum_all, sum_block / sum_all


def dense_inverse(H: Tensor, tau: float) -> Tensor:
    """pinv(H + tau I).  sampling_free/utils.py:47-53; classification_ll_dense.py:108-109."""
    return torch.linalg.pinv(H + tau * torch.eye(H.shape[0], dtype=H.dtype))


def dense_variance(J: Tensor, H_inv: Tensor) -> float:
    """|J H_inv J^T|.  sampling_free/classification/classification_ll_dense.py:160-161."""
    return torch.abs(J @ H_inv @ J.t()).item()


def kernel_coords(P: int, n_hid: Optional[int] = None) -> List[Tuple[int, int]]:
    """Block ranges of sampling_free/utils.py: P = 15080 (:65-93), 748 (:108-129), 141 (:146-167); n_hid given:
    the regression variant (:181-201), whose first loop advances by 1 while its blocks are n_hid wide."""
    if n_hid is not None:
        groups = ((n_hid, n_hid, 1, n_hid), (n_hid, n_hid, n_hid, n_hid), (1, n_hid, n_hid, 1))
    else:
        groups = {15080: ((5, 25, 25, 5), (10, 125, 125, 10), (80, 160, 160, 80), (10, 80, 80, 10)),
                  748: ((3, 9, 9, 3), (6, 27, 27, 6), (10, 54, 54, 10)),
                  141: ((10, 1, 1, 10), (10, 10, 10, 10), (1, 10, 10, 1))}[P]
    coords, curr = [], 0
    for count, size, step, bias in groups:
        for _ in range(count):
            coords.append((curr, curr + size))
            curr += step
        coords.append((curr, curr + bias))
        curr += bias
    return coords


def kernel_diag(H: Tensor, coords: Sequence[Tuple[int, int]], tau: float = 0.0, n: float = 1.0
                ) -> Tuple[Tensor, Tensor]:
    """H += tau I (in place); res[a:b, a:b] = H[a:b, a:b] per block; (res, inverse(n * res)).
    sampling_free/utils.py:95-103 (same body at :131-138, :169-177, :203-211)."""
    res = torch.zeros_like(H)
    H += tau * torch.eye(H.shape[0], dtype=H.dtype)
    for (a, b) in coords:
        res[a:b, a:b] = H[a:b, a:b]
    return res, torch.inverse(n * res)


def diag_approximation(H: Tensor, tau: float = 0.0) -> Tuple[Tensor, Tensor]:
    """(diag(diag(H) + tau), diag(1 / (diag(H) + tau))).  sampling_free/utils.py:42-45."""
    h = torch.diag(H) + tau
    return torch.diag(h), torch.diag(torch.reciprocal(h))


# =============================================================================== eigen-decomposition
def factor_eigenvectors(xxt: Tensor, ggt: Tensor) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Eigen-decomposition of F + F^T (SUM, not mean) in ascending order; restated with
    torch.linalg.eigh because torch.symeig (models/utilities.py:155-157) no longer exists.
    Returns (evals_A, evecs_A, evals_G, evecs_G)."""
    wa, va = torch.linalg.eigh(xxt + xxt.t())
    wg, vg = torch.linalg.eigh(ggt + ggt.t())
    return wa, va, wg, vg


def factor_eigenvalues(xxt: Tensor, ggt: Tensor) -> Tensor:
    """ger(eigvals(A), eigvals(G)).view(-1).  models/utilities.py:136-138 (symeig -> eigvalsh)."""
    return torch.outer(torch.linalg.eigvalsh(xxt), torch.linalg.eigvalsh(ggt)).contiguous().view(-1)


# =============================================================================== RNG restatement
def philox4x32_10(counter: np.ndarray, key: Tuple[int, int]) -> np.ndarray:
    """Philox4x32-10 (Salmon et al., SC'11) on uint32 counters [n, 4]; the generator the CUDA path
    uses in place of torch.randn (models/curvatures.py:404).  Bit-exact integer arithmetic."""
    M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
    W0, W1 = 0x9E3779B9, 0xBB67AE85
    c = counter.astype(np.uint64)
    k0, k1 = key[0] & 0xFFFFFFFF, key[1] & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = M0 * c[:, 0]
        p1 = M1 * c[:, 2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        n0 = hi1 ^ c[:, 1] ^ np.uint64(k0)
        n2 = hi0 ^ c[:, 3] ^ np.uint64(k1)
        c = np.stack([n0, lo1, n2, lo0], axis=1)
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c.astype(np.uint32)


def philox_normal_matrix(seed: int, sample: int, stream_id: int, rows: int, cols: int) -> np.ndarray:
    """The [rows, cols] normal matrix of (seed, sample, stream) as bk_philox_normal defines it
    (include/bk_kfac.h): element (r, c) = lane c % 4 of counter (c // 4, r, sample, stream_id);
    Box-Muller on 24-bit uniforms, evaluated in fp64 (agrees with the device's fast-math fp32 to ~1e-5)."""
    groups = (cols + 3) // 4
    g, r = np.meshgrid(np.arange(groups, dtype=np.uint64), np.arange(rows, dtype=np.uint64))
    ctr = np.stack([g.reshape(-1), r.reshape(-1), np.full(rows * groups, sample, dtype=np.uint64),
                    np.full(rows * groups, stream_id, dtype=np.uint64)], axis=1)
    u = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.float64)
    out = np.empty((rows * groups, 4))
    for a, b, o in ((0, 1, 0), (2, 3, 2)):
        u1 = (np.floor(u[:, a] / 256.0) + 1.0) / 16777216.0
        u2 = np.floor(u[:, b] / 256.0) / 16777216.0
        rad = np.sqrt(-2.0 * np.log(u1))
        out[:, o] = rad * np.cos(2.0 * math.pi * u2)
        out[:, o + 1] = rad * np.sin(2.0 * math.pi * u2)
    return out.reshape(rows, groups * 4)[:, :cols]


def philox_normal(seed: int, sample: int, stream_id: int, count: int) -> np.ndarray:
    """The first `count` normals of (seed, sample, stream): element e uses counter (e // 4, sample,
    stream_id), lane e % 4; Box-Muller on 24-bit uniforms (bk_philox_normal in include/bk_kfac.h).
    fp64 evaluation of the transform: agrees with the device's fast-math fp32 to ~1e-5."""
    groups = (count + 3) // 4
    gi = np.arange(groups, dtype=np.uint64)
    ctr = np.stack([gi & np.uint64(0xFFFFFFFF), gi >> np.uint64(32),
                    np.full(groups, sample, dtype=np.uint64),
                    np.full(groups, stream_id, dtype=np.uint64)], axis=1)
    r = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).astype(np.float64)
    out = np.empty((groups, 4))
    for a, b, o in ((0, 1, 0), (2, 3, 2)):
        u1 = (np.floor(r[:, a] / 256.0) + 1.0) / 16777216.0
        u2 = np.floor(r[:, b] / 256.0) / 16777216.0
        rad = np.sqrt(-2.0 * np.log(u1))
        out[:, o] = rad * np.cos(2.0 * math.pi * u2)
        out[:, o + 1] = rad * np.sin(2.0 * math.pi * u2)
    return out.reshape(-1)[:count]


# =============================================================================== BlockDiagonal / EFB
def layer_grad_matrix(layer: torch.nn.Module) -> Tensor:
    """[d_out, d_in(+1)]: weight.grad viewed [d_out, -1] with bias.grad as last column.
    models/curvatures.py:436-438 (EFB.update)."""
    g = layer.weight.grad.contiguous().view(layer.weight.grad.shape[0], -1)
    if layer.bias is not None:
        g = torch.cat([g, layer.bias.grad.unsqueeze(dim=1)], dim=1)
    return g


def blockdiag_update(state: Optional[Tensor], layer: torch.nn.Module, batch_size: int) -> Tensor:
    """state (+)= ger(g, g) * batch_size with g = [W.grad.view(-1), b.grad].  models/curvatures.py:224-232."""
    g = layer.weight.grad.contiguous().view(-1)
    if layer.bias is not None:
        g = torch.cat([g, layer.bias.grad])
    upd = torch.outer(g, g) * batch_size
    return upd if state is None else state + upd


def blockdiag_invert(state: Tensor, add: float, multiply: float) -> Tensor:
    """pinverse(multiply * state + add * I).  models/curvatures.py:266-268."""
    reg = torch.diag(state.new_full((state.shape[0],), add))
    return torch.pinverse(multiply * state + reg)


def blockdiag_sample(inv: Tensor, layer: torch.nn.Module, z: Tensor) -> Tensor:
    """x = z @ inv reshaped to [d_out, d_in(+1)].  models/curvatures.py:272-275."""
    x = z @ inv
    return torch.cat([x[:layer.weight.numel()].contiguous().view(*layer.weight.shape),
                      torch.unsqueeze(x[layer.weight.numel():], dim=1)], dim=1)


def efb_update(state: Optional[Tensor], diags: Optional[Tensor], layer: torch.nn.Module,
               eigvecs: Tuple[Tensor, Tensor], batch_size: int) -> Tuple[Tensor, Tensor]:
    """lambdas = (U_G^T g U_A)^2 accumulated; diags += g^2 * batch_size.  models/curvatures.py:436-446."""
    g = layer_grad_matrix(layer)
    lambdas = (eigvecs[1].t() @ g @ eigvecs[0]) ** 2
    d = g ** 2 * batch_size
    return (lambdas, d) if state is None else (state + lambdas, diags + d)


def efb_invert(state: Tensor, add: float, multiply: float) -> Tensor:
    """reciprocal(multiply * state + add).sqrt().  models/curvatures.py:462-463."""
    return torch.reciprocal(multiply * state + add).sqrt()


def efb_sample(eigvecs: Tuple[Tensor, Tensor], inv: Tensor, z: Tensor) -> Tensor:
    """(U_A (z * inv^T) U_G^T)^T with z [d_in', d_out].  models/curvatures.py:467-473."""
    first, second = eigvecs
    return (first @ (z * inv.t()) @ second.t()).t()


# =============================================================================== INF
def inf_dim_reduction(frst_eigvecs: Tensor, scnd_eigvecs: Tensor, lambda_vec: Tensor, rank: int
                      ) -> Tuple[Tensor, Tensor, Tensor]:
    """models/curvatures.py:615-658.  The `rank` largest |lambda| of the flat (d_in' x d_out) grid pick a
    set of grid rows (-> columns of U_A) and grid columns (-> columns of U_G); the low-rank lambda is the
    full rows x cols sub-grid, rows-major.  (The reference's list-of-0-dim-tensor indexing no longer runs;
    this follows its arithmetic: 1-based idx, i = int((idx - 1)/m + 1), j = idx - m (i - 1).)"""
    if rank >= lambda_vec.shape[0]:
        return frst_eigvecs, scnd_eigvecs, lambda_vec
    m = scnd_eigvecs.shape[1]
    order = np.argsort(-np.abs(lambda_vec.numpy()), kind="stable")[:rank]
    rows = np.unique(order // m)
    cols = np.unique(order % m)
    grid = (rows[:, None] * m + cols[None, :]).reshape(-1)
    return frst_eigvecs[:, rows], scnd_eigvecs[:, cols], lambda_vec[grid]


def inf_diagonal_accumulator(xxt_eigvecs: Tensor, ggt_eigvecs: Tensor, lambda_vec: Tensor) -> Tensor:
    """models/curvatures.py:660-682: diag_vec[i*m + p] = sum_{q,x} (U_A[i,q] U_G[p,x])^2 lambda[q*b + x]
    (the reference builds one Kronecker row block per i).  The reference accumulates into
    `torch.zeros(n * m)` (:675), a FLOAT32 buffer whatever the dtype of the factors, so its result is
    rounded to fp32; reproduced here."""
    a, b = xxt_eigvecs.shape[1], ggt_eigvecs.shape[1]
    lam = lambda_vec.view(a, b)
    return torch.einsum("iq,qx,px->ip", xxt_eigvecs ** 2, lam, ggt_eigvecs ** 2).reshape(-1).float()


def inf_update(eigvecs: Tuple[Tensor, Tensor], lambdas: Tensor, diags: Tensor, rank: int
               ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """INF.update, models/curvatures.py:500-520: (lr U_A, lr U_G, lr lambda, diag - sif_diag)."""
    xxt, ggt = eigvecs
    lambda_vec = lambdas.t().contiguous().view(-1)
    diag_vec = diags.t().contiguous().view(-1)
    lr_a, lr_g, lr_lambda = inf_dim_reduction(xxt, ggt, lambda_vec, rank)
    return lr_a, lr_g, lr_lambda, diag_vec - inf_diagonal_accumulator(lr_a, lr_g, lr_lambda)


def inf_pre_sampler(frst: Tensor, scnd: Tensor, reg_lambda: Tensor, reg_inv_correction: Tensor) -> Tensor:
    """INF.pre_sampler, models/curvatures.py:565-585, statement by statement (explicit Kronecker matrix,
    Cholesky factors, LU inverses)."""
    scale_sqrt = torch.diag(reg_lambda)
    v_s = reg_inv_correction.contiguous().view(-1, 1) * kron(frst, scnd) @ scale_sqrt
    vtv = v_s.t() @ v_s
    vtv = (vtv + vtv.t()) / 2.
    eye = torch.eye(scale_sqrt.shape[0], dtype=vtv.dtype)
    a_c_inv = torch.linalg.inv(torch.linalg.cholesky(vtv))
    b_c = torch.linalg.cholesky(vtv + eye)
    c = a_c_inv.t() @ (b_c - eye) @ a_c_inv
    l_c = torch.linalg.inv(torch.linalg.inv(c) + vtv)
    return scale_sqrt @ l_c @ scale_sqrt


def inf_invert(value: Tuple[Tensor, Tensor, Tensor, Tensor], add: float, multiply: float
               ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """INF.invert, models/curvatures.py:535-542 (returns a clamped copy; the reference clamps in place)."""
    lr_a, lr_g, lr_lambda, correction = value
    correction = correction.clamp(min=0)
    reg_lr_lambda = (multiply * lr_lambda).sqrt()
    reg_inv_correction = torch.reciprocal(multiply * correction + add).sqrt()
    return lr_a, lr_g, reg_inv_correction, inf_pre_sampler(lr_a, lr_g, reg_lr_lambda, reg_inv_correction)


def inf_sample(inv_value: Tuple[Tensor, Tensor, Tensor, Tensor], x: Tensor) -> Tensor:
    """INF.sampler + INF.sample, models/curvatures.py:544-548, 587-613, with the noise X given.
    Note the reference's two plain reshapes of flat vectors ((m, n) and (b, a)), kept as they are."""
    frst, scnd, c, pre = inv_value
    n, m = frst.shape[0], scnd.shape[0]
    y_l = c * x
    unvec_y_l = y_l.reshape(m, n)
    xq = scnd.t() @ unvec_y_l @ frst
    qx = pre @ xq.t().contiguous().view(-1)
    unvec_qx = qx.reshape(scnd.shape[1], frst.shape[1])
    x_p_s = scnd @ unvec_qx @ frst.t()
    y_r = c ** 2 * x_p_s.t().contiguous().view(-1)
    return (y_l - y_r).reshape(n, m).t()


# =============================================================================== calibration metrics (f3)
def metric_rows(probs: np.ndarray, labels: Optional[np.ndarray] = None) -> Dict[str, np.ndarray]:
    """Per-row quantities every metric of models/utilities.py:178-366 is built from."""
    p = np.asarray(probs)
    out = {"conf": p.max(axis=1), "pred": p.argmax(axis=1)}
    pn = p.astype(np.float64) / p.astype(np.float64).sum(axis=1, keepdims=True)
    with np.errstate(divide="ignore", invalid="ignore"):
        out["entropy"] = -np.where(pn > 0, pn * np.log(pn), 0.0).sum(axis=1)     # scipy.stats.entropy
    if labels is not None:
        out["correct"] = out["pred"] == labels
        out["nll"] = -np.log(p[np.arange(p.shape[0]), labels] + 1e-12)
    return out


def metric_accuracy(probs, labels) -> float:
    """models/utilities.py:178-189."""
    return 100.0 * float(np.mean(metric_rows(probs, labels)["correct"]))


def metric_nll(probs, labels) -> float:
    """models/utilities.py:236-247."""
    return float(np.mean(metric_rows(probs, labels)["nll"]))


def metric_ece(probs, labels, bins: int = 10):
    """models/utilities.py:300-332: equally spaced bins (lo, hi]; empty bins report zeros."""
    rows = metric_rows(probs, labels)
    conf, ok = rows["conf"], rows["correct"]
    edges = np.linspace(0, 1, bins + 1)
    which = np.full(conf.shape, -1)
    for i in range(bins):
        which[(conf > edges[i]) & (conf <= edges[i + 1])] = i
    ace, acc, cf = np.zeros(bins), np.zeros(bins), np.zeros(bins)
    ece = 0.0
    for i in range(bins):
        sel = which == i
        if sel.any():
            acc[i], cf[i] = ok[sel].mean(), conf[sel].mean()
            ace[i] = cf[i] - acc[i]
            ece += sel.mean() * abs(ace[i])
    return ece, ace, acc, cf


def metric_calibration_curve(probs, labels, bins: int = 20):
    """models/utilities.py:250-297: edges = every step-th sorted confidence (+ the maximum unless
    n % step == 1), intervals open on both sides, only non-empty bins reported."""
    rows = metric_rows(probs, labels)
    conf, ok = rows["conf"], rows["correct"]
    n = conf.shape[0]
    step = (n + bins - 1) // bins
    edges = np.sort(conf)[::step]
    if n % step != 1:
        edges = np.concatenate((edges, [conf.max()]))
    ece, xs, ys, zs = 0.0, [], [], []
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (conf > lo) & (conf < hi)
        if sel.mean() > 0:
            xs.append(conf[sel].mean())
            ys.append(ok[sel].mean())
            zs.append(sel.mean())
            ece += abs(xs[-1] - ys[-1]) * zs[-1]
    return ece, np.array(xs), np.array(ys), np.array(zs)


def metric_binned_kl(dist1, dist2, smooth: float = 1e-7, bins=None) -> float:
    """models/utilities.py:192-217: np.histogram both sample sets, smooth, normalise, KL both ways."""
    bins = np.logspace(-7, 1, num=200) if bins is None else bins
    p = np.histogram(dist1, bins)[0] + smooth
    q = np.histogram(dist2, bins)[0] + smooth
    p, q = p / p.sum(), q / q.sum()
    return float(np.sum(p * np.log(p / q)) + np.sum(q * np.log(q / p)))
