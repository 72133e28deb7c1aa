"""Kronecker-factored / diagonal Laplace curvature on B200 — drop-in for the reference's
`models/curvatures.py` (`Curvature`, `Diagonal`, `KFAC`).

Same constructor, methods and attribute layout as the reference (file:line citations are relative to
/root/reference):

    KFAC(model, layer_types=None) / Diagonal(model, layer_types=None)     curvatures.py:38-65, 295-317
    .update(batch_size)                                                   :155-172, :325-365
    .invert(add=0., multiply=1.)   scalars or per-layer lists             :190-202, :367-398
    .sample(layer) -> Tensor[d_out, d_in(+1)]                             :204-207, :400-405
    .sample_and_replace()                                                 :117-129
    .save(filename) / .load(filename)                                     :132-144
    .model .model_state .state .inv_state (.hooks .record for KFAC)

All arithmetic runs in libbk_kfac.so (hand-written sm_100a kernels, see include/bk_kfac.h).  There is
no CPU path: constructing an estimator for a model that is not on a CUDA device raises.

Differences that are deliberate and documented in DESIGN.md:
  * `KFAC.record[m][1]` holds the raw `grad_output[0]`; the reference's `* N` (curvatures.py:323) is
    folded into the factor kernel as an input scale (one HBM pass saved).
  * a factor that is not positive definite raises RuntimeError (the reference retries in NumPy on the
    CPU, curvatures.py:393-396; a CPU fallback is out of scope here by construction).
  * extra keyword-only knobs: `precision` ("bf16x3" default = parity mode, "bf16" = throughput mode),
    `seed` for the Philox generator; `sample(layer, z=...)` accepts an external noise tensor so that
    the reference and this engine can consume identical noise.
"""
from __future__ import annotations

import copy
import ctypes as C
from abc import ABC, abstractmethod
from typing import Any, Dict, List, Optional, Union

import torch
from torch import Tensor
from torch.nn import Module, Sequential

from . import _lib

_PRECISIONS = {"fp32": _lib.BK_PREC_FP32, "bf16": _lib.BK_PREC_BF16, "bf16x3": _lib.BK_PREC_BF16X3}


def gemm_precision(name: str) -> int:
    """Precision of the tensor-core contractions (sampling, forward, quadratic forms): "fp32" only
    changes the factor SYRK; everything downstream then runs split-bf16 (bf16x3)."""
    return _lib.BK_PREC_BF16 if name == "bf16" else _lib.BK_PREC_BF16X3


def _round8(v: int) -> int:
    return (v + 7) // 8 * 8


def _alloc_factor(d: int, device) -> Tensor:
    """fp32 [d, d] factor.  Wide factors get a row pitch that is a multiple of 4 elements (16 B) so
    that the accumulating TMA epilogue of the tensor-core SYRK can address them (d' = d_in + 1 is odd
    for every power-of-two layer width); the returned tensor is then a [d, d] view of a [d, pitch]
    buffer — same shape and values as the reference's factor, different stride(0)."""
    if d > _lib.BK_SMALL_D_MAX and d % 4:
        pitch = (d + 3) // 4 * 4
        return torch.empty(d, pitch, device=device, dtype=torch.float32)[:, :d]
    return torch.empty(d, d, device=device, dtype=torch.float32)


class _Workspace:
    """Grow-only device scratch buffer (256 B aligned) reused across kernel calls."""

    def __init__(self) -> None:
        self._buf: Optional[Tensor] = None

    def get(self, nbytes: int, device: torch.device) -> Tensor:
        nbytes = max(int(nbytes), 256)
        if self._buf is None or self._buf.numel() < nbytes or self._buf.device != device:
            self._buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._buf


class _FactorState(dict):
    """`KFAC.state`: the reference's dict (module -> [A, G], models/curvatures.py:363) whose values are brought
    up to date when they are READ.  The tensor-core factor update accumulates lower triangles only and, in the
    running-average mode, in units of a lazily applied scalar (see KFAC.update); any public read — `state[layer]`,
    `.get`, `.items()`, `.values()`, pickling — first runs bk_sym_finalize on the entries that need it, so callers
    always see the full symmetric factors of the reference.  Membership tests, `len`, iteration over keys and
    assignment (`state[layer] = [A, G]`, as the regression scripts' fixtures do) never touch the device."""

    def __init__(self, owner):
        super().__init__()
        self._owner = owner

    def __getitem__(self, key):
        self._owner._finalize_state(key)
        return dict.__getitem__(self, key)

    def get(self, key, default=None):
        return self[key] if key in self else default

    def items(self):
        self._owner._finalize_state()
        return dict.items(self)

    def values(self):
        self._owner._finalize_state()
        return dict.values(self)

    def pop(self, key, *default):
        if key in self:
            self._owner._finalize_state(key)
        return dict.pop(self, key, *default)

    def __setitem__(self, key, value):
        self._owner._state_assigned(key)
        dict.__setitem__(self, key, value)

    def raw(self, key):
        """The stored tensors as they are (possibly lower-only / unscaled): for the update path."""
        return dict.__getitem__(self, key)

    def __reduce__(self):   # pickles (save(), deepcopy) as a plain, finalised dict
        return (dict, (dict(self.items()),))


class Curvature(ABC):
    """Base class: holds the model, a deep copy of its MAP weights, `state` and `inv_state`.

    Mirrors reference `Curvature` (curvatures.py:17-144)."""

    def __init__(self,
                 model: Union[Module, Sequential],
                 layer_types: Union[List[str], str] = None,
                 *,
                 precision: str = "bf16x3",
                 seed: int = 0):
        self.model = model
        self.model_state = copy.deepcopy(model.state_dict())
        self.layer_types: List[str] = list()
        if isinstance(layer_types, str):
            self.layer_types.append(layer_types)
        elif isinstance(layer_types, list):
            if layer_types:
                self.layer_types.extend(layer_types)
            else:
                self.layer_types.extend(['Linear', 'Conv2d', 'MultiheadAttention'])
        elif layer_types is None:
            self.layer_types.extend(['Linear', 'Conv2d', 'MultiheadAttention'])
        else:
            raise TypeError
        for _type in self.layer_types:
            assert _type in ['Linear', 'Conv2d', 'MultiheadAttention']
        self.state: Dict[Any, Any] = dict()
        self.inv_state: Dict[Any, Any] = dict()

        if precision not in _PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(_PRECISIONS)}")
        self.precision = precision
        self.seed = int(seed)
        self._sample_counter = 0
        self._ws = _Workspace()
        self._lib = _lib.load()
        _lib.require_device()
        params = list(model.parameters())
        if params and not params[0].is_cuda:
            raise _lib.BkError("the model must live on a CUDA device (there is no CPU path)")

    # ------------------------------------------------------------------ helpers
    def _selected_layers(self):
        """(index, layer) of every Linear/Conv2d selected by layer_types, in model.modules() order —
        the order the reference iterates in and therefore its RNG consumption order."""
        idx = 0
        for layer in self.model.modules():
            name = layer.__class__.__name__
            if name in self.layer_types:
                if name in ['Linear', 'Conv2d']:
                    yield idx, layer
                    idx += 1
                elif name == 'MultiheadAttention':
                    raise NotImplementedError

    @staticmethod
    def _replace(sample: Tensor, weight: Tensor, bias: Tensor = None):
        """Adds `sample` ([out, in(+1)], last column = bias) to the parameters in place.
        Reference: curvatures.py:67-82."""
        if bias is not None:
            bias_sample = sample[:, -1].contiguous().view(*bias.shape)
            bias.data.add_(bias_sample)
            sample = sample[:, :-1]
        weight.data.add_(sample.contiguous().view(*weight.shape))

    @abstractmethod
    def update(self, *args: Any, **kwargs: Any):
        raise NotImplementedError

    @abstractmethod
    def invert(self, add: Union[float, list, tuple] = 0., multiply: Union[float, list, tuple] = 1.):
        raise NotImplementedError

    @abstractmethod
    def sample(self, layer: Module) -> Tensor:
        raise NotImplementedError

    def sample_and_replace(self):
        """Restores the MAP weights, then perturbs every selected layer by one posterior sample.
        Reference: curvatures.py:117-129."""
        self.model.load_state_dict(self.model_state)
        for layer in self.model.modules():
            if layer.__class__.__name__ in self.layer_types:
                if layer.__class__.__name__ in ['Linear', 'Conv2d']:
                    _sample = self.sample(layer)
                    self._replace(_sample, layer.weight, layer.bias)
                elif layer.__class__.__name__ == 'MultiheadAttention':
                    raise NotImplementedError

    def _names(self) -> Dict[Module, str]:
        return {m: n for n, m in self.model.named_modules()}

    def save(self, filename):
        """Reference format (curvatures.py:132-137) plus name-keyed copies so the file survives
        re-instantiating the model."""
        names = self._names()
        torch.save({
            'state': dict(self.state.items()),
            'inv_state': self.inv_state,
            'model': self.model,
            'state_by_name': {names[k]: v for k, v in self.state.items() if k in names},
            'inv_state_by_name': {names[k]: v for k, v in self.inv_state.items() if k in names},
        }, filename)
        print('Writting %s complete!\n' % filename)

    def load(self, filename):
        state_dict = torch.load(filename, weights_only=False)
        self.state = state_dict['state']
        self.inv_state = state_dict['inv_state']
        self.model = state_dict['model']
        self._invalidate_caches()
        print('Loading %s complete!\n' % filename)

    def load_by_name(self, filename):
        """Restores `state` / `inv_state` from the name-keyed copies of a file written by save() onto
        the CURRENT model's modules (same `named_modules()` names), keeping `self.model`: the checkpoint
        then survives re-instantiating the network, which the reference's module-object keys do not."""
        blob = torch.load(filename, weights_only=False)
        modules = dict(self.model.named_modules())
        dev = next(self.model.parameters()).device

        def to_dev(v):
            if isinstance(v, Tensor):
                return v.to(dev)
            if isinstance(v, (list, tuple)):
                return type(v)(to_dev(u) for u in v)
            return v

        for attr in ("state", "inv_state"):
            named = blob[attr + "_by_name"]
            missing = [k for k in named if k not in modules]
            if missing:
                raise KeyError(f"checkpoint layers {missing} do not exist in the current model")
            setattr(self, attr, {modules[k]: to_dev(v) for k, v in named.items()})
        self._invalidate_caches()
        print('Loading %s complete!\n' % filename)

    def _invalidate_caches(self):
        pass

    # The estimator is reachable from the pickled model through its hooks (bound methods), exactly as
    # in the reference; the ctypes handle and the scratch buffer are process-local and are dropped.
    def __getstate__(self):
        d = dict(self.__dict__)
        d.pop("_lib", None)
        d["_ws"] = None
        d["_staged"] = dict()
        if isinstance(self.state, _FactorState):
            d["state"] = dict(self.state.items())       # finalised, plain
            d["_dirty"] = set()
        return d

    def __setstate__(self, d):
        self.__dict__.update(d)
        self._lib = _lib.load()
        self._ws = _Workspace()
        if hasattr(self, "_dirty") and not isinstance(self.state, _FactorState):
            st = _FactorState(self)
            for k, v in self.state.items():
                dict.__setitem__(st, k, v)
            self.state = st

    def _next_sample_id(self) -> int:
        sid = self._sample_counter
        self._sample_counter += 1
        return sid


class Diagonal(Curvature):
    """Diagonal Fisher / GGN.  Reference: curvatures.py:146-207."""

    @_lib.nvtx_range("Diagonal.update")
    def update(self, batch_size: int):
        """state += [W.grad | b.grad]^2 * batch_size  (curvatures.py:155-172)."""
        st = _lib.stream_ptr()
        for _, layer in self._selected_layers():
            wg = layer.weight.grad
            if wg is None:
                raise RuntimeError("Diagonal.update needs .grad on every selected layer (call backward first)")
            d_out = wg.shape[0]
            wg2 = wg.contiguous().view(d_out, -1).float()
            d_in = wg2.shape[1]
            bg = None
            if layer.bias is not None:
                bg = layer.bias.grad.contiguous().float()
            if layer in self.state:
                state, beta = self.state[layer], 1.0
            else:
                state = torch.empty(d_out, d_in + (1 if bg is not None else 0), device=wg.device,
                                    dtype=torch.float32)
                beta = 0.0
                self.state[layer] = state
            _lib.check(self._lib.bk_diag_accum(state.data_ptr(), wg2.data_ptr(), _lib.ptr(bg), d_out,
                                               d_in, float(batch_size), beta, st), "bk_diag_accum")

    def invert(self, add: Union[float, list, tuple] = 0., multiply: Union[float, list, tuple] = 1.):
        """inv_state = 1/sqrt(multiply*state + add)  (curvatures.py:190-202)."""
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        st = _lib.stream_ptr()
        for index, (layer, value) in enumerate(self.state.items()):
            if isinstance(add, (list, tuple)) and isinstance(multiply, (list, tuple)):
                assert len(add) == len(multiply) == len(self.state)
                n, s = add[index], multiply[index]
            else:
                n, s = add, multiply
            inv = torch.empty_like(value)
            _lib.check(self._lib.bk_diag_invert(inv.data_ptr(), value.data_ptr(), value.numel(),
                                                float(n), float(s), st), "bk_diag_invert")
            self.inv_state[layer] = inv

    def sample(self, layer: Union[Module, str], z: Optional[Tensor] = None) -> Tensor:
        """N(0,1) * inv_state  (curvatures.py:204-207); `z` supplies external noise."""
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        inv = self.inv_state[layer]
        out = torch.empty_like(inv)
        if z is not None:
            z = z.to(inv.device, torch.float32).contiguous()
            assert z.shape == inv.shape
        stream_id = self._layer_stream_id(layer)
        _lib.check(self._lib.bk_diag_sample(out.data_ptr(), inv.data_ptr(), inv.numel(), 1, self.seed,
                                            self._next_sample_id(), stream_id, _lib.ptr(z),
                                            _lib.stream_ptr()), "bk_diag_sample")
        return out

    def _layer_stream_id(self, layer) -> int:
        for i, (l, _) in enumerate(self.inv_state.items()):
            if l is layer:
                return i
        return 0


_SIDE_STREAMS: dict = {}    # device -> four side streams shared by every estimator on it


class KFAC(Curvature):
    """Kronecker-factored Fisher.  Reference: curvatures.py:277-405."""

    def __init__(self,
                 model: Union[Module, Sequential],
                 layer_types: Union[List[str], str] = None,
                 *,
                 precision: str = "bf16x3",
                 seed: int = 0,
                 averaging: str = "sum",
                 decay: float = 0.95,
                 lower_only: bool = True):
        """averaging="sum" (default) is the reference: `state` is the plain sum of the per-batch factors
        (models/curvatures.py:359-363).  averaging="ema" keeps an exponential running average instead,
            state_1 = F_1,   state_t = decay * state_{t-1} + (1 - decay) * F_t,
        at no extra cost per update: the accumulators hold state / c with c = decay^(t-1), a new batch is
        added with weight (1 - decay) / c by the same `+=` kernels, and c is applied when `state` is read
        (bk_sym_finalize; also whenever c drops below 2^-20, i.e. every ~14 / (1 - decay) updates).
        lower_only: the tensor-core SYRK accumulates lower triangles only and `state` reads mirror them
        (halves the epilogue's reduction traffic per update); False restores the mirrored epilogue."""
        super().__init__(model, layer_types, precision=precision, seed=seed)
        if averaging not in ("sum", "ema"):
            raise ValueError('averaging must be "sum" or "ema"')
        if averaging == "ema" and not (0.0 < decay < 1.0):
            raise ValueError("decay must lie in (0, 1)")
        self.averaging, self.decay, self.lower_only = averaging, float(decay), bool(lower_only)
        self.state = _FactorState(self)
        self._dirty = set()     # layers whose wide factors currently hold a valid LOWER triangle only
        self._scale = 1.0       # state = _scale * stored accumulators (running-average mode)
        self._n_updates = 0
        self.hooks = list()
        self.record = dict()
        self._pending = []
        self._small_jobs = []
        self._staged = dict()  # layer -> bf16 K-major copies of (L_A, L_G) for the tensor-core GEMMs

        for layer in model.modules():
            if layer.__class__.__name__ in self.layer_types:
                if layer.__class__.__name__ in ['Linear', 'Conv2d']:
                    self.record[layer] = [None, None]
                    self.hooks.append(layer.register_forward_pre_hook(self._save_input))
                    self.hooks.append(layer.register_full_backward_hook(self._save_output))
                elif layer.__class__.__name__ == 'MultiheadAttention':
                    raise NotImplementedError

    def _save_input(self, module, input):
        self.record[module][0] = input[0]

    def _save_output(self, module, grad_input, grad_output):
        # reference stores grad_output[0] * N (curvatures.py:323); the factor N is applied inside the
        # factor kernel instead (in_scale), see update().
        self.record[module][1] = grad_output[0]

    # ------------------------------------------------------------------ lazy state
    def _raw(self, layer):
        st = self.state
        return st.raw(layer) if isinstance(st, _FactorState) else st[layer]

    def _raw_items(self):
        return dict.items(self.state)

    def _state_assigned(self, key):
        """A caller replaces `state[key]`: the new tensors are complete and unscaled."""
        if self._scale != 1.0:
            self._finalize_state()
        self._dirty.discard(key)

    def _finalize_state(self, key=None):
        """Mirror the lower-only accumulators (and apply the running-average scale) in place."""
        if self._scale == 1.0:
            todo = [k for k in ([key] if key is not None else list(self._dirty)) if k in self._dirty]
        else:
            todo = list(dict.keys(self.state))      # the scale is one scalar for the whole estimator
        if not todo:
            return
        tensors = [t for k in todo for t in self._raw(k)]
        n = len(tensors)
        mats = (C.c_void_p * n)(*[t.data_ptr() for t in tensors])
        lds = (C.c_longlong * n)(*[t.stride(0) for t in tensors])
        dims = (C.c_int * n)(*[t.shape[0] for t in tensors])
        _lib.check(self._lib.bk_sym_finalize(mats, lds, dims, n, float(self._scale), _lib.stream_ptr()),
                   "bk_sym_finalize")
        for k in todo:
            self._dirty.discard(k)
        self._scale = 1.0

    # ------------------------------------------------------------------ factor update
    def _syrk(self, state: Tensor, beta: float, x: Tensor, has_bias: bool, in_scale: float,
              alpha: float):
        """Queue state = beta*state + alpha*[in_scale*x ; 1]^T[...]; issued by _flush_syrks()."""
        self._pending.append((state, beta, x, has_bias, in_scale, alpha))

    def _direct_ok(self, state: Tensor, x: Tensor, has_bias: bool) -> bool:
        """Can this bf16 activation matrix feed the tensor cores as it is (no staging pass)?"""
        d = x.shape[1]
        return (x.dtype == torch.bfloat16 and self.precision != "fp32" and d >= 192
                and d + int(has_bias) > _lib.BK_SMALL_D_MAX and x.stride(1) == 1 and x.stride(0) % 8 == 0
                and x.data_ptr() % 16 == 0 and state.stride(0) % 4 == 0 and state.data_ptr() % 16 == 0)

    def _small_job(self, fn, *keep):
        """Queue one small independent factor kernel: fn(stream_ptr).  `keep` pins the tensors it reads until the
        join is enqueued (they may be temporaries of this update)."""
        self._small_jobs.append((fn, keep))

    def _flush_small_jobs(self):
        """The small factors of one update (conv layers, narrow Linear layers: 5 .. 176 wide) are independent
        kernels of a few CTAs each: round-robin over four side streams, forked from and joined into the current
        stream, so they run beside each other (and beside the tensor-core launch of the wide factors) instead
        of one after the other (LeNet-5: 10 factors, 21 launches, 283 us in line)."""
        jobs, self._small_jobs = self._small_jobs, []
        if not jobs:
            return
        main = torch.cuda.current_stream()
        if len(jobs) < 2 or not getattr(self, "multi_stream", True):
            for fn, _ in jobs:
                fn(main.cuda_stream)
            return
        dev = main.device
        side = _SIDE_STREAMS.get(dev)
        if side is None:
            side = _SIDE_STREAMS[dev] = [torch.cuda.Stream(device=dev) for _ in range(4)]
        fork = torch.cuda.Event()
        fork.record(main)
        used = min(len(side), len(jobs))
        for sst in side[:used]:
            sst.wait_event(fork)
        for i, (fn, _) in enumerate(jobs):
            fn(side[i % used].cuda_stream)
        for sst in side[:used]:
            ev = torch.cuda.Event()
            ev.record(sst)
            main.wait_event(ev)

    def _flush_syrks(self):
        """All queued factor updates of this update() in ONE grouped library call: the wide factors
        share persistent tensor-core launches (bk_syrk_accum_grouped).  bf16 activations (a model under
        bf16 autocast) are consumed directly — row-major X is the MN-major operand of X^T X — where the factor
        is wide enough; everything else is handed over as fp32."""
        items, self._pending = self._pending, []
        n = len(items)
        if n == 0:
            return
        fixed = []
        for (state, beta, x, has_bias, in_scale, alpha) in items:
            if x.dtype != torch.float32 and not self._direct_ok(state, x, has_bias):
                x = x.float()
            fixed.append((state, beta, x.contiguous() if x.stride(-1) != 1 else x, has_bias, in_scale, alpha))
        prec = _PRECISIONS[self.precision]
        # narrow factors: one small SIMT kernel each, handed to the side streams (see _flush_small_jobs)
        wide = []
        is_small = [it[2].dtype == torch.float32 and it[2].shape[1] + int(it[3]) <= _lib.BK_SMALL_D_MAX for it in fixed]
        # a single narrow factor stays in the grouped call (nothing to run it beside; it goes first there)
        divert = getattr(self, "multi_stream", True) and len(self._small_jobs) + sum(is_small) >= 2
        for it, small in zip(fixed, is_small):
            state, beta, x, has_bias, in_scale, alpha = it
            if small and divert:
                def job(stream, state=state, beta=beta, x=x, has_bias=has_bias, in_scale=in_scale, alpha=alpha):
                    _lib.check(self._lib.bk_syrk_accum(state.data_ptr(), state.stride(0), x.data_ptr(), x.stride(0),
                                                       x.shape[0], x.shape[1], int(has_bias), float(in_scale),
                                                       float(alpha), float(beta), prec, 0, 0, stream),
                               "bk_syrk_accum")
                self._small_job(job, state, x)
            else:
                wide.append(it)
        items = wide
        n = len(items)
        if n == 0:
            return
        ns = (C.c_int * n)(*[it[2].shape[0] for it in items])
        ds = (C.c_int * n)(*[it[2].shape[1] for it in items])
        hb = (C.c_int * n)(*[int(it[3]) for it in items])
        nbytes = self._lib.bk_syrk_grouped_workspace_bytes(ns, ds, hb, n, prec)
        ws = self._ws.get(nbytes, items[0][2].device)
        states = (C.c_void_p * n)(*[it[0].data_ptr() for it in items])
        lds = (C.c_longlong * n)(*[it[0].stride(0) for it in items])
        xs = (C.c_void_p * n)(*[it[2].data_ptr() for it in items])
        bf = (C.c_int * n)(*[int(it[2].dtype == torch.bfloat16) for it in items])
        ldx = (C.c_longlong * n)(*[it[2].stride(0) for it in items])
        insc = (C.c_float * n)(*[it[4] for it in items])
        alph = (C.c_float * n)(*[it[5] for it in items])
        beta = (C.c_float * n)(*[it[1] for it in items])
        flags = (_lib.SYRK_LOWER_ONLY if self.lower_only else 0) | int(getattr(self, "_syrk_flags_extra", 0))
        _lib.check(self._lib.bk_syrk_accum_grouped(states, lds, xs, bf, ldx, ns, ds, hb, insc, alph, beta, n, prec,
                                                   flags, ws.data_ptr(), nbytes, _lib.stream_ptr()),
                   "bk_syrk_accum_grouped")

    @_lib.nvtx_range("KFAC.update")
    def update(self, batch_size: int = None):
        """Accumulates A = [a;1][a;1]^T / cols and G = g g^T / cols per selected layer, with
        g = grad_output * N (curvatures.py:325-365).  `batch_size` is ignored, as in the reference."""
        del batch_size
        self._pending = []
        self._small_jobs = []
        # weight of this batch in units of the stored accumulators (see __init__): 1 for the plain sum
        w = 1.0
        if self.averaging == "ema" and self._n_updates > 0:
            if self._scale * self.decay < 2.0 ** -20:
                self._finalize_state()
            self._scale *= self.decay
            w = (1.0 - self.decay) / self._scale
        self._n_updates += 1
        wide_mode = self.lower_only and self.precision != "fp32"
        for _, layer in self._selected_layers():
            module_class = layer.__class__.__name__
            forward, backward = self.record[layer]
            if forward is None or backward is None:
                raise RuntimeError("KFAC.update needs a forward and a backward pass through the hooked model")
            forward = forward.detach()
            backward = backward.detach()
            n_batch = backward.shape[0]
            has_bias = layer.bias is not None
            if module_class == 'Conv2d':
                d_a = forward.shape[1] * layer.kernel_size[0] * layer.kernel_size[1] + int(has_bias)
                d_g = backward.shape[1]
            else:
                if forward.dim() != 2:
                    raise ValueError("Linear layers must see 2-D inputs [N, d_in] (as in the reference)")
                d_a = forward.shape[1] + int(has_bias)
                d_g = backward.shape[1]
            if layer in self.state:
                first, second = self._raw(layer)
                beta = 1.0
            else:
                first = _alloc_factor(d_a, forward.device)
                second = _alloc_factor(d_g, forward.device)
                self.state[layer] = [first, second]
                beta = 0.0
            if module_class == 'Conv2d':
                self._update_conv(layer, forward, backward, first, second, beta, has_bias, n_batch, w)
            else:
                # fp32 or bf16 as the model produced them (other dtypes are widened in _flush_syrks)
                x = forward if forward.stride(-1) == 1 else forward.contiguous()
                g = backward if backward.stride(-1) == 1 else backward.contiguous()
                self._syrk(first, beta, x, has_bias, 1.0, w / x.shape[0])
                self._syrk(second, beta, g, False, float(n_batch), w / g.shape[0])
            if wide_mode and max(d_a, d_g) > _lib.BK_SMALL_D_MAX:
                self._dirty.add(layer)      # the tensor-core SYRK leaves the upper triangle stale
        self._flush_syrks()
        self._flush_small_jobs()

    def _update_conv(self, layer, forward, backward, first, second, beta, has_bias, n_batch, w=1.0):
        st = _lib.stream_ptr()
        x = forward.float().contiguous()
        g = backward.float().contiguous()
        n, c, h, wd = x.shape
        kh, kw = layer.kernel_size
        ph, pw = layer.padding if not isinstance(layer.padding, str) else (0, 0)
        sh, sw = layer.stride
        if isinstance(layer.padding, str) or tuple(layer.dilation) != (1, 1) or layer.groups != 1:
            raise NotImplementedError("Conv2d with string padding, dilation or groups")
        oh = (h + 2 * ph - kh) // sh + 1
        ow = (wd + 2 * pw - kw) // sw + 1
        cols = n * oh * ow
        d_a = first.shape[0]
        if d_a <= _lib.BK_SMALL_D_MAX:
            def job_a(stream):
                _lib.check(self._lib.bk_conv_a_accum(first.data_ptr(), first.stride(0), x.data_ptr(), n, c,
                                                     h, wd, kh, kw, ph, pw, sh, sw, int(has_bias),
                                                     w / cols, beta, stream), "bk_conv_a_accum")
            self._small_job(job_a, first, x)
        elif self.precision == "fp32":
            # full-fp32 parity mode: the SIMT SYRK consumes an explicit fp32 patch matrix [N*L, C*kh*kw]
            u = torch.nn.functional.unfold(x, (kh, kw), padding=(ph, pw), stride=(sh, sw))
            u = u.permute(0, 2, 1).reshape(cols, -1).contiguous()
            self._syrk(first, beta, u, has_bias, 1.0, w / cols)
        else:
            # wide conv layers: the K-major bf16 patch operand straight from NCHW (bk_im2col_split: no fp32
            # unfold / permute / contiguous round trips), then the tensor-core SYRK
            self._staged_conv_syrk(first, beta, x, (n, c, h, wd, kh, kw, ph, pw, sh, sw), cols, has_bias, 1.0,
                                   w / cols)
        o = g.shape[1]
        hw = g.shape[2] * g.shape[3]
        gcols = n * hw
        if o <= _lib.BK_SMALL_D_MAX:
            def job_g(stream):
                _lib.check(self._lib.bk_conv_g_accum(second.data_ptr(), second.stride(0), g.data_ptr(), n, o,
                                                     hw, float(n_batch), w / gcols, beta, stream),
                           "bk_conv_g_accum")
            self._small_job(job_g, second, g)
        elif self.precision == "fp32":
            g2 = g.permute(0, 2, 3, 1).reshape(gcols, o).contiguous()
            self._syrk(second, beta, g2, False, float(n_batch), w / gcols)
        else:
            # [N, O, H'W'] -> [O, N*H'W'] is the 1 x 1 case of the same staging kernel
            self._staged_conv_syrk(second, beta, g, (n, o, g.shape[2], g.shape[3], 1, 1, 0, 0, 1, 1), gcols, False,
                                   float(n_batch), w / gcols)

    def _staged_conv_syrk(self, state: Tensor, beta: float, x: Tensor, shape, cols: int, has_bias: bool,
                          in_scale: float, alpha: float):
        """state = beta*state + alpha * U U^T with U = [in_scale * unfold(x) ; 1] staged K-major in bf16."""
        n, c, h, wd, kh, kw, ph, pw, sh, sw = shape
        lib, st = self._lib, _lib.stream_ptr()
        x3 = self.precision == "bf16x3"
        rows = c * kh * kw + int(has_bias)
        ldt = _round8(cols)
        nbytes = rows * ldt * 2
        buf = self._ws.get((2 if x3 else 1) * nbytes + 256, x.device)
        t_hi = buf.data_ptr()
        t_lo = t_hi + (nbytes + 255) // 256 * 256 if x3 else 0
        _lib.check(lib.bk_im2col_split(x.data_ptr(), n, c, h, wd, kh, kw, ph, pw, sh, sw, float(in_scale),
                                       int(has_bias), t_hi, t_lo, ldt, st), "bk_im2col_split")
        _lib.check(lib.bk_syrk_accum_staged(state.data_ptr(), state.stride(0), t_hi, t_lo, ldt, cols, rows,
                                            float(alpha), float(beta), _PRECISIONS[self.precision], st),
                   "bk_syrk_accum_staged")

    # ------------------------------------------------------------------ inversion
    @_lib.nvtx_range("KFAC.invert")
    def invert(self, add: Union[float, list, tuple] = 0., multiply: Union[float, list, tuple] = 1.):
        """inv_state[layer] = (chol(inv(R_A)), chol(inv(R_G))) with R = sqrt(s) F + sqrt(n) I,
        symmetrised (curvatures.py:367-398).  One batched launch sequence for all factors."""
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        factors, adds, mults, layers = [], [], [], []
        for index, (layer, value) in enumerate(self.state.items()):
            if not isinstance(add, (float, int)) and not isinstance(multiply, (float, int)):
                assert len(add) == len(multiply) == len(self.state)
                n, s = add[index], multiply[index]
            else:
                n, s = float(add), float(multiply)
            first, second = value
            factors += [first, second]
            adds += [float(n), float(n)]
            mults += [float(s), float(s)]
            layers.append(layer)
        outs = invert_factors(factors, adds, mults, self._ws)
        for i, layer in enumerate(layers):
            self.inv_state[layer] = (outs[2 * i], outs[2 * i + 1])
        self._invalidate_caches()

    def _invalidate_caches(self):
        self._staged = dict()

    def _adopt_loaded_state(self):
        """`load` / `load_by_name` replaced `state` by a plain dict of complete factors."""
        st = _FactorState(self)
        for k, v in self.state.items():
            dict.__setitem__(st, k, v)
        self.state = st
        self._dirty = set()
        self._scale = 1.0

    def load(self, filename):
        super().load(filename)
        self._adopt_loaded_state()

    def load_by_name(self, filename):
        super().load_by_name(filename)
        self._adopt_loaded_state()

    # ------------------------------------------------------------------ sampling
    def _staged_factors(self, layer):
        """bf16 (hi, lo) copies of L_A [d_in', d_in'] and L_G [d_out, d_out], padded to ld % 8 == 0."""
        if layer not in self._staged:
            self._staged[layer] = tuple(stage_operand(L, lower_only=True) for L in self.inv_state[layer])
        return self._staged[layer]

    def _implicit_operands(self, layer):
        """Staged bf16 operands of the implicit MC forward (predictive._implicit_linear), cached per
        layer until the next invert(): L_A^T (upper triangular) and the K-concatenated
        [M~ | 0 | L_G] with M~ = [W | b] in columns [0, d_in') and L_G (lower) from column kx on."""
        key = ("implicit", layer)
        if key not in self._staged:
            lib = self._lib
            LA, LG = self.inv_state[layer]
            d = LA.shape[0]
            kx = _round8(d)
            t_hi = torch.zeros(d, kx, dtype=torch.bfloat16, device=LA.device)
            t_lo = torch.zeros_like(t_hi)
            la = LA.float().contiguous()
            _lib.check(lib.bk_transpose_split(la.data_ptr(), la.stride(0), d, d, 1.0, 0, t_hi.data_ptr(),
                                              t_lo.data_ptr(), kx, _lib.stream_ptr()), "bk_transpose_split")
            d_out = LG.shape[0]
            m = layer.weight.detach().float().reshape(d_out, -1)
            if layer.bias is not None:
                m = torch.cat([m, layer.bias.detach().float().reshape(d_out, 1)], dim=1)
            m = m.contiguous()
            lg = LG.float().contiguous()
            ldcat = kx + _round8(d_out)
            c_hi = torch.zeros(d_out, ldcat, dtype=torch.bfloat16, device=LA.device)
            c_lo = torch.zeros_like(c_hi)
            st = _lib.stream_ptr()
            _lib.check(lib.bk_convert_split(m.data_ptr(), m.stride(0), d_out, m.shape[1], 1.0, 0,
                                            c_hi.data_ptr(), c_lo.data_ptr(), ldcat, st), "bk_convert_split")
            _lib.check(lib.bk_convert_split(lg.data_ptr(), lg.stride(0), d_out, d_out, 1.0, 1,
                                            c_hi.data_ptr() + 2 * kx, c_lo.data_ptr() + 2 * kx, ldcat, st),
                       "bk_convert_split")
            self._staged[key] = {"LAT": (t_hi, t_lo, kx), "MLG": (c_hi, c_lo, ldcat), "kx": kx,
                                 "la_dd": LA[d - 1, d - 1].reshape(1)}
        return self._staged[key]

    def sample(self, layer: Module, z: Optional[Tensor] = None) -> Tensor:
        """(L_A z L_G^T)^T -> [d_out, d_in'] with z ~ N(0,1)^{d_in' x d_out} (curvatures.py:400-405).
        `z` (shape [d_in', d_out]) replaces the internal Philox draw."""
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        return self.sample_batch(layer, 1, z=None if z is None else z.unsqueeze(0))[0]

    @_lib.nvtx_range("KFAC.sample_batch")
    def sample_batch(self, layer: Module, n_samples: int, z: Optional[Tensor] = None,
                     sample0: Optional[int] = None) -> Tensor:
        """`n_samples` posterior samples of one layer in one batched launch: [S, d_out, d_in'].
        Sample s uses Philox subsequence `sample0 + s` (default: the estimator's running counter), so
        a sample's value does not depend on how samples are sharded over GPUs."""
        from .sampling import matrix_normal_samples
        first, second = self.inv_state[layer]
        stA, stG = self._staged_factors(layer)
        if sample0 is None:
            sample0 = self._sample_counter
            self._sample_counter += n_samples
        layer_id = [l for _, l in self._selected_layers()].index(layer)
        return matrix_normal_samples(stA, stG, first.shape[0], second.shape[0], n_samples,
                                     precision=gemm_precision(self.precision), seed=self.seed,
                                     sample0=sample0, stream_id=layer_id, z=z, ws=self._ws)


# ---------------------------------------------------------------------------------------------------
def stage_operand(x: Tensor, lower_only: bool = False, scale: float = 1.0):
    """fp32 [rows, cols] -> (hi, lo, ld): bf16 split copies with row pitch ld (multiple of 8)."""
    lib = _lib.load()
    rows, cols = x.shape
    ld = _round8(cols)
    hi = torch.zeros(rows, ld, dtype=torch.bfloat16, device=x.device)
    lo = torch.zeros(rows, ld, dtype=torch.bfloat16, device=x.device)
    xc = x.float().contiguous()
    _lib.check(lib.bk_convert_split(xc.data_ptr(), xc.stride(0), rows, cols, scale, int(lower_only),
                                    hi.data_ptr(), lo.data_ptr(), ld, _lib.stream_ptr()),
               "bk_convert_split")
    return hi, lo, ld


def invert_factors(factors: List[Tensor], adds: List[float], mults: List[float],
                   ws: Optional[_Workspace] = None) -> List[Tensor]:
    """Batched damped chol(inv(.)) of square fp32 factors (bk_damp_chol_inv_batched).
    Raises RuntimeError naming the first factor that is not positive definite."""
    lib = _lib.load()
    n = len(factors)
    if n == 0:
        return []
    fs = [f.float().contiguous() for f in factors]
    outs = [torch.empty_like(f) for f in fs]
    dims = (C.c_int * n)(*[f.shape[0] for f in fs])
    fptr = (C.c_void_p * n)(*[f.data_ptr() for f in fs])
    optr = (C.c_void_p * n)(*[o.data_ptr() for o in outs])
    av = (C.c_float * n)(*adds)
    mv = (C.c_float * n)(*mults)
    nbytes = lib.bk_chol_inv_workspace_bytes(dims, n)
    ws = ws or _Workspace()
    buf = ws.get(nbytes, fs[0].device)
    rc = lib.bk_damp_chol_inv_batched(fptr, optr, dims, av, mv, n, buf.data_ptr(), nbytes,
                                      _lib.stream_ptr())
    _lib.check(rc, "bk_damp_chol_inv_batched")
    if rc > 0:
        raise RuntimeError(f"factor {rc - 1} (d={fs[rc - 1].shape[0]}) is not positive definite after "
                           f"damping (add={adds[rc - 1]}, multiply={mults[rc - 1]})")
    return outs


# ---------------------------------------------------------------------------------------------------
def _layer_grads(layer: Module) -> Tensor:
    """[d_out, d_in(+1)]: weight.grad viewed [d_out, -1] with the bias gradient as last column."""
    g = layer.weight.grad.contiguous().view(layer.weight.grad.shape[0], -1).float()
    if layer.bias is not None:
        g = torch.cat([g, layer.bias.grad.float().unsqueeze(1)], dim=1)
    return g.contiguous()


class BlockDiagonal(Curvature):
    """Per-layer full Fisher.  Reference: models/curvatures.py:210-275.

        update   state[layer] += ger(g, g) * batch_size,  g = [W.grad.view(-1), b.grad]      (:224-232)
        invert   inv_state[layer] = pinverse(multiply * state + add * I)                      (:258-268)
        sample   x = z @ inv_state[layer], reshaped to [d_out, d_in(+1)]                       (:270-275)

    `multiply*state + add*I` is symmetric positive definite for add > 0, so the pseudo-inverse is the
    inverse and comes from the batched Cholesky kernels (L L^T).  For add == 0 (rank-deficient sums of
    outer products) the pseudo-inverse is built from the Jacobi eigendecomposition; singular values
    below d * eps32 * max are dropped (torch.pinverse's rcond = 1e-15 presumes fp64 data)."""

    def update(self, batch_size: int):
        st = _lib.stream_ptr()
        for _, layer in self._selected_layers():
            g = layer.weight.grad.contiguous().view(-1).float()
            if layer.bias is not None:
                g = torch.cat([g, layer.bias.grad.float()])
            g = g.contiguous()
            P = g.numel()
            if layer in self.state:
                state, beta = self.state[layer], 1.0
            else:
                state, beta = _alloc_factor(P, g.device), 0.0
                self.state[layer] = state
            _lib.check(self._lib.bk_ger_accum(state.data_ptr(), state.stride(0), g.data_ptr(), P,
                                              float(batch_size), beta, st), "bk_ger_accum")

    def invert(self, add: Union[float, list, tuple] = 0., multiply: Union[float, list, tuple] = 1.):
        from .predictive import inverse_from_chol
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        for index, (layer, value) in enumerate(self.state.items()):
            if not isinstance(add, float) and not isinstance(multiply, float):
                assert len(add) == len(multiply) == len(self.state)
                n, s = add[index], multiply[index]
            else:
                n, s = add, multiply
            if n > 0:
                (Lc,) = invert_factors([value], [float(n) ** 2], [float(s) ** 2], self._ws)
                self.inv_state[layer] = inverse_from_chol(Lc)
            else:
                self.inv_state[layer] = _pinv_eigh(value, float(s))
        self._staged = dict()

    def _staged_inverse(self, layer):
        if not hasattr(self, "_staged"):
            self._staged = dict()
        if layer not in self._staged:
            self._staged[layer] = stage_operand(self.inv_state[layer])
        return self._staged[layer]

    def sample_batch(self, layer: Module, n_samples: int, z: Optional[Tensor] = None,
                     sample0: Optional[int] = None) -> Tensor:
        """[S, d_out, d_in(+1)] samples x_s = z_s @ inv (one batched tensor-core GEMM; inv is symmetric)."""
        inv = self.inv_state[layer]
        P = inv.shape[0]
        dev = inv.device
        lib = self._lib
        if sample0 is None:
            sample0 = self._sample_counter
            self._sample_counter += n_samples
        if z is None:
            layer_id = [l for _, l in self._selected_layers()].index(layer)
            zf = torch.empty(n_samples, P, device=dev, dtype=torch.float32)
            # sample s = row 0 of the [1, P] matrix of Philox subsequence sample0 + s (shard-invariant)
            _lib.check(lib.bk_philox_normal(self.seed, sample0, layer_id, 1, P, n_samples, zf.data_ptr(), P, P,
                                            0, 0, 0, 0, _lib.stream_ptr()), "bk_philox_normal")
        else:
            zf = z.to(dev, torch.float32).reshape(n_samples, P).contiguous()
        z_hi, z_lo, ldz = stage_operand(zf)
        i_hi, i_lo, ldi = self._staged_inverse(layer)
        prec = gemm_precision(self.precision)
        x3 = prec == _lib.BK_PREC_BF16X3
        out = torch.empty(n_samples, P, device=dev, dtype=torch.float32)
        _lib.check(lib.bk_gemm_nt(z_hi.data_ptr(), z_lo.data_ptr() if x3 else 0, ldz, 0,
                                  i_hi.data_ptr(), i_lo.data_ptr() if x3 else 0, ldi, 0,
                                  n_samples, P, P, 1, prec, 0, 1.0, 0.0, out.data_ptr(), P, 0, 0, 0, 0, 0, 0, 0,
                                  _lib.stream_ptr()), "bk_gemm_nt(z inv)")
        d_out = layer.weight.shape[0]
        nw = layer.weight.numel()
        w = out[:, :nw].reshape(n_samples, d_out, -1)
        if layer.bias is None:
            return w
        return torch.cat([w, out[:, nw:].unsqueeze(2)], dim=2)

    def sample(self, layer: Module, z: Optional[Tensor] = None) -> Tensor:
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        return self.sample_batch(layer, 1, z=None if z is None else z.reshape(1, -1))[0]


def _pinv_eigh(value: Tensor, s: float) -> Tensor:
    """pinverse(s * value) for a symmetric PSD matrix through the Jacobi eigensolver:
    V diag(1 / w_i for w_i > 1e-15 * w_max) V^T (torch.pinverse's default rcond)."""
    from .utilities import eigh_factors
    lib = _lib.load()
    (w,), (v,) = eigh_factors([value], sym_scale=0.5 * s)
    # torch.pinverse's rcond is 1e-15, meaningful for the reference's fp64 runs; in fp32 arithmetic
    # eigenvalues below d * eps32 * w_max are rounding noise of the (rank-deficient) sum of outer
    # products and are dropped
    cut = max(1e-15, value.shape[0] * 1.2e-7) * w.abs().max()
    winv = torch.where(w.abs() > cut, 1.0 / w, torch.zeros_like(w))
    # (V * winv) V^T as one contraction: A = V * winv (columns scaled), B = V
    a = (v * winv.unsqueeze(0)).contiguous()
    a_hi, a_lo, lda = stage_operand(a)
    b_hi, b_lo, ldb = stage_operand(v)
    d = v.shape[0]
    out = torch.empty(d, d, device=v.device, dtype=torch.float32)
    _lib.check(lib.bk_gemm_nt(a_hi.data_ptr(), a_lo.data_ptr(), lda, 0, b_hi.data_ptr(), b_lo.data_ptr(), ldb, 0,
                              d, d, d, 1, _lib.BK_PREC_BF16X3, 0, 1.0, 0.0, out.data_ptr(), d, 0, 0, 0, 0, 0, 0, 0,
                              _lib.stream_ptr()), "bk_gemm_nt(V w^-1 V^T)")
    return out


class EFB(Curvature):
    """Eigenvalue-corrected Kronecker-factored Fisher.  Reference: models/curvatures.py:408-473 (which
    no longer runs: get_eigenvectors calls the removed torch.symeig).

        __init__  eigvecs[layer] = (U_A, U_G) = eigenvectors of (A + A^T, G + G^T)             (:423-424)
        update    state[layer] += (U_G^T g U_A)^2 ;  diags[layer] += g^2 * batch_size          (:438-446)
        invert    inv_state[layer] = 1 / sqrt(multiply * state + add)                           (:462-463)
        sample    z ~ N(0,1)[d_in', d_out];  z *= inv_state^T;  (U_A z U_G^T)^T -> [d_out, d_in'] (:467-473)
    """

    def __init__(self, model, factors: Dict[Module, Any], layer_types=None, *, precision: str = "bf16x3",
                 seed: int = 0):
        from .utilities import get_eigenvectors
        super().__init__(model, layer_types, precision=precision, seed=seed)
        self.eigvecs = get_eigenvectors(factors)
        self.diags: Dict[Any, Tensor] = dict()
        self._staged = dict()

    def _staged_eigvecs(self, layer):
        """bf16 operands: U_A [d_in', d_in'] and U_G^T [d_out, d_out] (row-major, K contiguous)."""
        if layer not in self._staged:
            ua, ug = self.eigvecs[layer]
            self._staged[layer] = {"UA": stage_operand(ua), "UAT": stage_operand(ua.t().contiguous()),
                                   "UG": stage_operand(ug), "UGT": stage_operand(ug.t().contiguous())}
        return self._staged[layer]

    def _sandwich(self, left, mid: Tensor, right, rows: int, cols: int) -> Tensor:
        """left[rows, rows] @ mid[rows, cols] @ right[cols, cols]^T with staged bf16x3 operands:
        P^T = right mid^T (K = cols), out = left P (K = rows)."""
        lib = self._lib
        dev = mid.device
        m_hi, m_lo, ldm = stage_operand(mid)                    # [rows, cols]
        ldp = _round8(rows)
        p_hi = torch.zeros(cols, ldp, dtype=torch.bfloat16, device=dev)
        p_lo = torch.zeros_like(p_hi)
        r_hi, r_lo, ldr = right
        # P^T[j, o] = sum_i right[j, i] mid[o, i]
        _lib.check(lib.bk_gemm_nt(r_hi.data_ptr(), r_lo.data_ptr(), ldr, 0, m_hi.data_ptr(), m_lo.data_ptr(), ldm, 0,
                                  cols, rows, cols, 1, _lib.BK_PREC_BF16X3, 0, 1.0, 0.0, 0, 0, 0, 0, 0,
                                  p_hi.data_ptr(), p_lo.data_ptr(), ldp, 0, _lib.stream_ptr()), "bk_gemm_nt(R M^T)")
        l_hi, l_lo, ldl = left
        out = torch.empty(rows, cols, device=dev, dtype=torch.float32)
        # out[p, j] = sum_o left[p, o] P^T[j, o]
        _lib.check(lib.bk_gemm_nt(l_hi.data_ptr(), l_lo.data_ptr(), ldl, 0, p_hi.data_ptr(), p_lo.data_ptr(), ldp, 0,
                                  rows, cols, rows, 1, _lib.BK_PREC_BF16X3, 0, 1.0, 0.0, out.data_ptr(), cols, 0,
                                  0, 0, 0, 0, 0, 0, _lib.stream_ptr()), "bk_gemm_nt(L P)")
        return out

    def update(self, batch_size: int):
        st = _lib.stream_ptr()
        for _, layer in self._selected_layers():
            g = _layer_grads(layer)                               # [d_out, d_in']
            d_out, d_inp = g.shape
            ops = self._staged_eigvecs(layer)
            # U_G^T g U_A = UGT @ g @ (UAT)^T
            proj = self._sandwich(ops["UGT"], g, ops["UAT"], d_out, d_inp)
            if layer in self.state:
                state, diag, beta = self.state[layer], self.diags[layer], 1.0
            else:
                state = torch.empty(d_out, d_inp, device=g.device, dtype=torch.float32)
                diag = torch.empty_like(state)
                self.state[layer], self.diags[layer], beta = state, diag, 0.0
            _lib.check(self._lib.bk_diag_accum(state.data_ptr(), proj.data_ptr(), 0, d_out, d_inp, 1.0, beta, st),
                       "bk_diag_accum(lambdas)")
            _lib.check(self._lib.bk_diag_accum(diag.data_ptr(), g.data_ptr(), 0, d_out, d_inp, float(batch_size),
                                               beta, st), "bk_diag_accum(diags)")

    def invert(self, add: Union[float, list, tuple] = 0., multiply: Union[float, list, tuple] = 1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        st = _lib.stream_ptr()
        for index, (layer, value) in enumerate(self.state.items()):
            if not isinstance(add, float) and not isinstance(multiply, float):
                assert len(add) == len(multiply) == len(self.state)
                n, s = add[index], multiply[index]
            else:
                n, s = add, multiply
            inv = torch.empty_like(value)
            _lib.check(self._lib.bk_diag_invert(inv.data_ptr(), value.data_ptr(), value.numel(), float(n), float(s),
                                                st), "bk_diag_invert")
            self.inv_state[layer] = inv

    def sample(self, layer: Module, z: Optional[Tensor] = None) -> Tensor:
        """(U_A (z * inv_state^T) U_G^T)^T = U_G (z^T * inv_state) U_A^T -> [d_out, d_in'].
        `z` ([d_in', d_out], the reference's orientation) replaces the Philox draw."""
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        inv = self.inv_state[layer]
        d_out, d_inp = inv.shape
        zt = None
        if z is not None:
            zt = z.to(inv.device, torch.float32).t().contiguous()
            assert zt.shape == inv.shape
        scaled = torch.empty_like(inv)
        layer_id = [l for _, l in self._selected_layers()].index(layer)
        _lib.check(self._lib.bk_diag_sample(scaled.data_ptr(), inv.data_ptr(), inv.numel(), 1, self.seed,
                                            self._next_sample_id(), layer_id, _lib.ptr(zt), _lib.stream_ptr()),
                   "bk_diag_sample")
        ops = self._staged_eigvecs(layer)
        return self._sandwich(ops["UG"], scaled, ops["UA"], d_out, d_inp)


# ---------------------------------------------------------------------------------------------------
def stage_operand_t(x: Tensor):
    """fp32 [rows, cols] -> staged bf16 (hi, lo, ld) of x^T [cols, rows] (bk_transpose_split)."""
    lib = _lib.load()
    rows, cols = x.shape
    ld = _round8(rows)
    hi = torch.zeros(cols, ld, dtype=torch.bfloat16, device=x.device)
    lo = torch.zeros_like(hi)
    xc = x.float()
    if xc.stride(1) != 1:
        xc = xc.contiguous()
    _lib.check(lib.bk_transpose_split(xc.data_ptr(), xc.stride(0), rows, cols, 1.0, 0, hi.data_ptr(),
                                      lo.data_ptr(), ld, _lib.stream_ptr()), "bk_transpose_split")
    return hi, lo, ld


def mm_nt_staged(a, b, m: int, n: int, k: int, alpha: float = 1.0, beta: float = 0.0,
                 out: Optional[Tensor] = None) -> Tensor:
    """out[m, n] = alpha * A B^T + beta * out for staged K-major operands A [m, k], B [n, k]
    (split-bf16, three tensor-core passes: fp32-class accuracy)."""
    lib = _lib.load()
    a_hi, a_lo, lda = a
    b_hi, b_lo, ldb = b
    if out is None:
        assert beta == 0.0
        out = torch.empty(m, n, device=a_hi.device, dtype=torch.float32)
    _lib.check(lib.bk_gemm_nt(a_hi.data_ptr(), a_lo.data_ptr(), lda, 0, b_hi.data_ptr(), b_lo.data_ptr(), ldb, 0,
                              m, n, k, 1, _lib.BK_PREC_BF16X3, 0, alpha, beta, out.data_ptr(), out.stride(0), 0,
                              0, 0, 0, 0, 0, 0, _lib.stream_ptr()), "bk_gemm_nt")
    return out


def mm_nt(x: Tensor, y: Tensor, x_t: bool = False, y_t: bool = False, alpha: float = 1.0, beta: float = 0.0,
          out: Optional[Tensor] = None) -> Tensor:
    """op(x) @ op(y)^T on the tensor cores; op transposes first when the flag is set
    (x [K, M] / y [K, N] are then staged transposed)."""
    a = stage_operand_t(x) if x_t else stage_operand(x)
    b = stage_operand_t(y) if y_t else stage_operand(y)
    m, k = (x.shape[1], x.shape[0]) if x_t else x.shape
    n, k2 = (y.shape[1], y.shape[0]) if y_t else y.shape
    assert k == k2
    return mm_nt_staged(a, b, m, n, k, alpha, beta, out)


def _square(x: Tensor) -> Tensor:
    """Elementwise x^2 (the squaring form of the diagonal accumulate kernel)."""
    lib = _lib.load()
    xc = x.float().contiguous()
    out = torch.empty_like(xc)
    _lib.check(lib.bk_diag_accum(out.data_ptr(), xc.data_ptr(), 0, xc.shape[0], xc.shape[1], 1.0, 0.0,
                                 _lib.stream_ptr()), "bk_diag_accum(square)")
    return out


class INF(Curvature):
    """Information-form low-rank approximation with diagonal correction.  Reference:
    models/curvatures.py:476-682 (paths relative to /root/reference).

        __init__  eigvecs = get_eigenvectors(factors); lambdas (EFB.state), diags (Diagonal.state)   (:494-498)
        update    per layer: keep the eigen-directions of the `rank` largest |lambda| (index sets
                  idx_left x idx_right, _dim_reduction :615-658), correction = diag - sif_diag with
                  sif_diag[i*m + p] = sum_{q,x} U_A[i,q]^2 U_G[p,x]^2 lambda[q,x] (_diagonal_accumulator
                  :660-682, there an O(n) Python loop over Kronecker rows)                              (:500-520)
        invert    clamp the correction at 0, regularise, pre_sampler -> P_c                           (:522-542)
        sample    sampler :587-613, reshaped [n, m] and transposed -> [d_out, d_in']                  (:544-548)

    The dense products run on the tensor cores (split-bf16), the r x r pre-sampler chain in fp64
    (bk_inf_presample, see bk_inf.cu for why), element-wise stages in bk_inf_regularise / bk_inf_combine.
    Index selection (top-k, unique, gathers) is tensor bookkeeping in torch."""

    def __init__(self, model, diags: Dict[Module, Tensor], factors: Dict[Module, Any],
                 lambdas: Dict[Module, Tensor], layer_types=None, *, precision: str = "bf16x3", seed: int = 0):
        from .utilities import get_eigenvectors
        super().__init__(model, layer_types, precision=precision, seed=seed)
        assert diags.keys() == factors.keys() == lambdas.keys()
        self.eigvecs = get_eigenvectors(factors)
        self.lambdas = lambdas
        self.diags = diags
        self._staged = dict()

    # --------------------------------------------------------------------------- update
    @staticmethod
    def _dim_reduction(frst_eigvecs: Tensor, scnd_eigvecs: Tensor, lambda_vec: Tensor, rank: int):
        """Rows / columns of the (d_in' x d_out) lambda grid touched by the `rank` largest |lambda|;
        returns (U_A[:, rows], U_G[:, cols], lambda[rows x cols] flattened row-major)."""
        if rank >= lambda_vec.shape[0]:
            return frst_eigvecs, scnd_eigvecs, lambda_vec
        m = scnd_eigvecs.shape[1]
        top = torch.topk(lambda_vec.abs(), rank).indices
        idx_left = torch.unique(torch.div(top, m, rounding_mode="floor"))
        idx_right = torch.unique(top - torch.div(top, m, rounding_mode="floor") * m)
        grid = (idx_left.unsqueeze(1) * m + idx_right.unsqueeze(0)).reshape(-1)
        return (frst_eigvecs.index_select(1, idx_left).contiguous(),
                scnd_eigvecs.index_select(1, idx_right).contiguous(),
                lambda_vec.index_select(0, grid).contiguous())

    @staticmethod
    def _corrected_diagonal(xxt_eigvecs: Tensor, ggt_eigvecs: Tensor, lambda_vec: Tensor, diag_vec: Tensor):
        """diag_vec - sif_diag with sif_diag = (U_A^2 Lambda U_G^2^T) flattened [n, m] row-major: two
        contractions, the subtraction in the epilogue of the second (alpha = -1, beta = 1)."""
        n, a = xxt_eigvecs.shape
        m, b = ggt_eigvecs.shape
        lam_t = lambda_vec.view(a, b).t()                              # [b, a]
        t = mm_nt(_square(xxt_eigvecs), lam_t.contiguous())            # [n, b] = U_A^2 Lambda
        out = diag_vec.clone().view(n, m)
        mm_nt(t, _square(ggt_eigvecs), alpha=-1.0, beta=1.0, out=out)  # diag - T U_G^2^T
        return out.view(-1)

    def update(self, rank: int = 100):
        for layer in list(self.diags.keys()):
            xxt_eigvecs, ggt_eigvecs = self.eigvecs[layer]
            lambda_vec = self.lambdas[layer].float().t().contiguous().view(-1)
            diag_vec = self.diags[layer].float().t().contiguous().view(-1)
            lr_a, lr_g, lr_lambda = self._dim_reduction(xxt_eigvecs, ggt_eigvecs, lambda_vec, rank)
            correction = self._corrected_diagonal(lr_a, lr_g, lr_lambda, diag_vec)
            self.state[layer] = (lr_a, lr_g, lr_lambda, correction)

    # --------------------------------------------------------------------------- invert
    def invert(self, add: Union[float, list, tuple] = 0., multiply: Union[float, list, tuple] = 1.):
        assert self.state, "State dict is empty. Did you call 'update' prior to this?"
        st = _lib.stream_ptr()
        lib = self._lib
        for index, (layer, value) in enumerate(self.state.items()):
            if not isinstance(add, float) and not isinstance(multiply, float):
                assert len(add) == len(multiply) == len(self.state)
                n_, s_ = add[index], multiply[index]
            else:
                n_, s_ = add, multiply
            lr_a, lr_g, lr_lambda, correction = value
            lr_a, lr_g = lr_a.contiguous(), lr_g.contiguous()
            n, a = lr_a.shape
            m, b = lr_g.shape
            r = a * b
            if r > 16384:
                raise ValueError(f"INF.invert: layer {index} keeps a {a} x {b} grid of eigen-directions (r = {r}); the "
                                 f"pre-sampler works on r x r matrices - choose a smaller `rank` in update()")
            reg_inv_correction = torch.empty_like(correction)
            reg_lambda = torch.empty_like(lr_lambda)
            _lib.check(lib.bk_inf_regularise(correction.data_ptr(), correction.numel(), lr_lambda.data_ptr(), r,
                                             float(n_), float(s_), reg_inv_correction.data_ptr(),
                                             reg_lambda.data_ptr(), st), "bk_inf_regularise")
            pre_sample = torch.empty(r, r, device=lr_a.device, dtype=torch.float32)
            nbytes = lib.bk_inf_presample_workspace_bytes(n, a, m, b)
            ws = self._ws.get(nbytes, lr_a.device)
            rc = _lib.check(lib.bk_inf_presample(lr_a.data_ptr(), lr_a.stride(0), n, a, lr_g.data_ptr(),
                                                 lr_g.stride(0), m, b, reg_inv_correction.data_ptr(),
                                                 reg_lambda.data_ptr(), pre_sample.data_ptr(), ws.data_ptr(),
                                                 nbytes, st), "bk_inf_presample")
            if rc > 0:
                raise RuntimeError(f"INF.invert: V^T V{' + I' if rc == 2 else ''} of layer {index} (rank {r}) is "
                                   f"not positive definite (the reference's cholesky() raises here too)")
            self.inv_state[layer] = (lr_a, lr_g, reg_inv_correction, pre_sample)
        self._staged = dict()

    # --------------------------------------------------------------------------- sample
    def _staged_inv(self, layer):
        if layer not in self._staged:
            a_, b_, _, p = self.inv_state[layer]
            self._staged[layer] = {"A": stage_operand(a_), "AT": stage_operand_t(a_), "G": stage_operand(b_),
                                   "GT": stage_operand_t(b_), "P": stage_operand(p)}
        return self._staged[layer]

    def sample(self, layer: Module, z: Optional[Tensor] = None) -> Tensor:
        """One INF posterior sample [d_out, d_in'].  `z` (flat [d_in' * d_out], the reference's X)
        replaces the Philox draw."""
        assert self.inv_state, "Inverse state dict is empty. Did you call 'invert' prior to this?"
        a_, b_, c, p = self.inv_state[layer]
        n, a = a_.shape
        m, b = b_.shape
        r = a * b
        ops = self._staged_inv(layer)
        if z is not None:
            z = z.to(c.device, torch.float32).contiguous().view(-1)
            assert z.numel() == n * m
        y_l = torch.empty(n * m, device=c.device, dtype=torch.float32)
        layer_id = list(self.inv_state.keys()).index(layer)
        _lib.check(self._lib.bk_diag_sample(y_l.data_ptr(), c.data_ptr(), n * m, 1, self.seed,
                                            self._next_sample_id(), layer_id, _lib.ptr(z), _lib.stream_ptr()),
                   "bk_diag_sample")
        unvec = y_l.view(m, n)                                              # the reference's reshape (:604)
        t1 = mm_nt_staged(ops["GT"], stage_operand_t(unvec), b, n, m)       # U_G^T unvec          [b, n]
        xq_t = mm_nt_staged(ops["AT"], stage_operand(t1), a, b, n)          # (U_G^T unvec U_A)^T  [a, b]
        qx = mm_nt_staged(stage_operand(xq_t.view(1, r)), ops["P"], 1, r, r)  # (P vec)^T          [1, r]
        wq = qx.view(b, a)                                                  # the reference's reshape (:608)
        t2 = mm_nt_staged(ops["A"], stage_operand(wq), n, b, a)             # U_A wq^T             [n, b]
        xps_t = mm_nt_staged(stage_operand(t2), ops["G"], n, m, b)          # (U_G wq U_A^T)^T     [n, m]
        out = torch.empty_like(y_l)
        _lib.check(self._lib.bk_inf_combine(out.data_ptr(), y_l.data_ptr(), c.data_ptr(), xps_t.data_ptr(),
                                            n * m, _lib.stream_ptr()), "bk_inf_combine")
        return out.view(n, m).t()
