"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads, and exports exactly the
symbols include/bk_kfac.h declares; the Python mirror of the reference API keeps its surface."""
import inspect
import re
import subprocess
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "bk_kfac.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bk_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from bnn_kfac_b200 import _build, _lib
    path = _build.build()
    assert path.exists()
    out = subprocess.run(["nm", "-D", "--defined-only", str(path)], capture_output=True, text=True).stdout
    exported = sorted(set(re.findall(r" T (bk_[a-z0-9_]+)", out)))
    declared = _declared()
    assert declared, "no declarations parsed from include/bk_kfac.h"
    assert exported == declared
    assert sorted(_lib.SIGNATURES) == declared          # ctypes table covers the header
    lib = _lib.load()                                   # strict: raises on a missing symbol
    assert lib.bk_version().decode().startswith("bk_kfac")


def test_no_torch_types_in_abi_and_no_undefined_torch_symbols():
    from bnn_kfac_b200 import _build
    out = subprocess.run(["nm", "-D", "--undefined-only", str(_build.build())], capture_output=True,
                         text=True).stdout
    assert "c10" not in out and "at::" not in out and "torch" not in out


def test_product_has_no_cpu_fallback_and_no_oracle_import():
    pkg = ROOT / "bnn_kfac_b200"
    for p in pkg.glob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
    if not torch.cuda.is_available():
        from bnn_kfac_b200 import _lib
        from bnn_kfac_b200.curvatures import KFAC
        with pytest.raises(_lib.BkError):
            KFAC(torch.nn.Sequential(torch.nn.Linear(3, 2)))


def test_reference_api_surface():
    """Constructor / method names and arguments of models/curvatures.py and models/wrapper.py."""
    from bnn_kfac_b200 import curvatures as C, wrapper as W
    for cls in (C.KFAC, C.Diagonal):
        params = list(inspect.signature(cls.__init__).parameters)
        assert params[:3] == ["self", "model", "layer_types"]
        for m in ("update", "invert", "sample", "sample_and_replace", "save", "load", "_replace"):
            assert hasattr(cls, m)
        inv = inspect.signature(cls.invert).parameters
        assert list(inv)[:3] == ["self", "add", "multiply"]
        assert inv["add"].default == 0. and inv["multiply"].default == 1.
    for fn in ("get_nb_parameters", "save", "load", "train", "eval", "accuracy"):
        assert callable(getattr(W, fn))
    assert list(inspect.signature(W.train).parameters) == ["model", "device", "data", "criterion",
                                                           "optimizer", "epochs"]
    assert list(inspect.signature(W.eval).parameters) == ["model", "device", "data"]
    assert sum(p.numel() for p in W.BaseNet_750().parameters()) == 748
    assert sum(p.numel() for p in W.BaseNet_15k().parameters()) == 15080


def test_wrapper_eval_and_accuracy_cpu():
    from bnn_kfac_b200 import wrapper as W
    torch.manual_seed(0)
    net = W.BaseNet_750()
    net.weight_init_uniform(0.2)
    data = [(torch.rand(4, 1, 28, 28), torch.randint(0, 10, (4,))) for _ in range(3)]
    preds, targets = W.eval(net, "cpu", data)
    assert preds.shape == (12, 10) and targets.shape == (12,)
    assert torch.allclose(preds.sum(1), torch.ones(12), atol=1e-5)
    ref = torch.softmax(torch.cat([net(x) for x, _ in data]), dim=1)
    assert torch.allclose(preds, ref, atol=1e-6)
    acc = W.accuracy(preds, targets)
    assert 0.0 <= acc <= 100.0
    for layer in net.modules():
        if layer.__class__.__name__ in ("Linear", "Conv2d"):
            assert layer.bias.abs().max().item() == 0.0
            assert layer.weight.abs().max().item() <= 0.2
