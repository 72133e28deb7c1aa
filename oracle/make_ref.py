"""Stages the reference's own hot-path modules as an importable, UNMODIFIED copy under oracle/_ref/.

TEST / BENCH INFRASTRUCTURE, not product: only `bench.py --impl reference`, `bench.py`'s `cpu_baseline`
leg and `tests/` may import what this script produces.

The reference (TianmingQiu/BNN_KFAC) is pure Python: there is nothing to compile.  `/root/reference` does not
exist on the GPU box, so the files the timed CPU arm needs travel as a build output (oracle/_ref/ is
git-ignored, not gpurun-ignored), exactly like a compiled `oracle/_ref/*.so` would:

    oracle/_ref/models/{__init__,curvatures,utilities,wrapper}.py   byte-for-byte copies (sha256 recorded)
    oracle/_ref/matplotlib/{__init__,pyplot}.py                      empty stubs: models/utilities.py:19 imports
                                                                     pyplot at module import and matplotlib is
                                                                     not installed in this image
    oracle/_ref/MANIFEST.json                                        source path + sha256 of every copied file

Run in the build container (`python oracle/make_ref.py`, also called by `__graft_entry__.build()` whenever
/root/reference is present).  Nothing under oracle/_ref/ is ever committed.
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

REF = Path("/root/reference")
OUT = Path(__file__).resolve().parent / "_ref"
FILES = ["models/__init__.py", "models/curvatures.py", "models/utilities.py", "models/wrapper.py"]


def make(ref: Path = REF, out: Path = OUT) -> Path:
    if not ref.exists():
        raise FileNotFoundError(f"{ref} is not present (build container only)")
    manifest = {}
    for rel in FILES:
        src, dst = ref / rel, out / rel
        dst.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(src, dst)
        manifest[rel] = {"source": str(src), "sha256": hashlib.sha256(dst.read_bytes()).hexdigest()}
    stub = out / "matplotlib"
    stub.mkdir(exist_ok=True)
    (stub / "__init__.py").write_text('"""empty stub: the reference imports matplotlib.pyplot at import time"""\n')
    (stub / "pyplot.py").write_text('"""empty stub (never called on the curvature path)"""\n')
    (out / "MANIFEST.json").write_text(json.dumps(manifest, indent=1) + "\n")
    return out


def available(out: Path = OUT) -> bool:
    return (out / "MANIFEST.json").exists() and all((out / rel).exists() for rel in FILES)


def load(out: Path = OUT):
    """Imports the staged reference and returns its `models.curvatures` module (KFAC, Diagonal, ...)."""
    import importlib
    import warnings
    if not available(out):
        raise FileNotFoundError(f"{out} is missing: run `python oracle/make_ref.py` in the build container")
    manifest = json.loads((out / "MANIFEST.json").read_text())
    for rel, meta in manifest.items():
        got = hashlib.sha256((out / rel).read_bytes()).hexdigest()
        if got != meta["sha256"]:
            raise RuntimeError(f"oracle/_ref/{rel} does not match its recorded sha256: not the unmodified reference")
    if str(out) not in sys.path:
        sys.path.insert(0, str(out))
    warnings.filterwarnings("ignore", category=UserWarning)
    warnings.filterwarnings("ignore", category=FutureWarning)
    return importlib.import_module("models.curvatures")


if __name__ == "__main__":
    print(make())
