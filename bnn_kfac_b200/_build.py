"""In-tree build of libbk_kfac.so (sm_100a only) with plain nvcc.

The shared library is a C-ABI library (include/bk_kfac.h): it does not link against torch or
libcuda (cuTensorMapEncodeTiled is resolved at run time through cudaGetDriverEntryPoint).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
OUT_DIR = PKG_DIR / "_C"
LIB_PATH = OUT_DIR / "libbk_kfac.so"
STAMP = OUT_DIR / "libbk_kfac.stamp"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "-shared", "-cudart", "shared",
]


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) +
                    [PKG_DIR.parent / "include" / "bk_kfac.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found: libbk_kfac.so cannot be built")
    return nvcc


def is_current() -> bool:
    """Does the built library match the sources (digest stamp written by build())?"""
    return LIB_PATH.exists() and STAMP.exists() and STAMP.read_text() == _digest()


def _headers_digest() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cuh")) + [PKG_DIR.parent / "include" / "bk_kfac.h"]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(args):
    nvcc, src, obj, verbose = args
    cmd = [nvcc, *[f for f in NVCC_FLAGS if f not in ("-shared",)], "-c", "-o", str(obj), str(src)]
    if verbose:
        cmd += ["-Xptxas", "-v"]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    return src, proc.returncode, " ".join(cmd), proc.stdout + proc.stderr


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA source (one object per file, in parallel, re-using objects whose source and headers
    did not change) and link them into one shared library. Returns the path to the .so."""
    from concurrent.futures import ThreadPoolExecutor
    OUT_DIR.mkdir(exist_ok=True)
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text() == digest:
        return LIB_PATH
    nvcc = find_nvcc()
    obj_dir = OUT_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    hdr = _headers_digest()
    jobs, objs = [], []
    for src in sources():
        obj = obj_dir / (src.stem + ".o")
        tag = obj_dir / (src.stem + ".digest")
        want = hashlib.sha256(hdr.encode() + src.read_bytes()).hexdigest()
        objs.append(obj)
        if force or verbose or not obj.exists() or not tag.exists() or tag.read_text() != want:
            jobs.append((nvcc, src, obj, verbose, tag, want))
    with ThreadPoolExecutor(max_workers=min(len(jobs), os.cpu_count() or 1) or 1) as pool:
        for (src, rc, cmd, out), job in zip(pool.map(_compile_one, [j[:4] for j in jobs]), jobs):
            if rc != 0:
                raise RuntimeError("nvcc failed:\n" + cmd + "\n" + out)
            if verbose:
                print(out)
            job[4].write_text(job[5])
    cmd = [nvcc, *NVCC_FLAGS, "-o", str(LIB_PATH), *map(str, objs)]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    if proc.returncode != 0:
        raise RuntimeError("nvcc link failed:\n" + " ".join(cmd) + "\n" + proc.stdout + proc.stderr)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    import sys
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
