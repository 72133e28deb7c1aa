"""Development timing: the grouped factor SYRK (7 x 4096^2 x 4096, bf16) under its A/B switches, interleaved
repetitions so that clock / thermal drift hits every variant alike."""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib

L = _lib.load()
_lib.require_device()
dev = torch.device("cuda:0")
n, d = 4096, 4096


def make(cnt):
    his = [torch.randn(d, n, device=dev).to(torch.bfloat16) for _ in range(cnt)]
    sts = [torch.zeros(d, d, device=dev) for _ in range(cnt)]
    a = dict(states=(C.c_void_p * cnt)(*[t.data_ptr() for t in sts]), lds=(C.c_longlong * cnt)(*[d] * cnt),
             hi=(C.c_void_p * cnt)(*[t.data_ptr() for t in his]), lo=(C.c_void_p * cnt)(*[t.data_ptr() for t in his]),
             ldt=(C.c_longlong * cnt)(*[n] * cnt), ns=(C.c_int * cnt)(*[n] * cnt), ds=(C.c_int * cnt)(*[d] * cnt),
             al=(C.c_float * cnt)(*[1.0 / n] * cnt), be=(C.c_float * cnt)(*[1.0] * cnt), keep=(his, sts))
    xs = [torch.randn(n, d, device=dev).to(torch.bfloat16) for _ in range(cnt)]      # row-major activations
    a["x"] = (C.c_void_p * cnt)(*[t.data_ptr() for t in xs])
    a["ldx"] = (C.c_longlong * cnt)(*[d] * cnt)
    a["keep2"] = xs
    return a


def timed(a, cnt, flags, reps=10):
    row_major = bool(flags & 4)
    f = lambda: _lib.check(L.bk_syrk_accum_staged_grouped(a["states"], a["lds"], a["x"] if row_major else a["hi"],
                                                          a["x"] if row_major else a["lo"],
                                                          a["ldx"] if row_major else a["ldt"], a["ns"],
                                                          a["ds"], a["al"], a["be"], cnt, 1, flags,
                                                          _lib.stream_ptr()), "grouped")
    f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        f()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


for cnt in (7, 3, 1):
    a = make(cnt)
    res = {}
    for rnd in range(3):
        for tune in (0, 1):
            for flags in (1, 0, 5):
                L.bk_set_syrk_tuning(tune)
                res.setdefault((tune, flags), []).append(timed(a, cnt, flags))
    L.bk_set_syrk_tuning(0)
    for (tune, flags), v in sorted(res.items()):
        us = min(v)
        print(f"factors={cnt} tuning={tune} (dedup={'off' if tune & 1 else 'on'}, tail-split={'on' if tune & 2 else 'off'}) "
              f"{'lower-only' if flags & 1 else 'mirrored'}{' ROW-MAJOR bf16 activations (no staging)' if flags & 4 else ''}: best {us:7.1f} us  all {[round(x, 1) for x in v]}  "
              f"{cnt * d * (d + 1) * n / us / 1e6:7.1f} TFLOP/s alg", flush=True)
