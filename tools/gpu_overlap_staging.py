"""Probe: does the fp32 -> bf16 K-major staging kernel (transpose_split64, 16.5 KB of shared memory per CTA) share
SMs with the persistent grouped SYRK (lower-only epilogue: 209 KB per CTA) when both are in flight on two streams?
Prints each alone, both back to back, and both concurrently, for 3 operands of 4096 x 4096."""
import ctypes as C
import sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
n = d = 4096
cnt = 3
xs = [torch.randn(n, d, device=dev) for _ in range(cnt)]
t_hi = [torch.empty(d, n, dtype=torch.bfloat16, device=dev) for _ in range(cnt)]
s_hi = [torch.randn(d, n, device=dev).to(torch.bfloat16) for _ in range(cnt)]
sts = [torch.zeros(d, d, device=dev) for _ in range(cnt)]
args = ((C.c_void_p * cnt)(*[t.data_ptr() for t in sts]), (C.c_longlong * cnt)(*[d] * cnt),
        (C.c_void_p * cnt)(*[t.data_ptr() for t in s_hi]), (C.c_void_p * cnt)(*[t.data_ptr() for t in s_hi]),
        (C.c_longlong * cnt)(*[n] * cnt), (C.c_int * cnt)(*[n] * cnt), (C.c_int * cnt)(*[d] * cnt),
        (C.c_float * cnt)(*[1.0 / n] * cnt), (C.c_float * cnt)(*[1.0] * cnt))


def syrk():
    _lib.check(L.bk_syrk_accum_staged_grouped(*args, cnt, 1, 1, _lib.stream_ptr()), "grouped")


def stage():
    for x, t in zip(xs, t_hi):
        _lib.check(L.bk_transpose_split(x.data_ptr(), d, n, d, 1.0, 0, t.data_ptr(), 0, n, _lib.stream_ptr()), "stage")


def ev(fn, reps=10):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(3):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    return best * 1e3


side = torch.cuda.Stream()


def both(first):
    main = torch.cuda.current_stream()
    side.wait_stream(main)
    if first == "stage":
        with torch.cuda.stream(side):
            stage()
        syrk()
    else:
        syrk()
        with torch.cuda.stream(side):
            stage()
    main.wait_stream(side)


print(f"stage {ev(stage):.1f} us  syrk {ev(syrk):.1f} us  serial {ev(lambda: (stage(), syrk())):.1f} us  "
      f"concurrent(stage first) {ev(lambda: both('stage')):.1f} us  concurrent(syrk first) {ev(lambda: both('syrk')):.1f} us")
