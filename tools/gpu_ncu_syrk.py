"""ncu target: the grouped factor SYRK of one cfg5 update (7 x 4096^2 x 4096 bf16, lower-only accumulation, staged
K-major operands) launched three times.  Used as
    ncu --set full --clock-control none --import-source on -k regex:umma_syrk_grouped -s 1 -c 1 -o ... python tools/gpu_ncu_syrk.py
"""
import ctypes as C
import sys

import torch

sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib

L = _lib.load()
_lib.require_device()
dev = torch.device("cuda:0")
n, d, cnt = 4096, 4096, 7
flags = int(sys.argv[1]) if len(sys.argv) > 1 else 1
his = [torch.randn(d, n, device=dev).to(torch.bfloat16) for _ in range(cnt)]
sts = [torch.zeros(d, d, device=dev) for _ in range(cnt)]
args = ((C.c_void_p * cnt)(*[t.data_ptr() for t in sts]), (C.c_longlong * cnt)(*[d] * cnt),
        (C.c_void_p * cnt)(*[t.data_ptr() for t in his]), (C.c_void_p * cnt)(*[t.data_ptr() for t in his]),
        (C.c_longlong * cnt)(*[n] * cnt), (C.c_int * cnt)(*[n] * cnt), (C.c_int * cnt)(*[d] * cnt),
        (C.c_float * cnt)(*[1.0 / n] * cnt), (C.c_float * cnt)(*[1.0] * cnt))
for _ in range(3):
    _lib.check(L.bk_syrk_accum_staged_grouped(*args, cnt, 1, flags, _lib.stream_ptr()), "grouped")
torch.cuda.synchronize()
print("ok")
