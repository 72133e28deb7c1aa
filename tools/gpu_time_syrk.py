"""Development timing: SYRK variants (cta_group, beta, precision, d) on a staged operand."""
import os, sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
n = 4096
def run(d, cg, prec, beta, ld=None, reps=20):
    ld = ld or d
    hi = torch.randn(d, n, device=dev).to(torch.bfloat16)
    lo = torch.zeros_like(hi)
    st = torch.zeros(d, ld, device=dev)
    L.bk_set_cta_group(cg)
    f = lambda: L.bk_syrk_accum_staged(st.data_ptr(), ld, hi.data_ptr(), lo.data_ptr(), n, n, d, 1.0 / n, beta, prec, _lib.stream_ptr())
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    print(f"d={d} ld={ld} cg={cg} prec={prec} beta={beta}: {us:.1f} us  {d*(d+1)*n/us/1e6:.1f} TFLOP/s alg", flush=True)
for d in (4096, 2048, 8192):
    for cg in (1, 2):
        for beta in (1.0, 0.0):
            run(d, cg, 1, beta)
run(4096, 2, 3, 1.0)
run(4096, 2, 1, 1.0, ld=4100)
