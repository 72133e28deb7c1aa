"""Probe: do the persistent Philox generator and a persistent tcgen05 GEMM share SMs when launched on two streams
with no dependency between them?  Prints the time of each alone, both back to back on one stream, and both
concurrently (two streams), for the implicit-predictive shapes (M = 256 / 1024, N = 4096, K = 4097, 20 samples)."""
import sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
d_in_p, d_out, S = 4097, 4096, 20
ldz = 4104
z = torch.empty(2, S, d_out, ldz, dtype=torch.bfloat16, device=dev)


def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for M in (256, 1024, 4097):
    a = torch.randn(M, ldz, device=dev).to(torch.bfloat16)
    o = torch.empty(S, M, d_out, dtype=torch.bfloat16, device=dev)

    def philox(buf):
        _lib.check(L.bk_philox_normal(7, 0, 0, d_out, d_in_p, S, 0, 0, 0, z[buf].data_ptr(), 0, ldz, d_out * ldz,
                                      _lib.stream_ptr()), "philox")

    def gemm(buf):
        _lib.check(L.bk_gemm_nt(a.data_ptr(), 0, ldz, 0, z[buf].data_ptr(), 0, ldz, d_out * ldz, M, d_out, d_in_p, S,
                                _lib.BK_PREC_BF16, 0, 1.0, 0.0, 0, 0, 0, 0, 0, o.data_ptr(), 0, d_out, M * d_out,
                                _lib.stream_ptr()), "gemm")

    philox(0); philox(1)
    side = torch.cuda.Stream()

    def both(first):
        main = torch.cuda.current_stream()
        side.wait_stream(main)
        if first == "philox":
            with torch.cuda.stream(side):
                philox(1)
            gemm(0)
        else:
            gemm(0)
            with torch.cuda.stream(side):
                philox(1)
        main.wait_stream(side)

    t_p, t_g = ev(lambda: philox(1)), ev(lambda: gemm(0))
    t_s = ev(lambda: (philox(1), gemm(0)))
    t_c1, t_c2 = ev(lambda: both("philox")), ev(lambda: both("gemm"))
    print(f"M={M}: philox {t_p:.3f}  gemm {t_g:.3f}  serial {t_s:.3f}  concurrent(philox first) {t_c1:.3f}  "
          f"concurrent(gemm first) {t_c2:.3f}  ms", flush=True)
