"""A/B for "Philox fused into the sampling GEMM" (VERDICT r1 item 9): what does it cost that the N(0,1) noise of
`KFAC.sample` (models/curvatures.py:404-405) travels through HBM between the Philox kernel and the tensor-core GEMM
that consumes it?  Per configuration (S weight samples, B test inputs, cfg5 layer 4096 -> 4096) and per consumer

    implicit      Y2_s = (x~ L_A) Z_s       M = B,    N = 4096, K = 4097   (predictive._implicit_linear)
    materialised  T_s  = L_A Z_s            M = 4097, N = 4096, K = 4097   (sampling.matrix_normal_samples, TRI_A)

it times, with CUDA events on one stream (best of 5 after a warm-up):
    gemm_hbm     the batched GEMM reading S different noise matrices from HBM            (what ships)
    gemm_l2      the same GEMM with ONE noise matrix shared by all samples (stride 0): 33.5 MB, L2-resident -
                 the operand traffic a producer-side generator would have (no HBM read of Z at all)
    philox       the generator alone (writes S x 33.5 MB of bf16 noise)
    serial       philox, then gemm_hbm on the same stream                                (no overlap at all)
    overlapped   philox of chunk i + 1 on a side stream under gemm_hbm of chunk i          (what ships)
"""
import sys

import torch

sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib

L = _lib.load()
_lib.require_device()
dev = torch.device("cuda:0")
d_in_p, d_out = 4097, 4096
ldz = (d_in_p + 7) // 8 * 8


def ev(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def run(S, B, chunk):
    out = {}
    z = torch.empty(2, chunk, d_out, ldz, dtype=torch.bfloat16, device=dev)     # two chunks (double buffer)
    la = torch.randn(d_in_p, ldz, device=dev).to(torch.bfloat16)
    for name, M, flags in (("implicit", B, 0), ("materialised", d_in_p, _lib.GEMM_TRI_A)):
        a = la if name == "materialised" else torch.randn(B, ldz, device=dev).to(torch.bfloat16)
        o_hi = torch.empty(chunk, M, d_out, dtype=torch.bfloat16, device=dev)
        nchunks = (S + chunk - 1) // chunk

        def philox(buf, c):
            _lib.check(L.bk_philox_normal(7, c * chunk, 0, d_out, d_in_p, chunk, 0, 0, 0, z[buf].data_ptr(), 0, ldz,
                                          d_out * ldz, _lib.stream_ptr()), "philox")

        def gemm(buf, shared):
            _lib.check(L.bk_gemm_nt(a.data_ptr(), 0, ldz, 0, z[buf].data_ptr(), 0, ldz, 0 if shared else d_out * ldz,
                                    M, d_out, d_in_p, chunk, _lib.BK_PREC_BF16, flags, 1.0, 0.0, 0, 0, 0, 0, 0,
                                    o_hi.data_ptr(), 0, d_out, M * d_out, _lib.stream_ptr()), "gemm")

        side = torch.cuda.Stream()
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        freed = [torch.cuda.Event(), torch.cuda.Event()]

        def overlapped():
            main = torch.cuda.current_stream()
            side.wait_stream(main)
            with torch.cuda.stream(side):
                philox(0, 0)
                ready[0].record(side)
            for c in range(nchunks):
                b = c & 1
                if c + 1 < nchunks:
                    with torch.cuda.stream(side):
                        if c >= 1:
                            side.wait_event(freed[b ^ 1])
                        philox(b ^ 1, c + 1)
                        ready[b ^ 1].record(side)
                main.wait_event(ready[b])
                gemm(b, False)
                freed[b].record(main)

        t = {"gemm_hbm": ev(lambda: [gemm(c & 1, False) for c in range(nchunks)]),
             "gemm_l2": ev(lambda: [gemm(c & 1, True) for c in range(nchunks)]),
             "philox": ev(lambda: [philox(c & 1, c) for c in range(nchunks)]),
             "serial": ev(lambda: [(philox(c & 1, c), gemm(c & 1, False)) for c in range(nchunks)]),
             "overlapped": ev(overlapped)}
        out[name] = t
        flops = 2.0 * M * d_out * d_in_p * chunk * nchunks * (0.5 if flags else 1.0)
        print(f"S={S} B={B} chunk={chunk} {name:12s}: " + "  ".join(f"{k} {v:7.3f} ms" for k, v in t.items())
              + f"  | gemm_hbm {flops / t['gemm_hbm'] / 1e9:6.0f} TFLOP/s  Z-through-HBM cost "
                f"{(t['gemm_hbm'] / t['gemm_l2'] - 1) * 100:+.1f} %  generator exposed after overlap "
                f"{(t['overlapped'] / t['gemm_hbm'] - 1) * 100:+.1f} %", flush=True)
    return out


if __name__ == "__main__":
    run(16, 1024, 16)
    run(100, 256, 20)
