"""Bring-up: the batched slab patterns of the block-Jacobi rounds through bk_gemm_nt (bf16x3)."""
import sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import stage_operand
dev = torch.device("cuda:0")
lib = _lib.load()
torch.manual_seed(0)
def run(dp, off, npairs):
    S = torch.randn(dp, dp, device=dev); S = S + S.t()
    QT = torch.randn(npairs, 128, 128, device=dev)
    s_hi, s_lo, ld = stage_operand(S)
    q_hi, q_lo, ldq = stage_operand(QT.reshape(-1, 128))
    # step 1: T[pair rows, :] = QT S[:, pair cols]^T  -> O only
    t_hi = torch.zeros(dp, dp, dtype=torch.bfloat16, device=dev); t_lo = torch.zeros_like(t_hi)
    rc = lib.bk_gemm_nt(q_hi.data_ptr(), q_lo.data_ptr(), 128, 128 * 128,
                        s_hi.data_ptr() + 2 * off, s_lo.data_ptr() + 2 * off, ld, 128,
                        128, dp, 128, npairs, 3, 0, 1.0, 0.0, 0, 0, 0, 0, 0,
                        t_hi.data_ptr() + 2 * off * dp, t_lo.data_ptr() + 2 * off * dp, dp, 128 * dp, 0)
    T = (t_hi.float() + t_lo.float())
    want = torch.zeros(dp, dp, device=dev, dtype=torch.float64)
    for b in range(npairs):
        o = off + 128 * b
        want[o:o + 128, :] = QT[b].double() @ S[:, o:o + 128].double().t()
    sel = slice(off, off + 128 * npairs)
    e1 = (T[sel].double() - want[sel]).norm() / want[sel].norm()
    # step 2: S'[:, pair cols] = T[:, pair cols] QT^T -> C and O
    Tm = torch.randn(dp, dp, device=dev)
    a_hi, a_lo, lda = stage_operand(Tm)
    C = torch.zeros(dp, dp, device=dev)
    o_hi = torch.zeros(dp, dp, dtype=torch.bfloat16, device=dev); o_lo = torch.zeros_like(o_hi)
    rc2 = lib.bk_gemm_nt(a_hi.data_ptr() + 2 * off, a_lo.data_ptr() + 2 * off, lda, 128,
                         q_hi.data_ptr(), q_lo.data_ptr(), 128, 128 * 128,
                         dp, 128, 128, npairs, 3, 0, 1.0, 0.0, C.data_ptr() + 4 * off, dp, 128, 0, 0,
                         o_hi.data_ptr() + 2 * off, o_lo.data_ptr() + 2 * off, dp, 128, 0)
    want2 = torch.zeros(dp, dp, device=dev, dtype=torch.float64)
    for b in range(npairs):
        o = off + 128 * b
        want2[:, o:o + 128] = Tm[:, o:o + 128].double() @ QT[b].double().t()
    e2 = (C[:, sel].double() - want2[:, sel]).norm() / want2[:, sel].norm()
    e3 = ((o_hi.float() + o_lo.float())[:, sel].double() - want2[:, sel]).norm() / want2[:, sel].norm()
    print(f"dp={dp} off={off} npairs={npairs}: rc={rc},{rc2} step1 {e1:.2e} step2 C {e2:.2e} O {e3:.2e}", flush=True)
for cfg in [(128, 0, 1), (256, 0, 2), (256, 64, 1), (384, 0, 3), (384, 64, 2), (1152, 0, 9), (1152, 64, 8)]:
    run(*cfg)
