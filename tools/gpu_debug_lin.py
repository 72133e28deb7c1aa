import sys, time
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200 import predictive as P
from bnn_kfac_b200.wrapper import LeNet5
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = LeNet5().to(dev); model.weight_init_uniform(0.05)
est = KFAC(model)
x = torch.rand(256, 1, 28, 28, device=dev); y = torch.randint(0, 10, (256,), device=dev)
loss = torch.nn.functional.cross_entropy(model(x), y); model.zero_grad(); loss.backward(); est.update(256); est.invert(1e2, 1e4)
def T(label, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    print(f"{label}: {(time.perf_counter() - t0) * 1e3:.2f} ms", flush=True); return r
for it in range(3):
    print("iter", it)
    pm = T("forward", lambda: torch.softmax(est.model(x), 1))
    go = P.argmax_grad_outputs(pm)
    layers = [l for l in list(est.model.modules())[1:] if l in est.state]
    Js = T("jacobians", lambda: P._layers_jacobians(pm, layers, go))
    total = torch.zeros(1, device=dev)
    for l, J in zip(layers, Js):
        Q, H = est.inv_state[l]
        st = T(f"  stage {tuple(Q.shape)} {tuple(H.shape)}", lambda: est._staged_factors(l))
        T("  quadform", lambda: P.kron_quadform(J.reshape(1, Q.shape[0], H.shape[0]), Q, H, triangular=True, out=total, accumulate=True, staged=st))
print("---- op-level timing of kron_quadform body (fc1: 401 x 120)")
import ctypes
from bnn_kfac_b200 import _lib
from bnn_kfac_b200.curvatures import _round8
lib = _lib.load(); st = _lib.stream_ptr()
l = layers[2]; Q, H = est.inv_state[l]; J = Js[2]
(q_hi, q_lo, ldq), (h_hi, h_lo, ldh) = est._staged_factors(l)
for it in range(3):
    print("rep", it)
    V = J.reshape(1, Q.shape[0], H.shape[0]).float().contiguous()
    Bn, dinp, dout = V.shape; ldv = _round8(dinp); ldu = _round8(dout)
    Vp = T("pad", lambda: (lambda z: (z[:, :dinp].copy_(V), z)[1])(torch.zeros(Bn, ldv, dout, device=dev)))
    t_hi = torch.empty(dout, Bn * ldv, dtype=torch.bfloat16, device=dev); t_lo = torch.empty_like(t_hi)
    T("transpose_split", lambda: lib.bk_transpose_split(Vp.data_ptr(), dout, Bn * ldv, dout, 1.0, 0, t_hi.data_ptr(), t_lo.data_ptr(), Bn * ldv, st))
    vt_hi = T("permute", lambda: t_hi.view(dout, Bn, ldv).permute(1, 0, 2).contiguous())
    vt_lo = t_lo.view(dout, Bn, ldv).permute(1, 0, 2).contiguous()
    u_hi = torch.zeros(Bn, dinp, ldu, dtype=torch.bfloat16, device=dev); u_lo = torch.zeros_like(u_hi)
    T("gemm QV", lambda: lib.bk_gemm_nt(q_hi.data_ptr(), q_lo.data_ptr(), ldq, 0, vt_hi.data_ptr(), vt_lo.data_ptr(), ldv, dout * ldv,
                              dinp, dout, dinp, Bn, 3, _lib.GEMM_TRI_A, 1.0, 0.0, 0, 0, 0, 0, 0, u_hi.data_ptr(), u_lo.data_ptr(), ldu, dinp * ldu, st))
    Wm = torch.empty(Bn, dinp, dout, device=dev)
    T("gemm UH", lambda: lib.bk_gemm_nt(u_hi.data_ptr(), u_lo.data_ptr(), ldu, dinp * ldu, h_hi.data_ptr(), h_lo.data_ptr(), ldh, 0,
                              dinp, dout, dout, Bn, 3, _lib.GEMM_TRI_B, 1.0, 0.0, Wm.data_ptr(), dout, dinp * dout, 0, 0, 0, 0, 0, 0, st))
    out = torch.zeros(1, device=dev)
    T("frob", lambda: lib.bk_frob_dot(out.data_ptr(), V.data_ptr(), dinp * dout, Wm.data_ptr(), dinp * dout, dinp * dout, Bn, 1, 1, st))
