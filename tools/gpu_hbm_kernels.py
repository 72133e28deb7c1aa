"""Achieved HBM bandwidth of the elementwise / reduction / staging kernels at cfg5 sizes
(4096 x 4097 layers, batch 4096), CUDA-event timed, L2 flushed between repetitions by cycling over
buffers larger than the 126 MB L2.  Prints a markdown table (copied to profiles/)."""
import json, os, sys
import torch
sys.path.insert(0, ".")
from bnn_kfac_b200 import _lib
L = _lib.load(); _lib.require_device()
dev = torch.device("cuda:0")
PEAK = 6544.7
if os.path.exists("MEASURED_PEAKS.json"):
    PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
st = _lib.stream_ptr()
rows = []

def timeit(name, fn_list, bytes_per_call, reps=4):
    """fn_list: independent closures over different buffers (rotated to defeat the L2)."""
    for f in fn_list: f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps):
        for f in fn_list: f()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * len(fn_list))
    gbs = bytes_per_call / us / 1e3
    rows.append((name, bytes_per_call / 1e6, us, gbs, gbs / PEAK))
    print(f"{name:58s} {bytes_per_call/1e6:8.1f} MB {us:8.1f} us {gbs:8.0f} GB/s {gbs/PEAK:5.2f}", flush=True)

NB = 4  # buffers per kernel: 4 x (67..200 MB) > L2
do, di, n = 4096, 4096, 4096
dp = di + 1
# staging: fp32 [n, d] -> bf16 [d, n] (+ column sums)
xs = [torch.randn(n, di, device=dev) for _ in range(NB)]
hi = [torch.empty(di, n, dtype=torch.bfloat16, device=dev) for _ in range(NB)]
lo = [torch.empty(di, n, dtype=torch.bfloat16, device=dev) for _ in range(NB)]
for variant in (0,):
    L.bk_set_cta_group(variant << 8)
    timeit(f"transpose_split64 v{variant} fp32[4096,4096] -> bf16 K-major (bf16)", [lambda i=i: L.bk_transpose_split(xs[i].data_ptr(), di, n, di, 1.0, 0, hi[i].data_ptr(), 0, n, st) for i in range(NB)], n * di * 6)
    timeit(f"transpose_split64 v{variant} ... hi + lo (bf16x3)", [lambda i=i: L.bk_transpose_split(xs[i].data_ptr(), di, n, di, 1.0, 0, hi[i].data_ptr(), lo[i].data_ptr(), n, st) for i in range(NB)], n * di * 8)
L.bk_set_cta_group(0)
timeit("convert_split fp32[4096,4096] -> bf16 hi+lo", [lambda i=i: L.bk_convert_split(xs[i].data_ptr(), di, n, di, 1.0, 0, hi[i].data_ptr(), lo[i].data_ptr(), n, st) for i in range(NB)], n * di * 8)
# diagonal curvature
wg = [torch.randn(do, di, device=dev) for _ in range(NB)]
bg = torch.randn(do, device=dev)
sts = [torch.rand(do, dp, device=dev) for _ in range(NB)]
invs = [torch.empty(do, dp, device=dev) for _ in range(NB)]
timeit("diag_accum state += [W.grad|b.grad]^2*bs  [4096,4097]", [lambda i=i: L.bk_diag_accum(sts[i].data_ptr(), wg[i].data_ptr(), bg.data_ptr(), do, di, 32.0, 1.0, st) for i in range(NB)], do * dp * 12)
timeit("diag_invert 1/sqrt(s*state+n)", [lambda i=i: L.bk_diag_invert(invs[i].data_ptr(), sts[i].data_ptr(), do * dp, 0.04, 200.0, st) for i in range(NB)], do * dp * 8)
timeit("diag_sample Philox * inv (1 sample)", [lambda i=i: L.bk_diag_sample(sts[i].data_ptr(), invs[i].data_ptr(), do * dp, 1, 7, 0, 0, 0, st) for i in range(NB)], do * dp * 8)
J = [torch.randn(8, do * dp, device=dev) for _ in range(NB)]
q = torch.empty(8, device=dev)
timeit("diag_quadform sum_j J^2 h  (8 rows x 16.8M)", [lambda i=i: L.bk_diag_quadform(q.data_ptr(), J[i].data_ptr(), do * dp, invs[i].data_ptr(), do * dp, 8, st) for i in range(NB)], do * dp * 4 * 9)
# Philox noise operand
zs = [torch.empty(do, 4104, dtype=torch.bfloat16, device=dev) for _ in range(NB)]
timeit("philox_normal bf16 Z^T [4096, 4097] (1 sample, compute-bound)", [lambda i=i: L.bk_philox_normal(7, i, 0, do, dp, 1, 0, 0, 0, zs[i].data_ptr(), 0, 4104, do * 4104, st) for i in range(NB)], do * dp * 2)
# dense Fisher dominance (P = 15080) and rank-1 accumulate
P = 15080
H = [torch.randn(P, P, device=dev) for _ in range(2)]
bb = torch.tensor([0, 5000], dtype=torch.int32, device=dev); be = torch.tensor([5000, P], dtype=torch.int32, device=dev)
out3 = torch.empty(3, dtype=torch.float64, device=dev)
timeit("dominance |H + tau I| sums, P = 15080", [lambda i=i: L.bk_dominance(H[i].data_ptr(), P, P, 1e-5, bb.data_ptr(), be.data_ptr(), 2, out3.data_ptr(), st) for i in range(2)], P * P * 4)
g = torch.randn(P, device=dev)
timeit("ger_accum H += bs * g g^T, P = 15080", [lambda i=i: L.bk_ger_accum(H[i].data_ptr(), P, g.data_ptr(), P, 32.0, 1.0, st) for i in range(2)], P * P * 8)
# predictive moments over samples
lg = [torch.randn(64, 4096, 1000, device=dev) for _ in range(2)]
mean = torch.empty(4096, 1000, device=dev); msq = torch.empty_like(mean)
timeit("predictive_moments softmax mean/meansq [64, 4096, 1000]", [lambda i=i: L.bk_predictive_moments(lg[i].data_ptr(), 64, 4096, 1000, 0, mean.data_ptr(), msq.data_ptr(), st) for i in range(2)], 64 * 4096 * 1000 * 4)
# peer-exchange kernels with every source in LOCAL memory (their HBM side): tile pack of a 4097-wide factor, fused
# sum-of-8 + unpack (what an owner runs per factor at world size 8, there with 7 of the 8 sources behind NVLink)
import ctypes as C
d = 4097
T = (d + 31) // 32
tp = T * (T + 1) // 2 * 1024
facs = [torch.randn(d, d, device=dev) for _ in range(NB)]
packs = [torch.empty(tp, device=dev) for _ in range(8)]
one = lambda t: ((C.c_void_p * 1)(t.data_ptr()), (C.c_longlong * 1)(t.stride(0)), (C.c_int * 1)(d))
timeit("tile_pack lower triangle of [4097, 4097] -> 32x32 tiles", [lambda i=i: L.bk_tile_pack(*one(facs[i]), None, 1, packs[i].data_ptr(), st) for i in range(NB)], d * (d + 1) // 2 * 4 + tp * 4)
outs = [torch.empty(d, d, device=dev) for _ in range(2)]
srcs = (C.c_void_p * 8)(*[t.data_ptr() for t in packs])
timeit("peer_tile_unpack<8> sum of 8 packed sources -> mirrored [4097, 4097]", [lambda i=i: L.bk_peer_tile_unpack(*one(outs[i]), 1, srcs, 8, 0.125, 1, st) for i in range(2)], 8 * tp * 4 + d * d * 4)
srcs1 = [(C.c_void_p * 1)(packs[i].data_ptr()) for i in range(NB)]
timeit("peer_tile_unpack<1> packed -> lower-triangular [4097, 4097]", [lambda i=i: L.bk_peer_tile_unpack(*one(outs[i % 2]), 1, srcs1[i], 1, 1.0, 0, st) for i in range(NB)], tp * 4 + d * d * 4)
with open("gpurun_out/hbm_kernels.md", "w") as f:
    f.write("| kernel (cfg5-sized operands) | algorithmic MB | us | GB/s | of measured HBM peak (%.1f GB/s) |\n|---|---:|---:|---:|---:|\n" % PEAK)
    for r in rows:
        f.write(f"| {r[0]} | {r[1]:.1f} | {r[2]:.1f} | {r[3]:.0f} | {r[4]:.2f} |\n")
