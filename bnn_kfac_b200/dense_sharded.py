"""Dense-Fisher Laplace across GPUs (BASELINE config 3 at P = 15 080: H is 910 MB).

The reference is single-device (hessian/classification_ll_dense_kernel_diag.py:68-91 accumulates H by rank-1
updates, sampling_free/classification/classification_ll_dense.py:108-109,160-161 takes pinverse(H + tau I) and
|J H_inv J^T|).  Sharding defined here:

  accumulate   the gradient rows are split over ranks; each rank runs ONE tensor-core SYRK over its rows
               (dense.dense_fisher, normalised by the GLOBAL gradient count).
  exchange     ONE reduce-scatter by row block: H is cut into nb-row blocks dealt round-robin to the ranks
               (block-cyclic), the send buffer is laid out in owner order, `reduce_scatter_tensor` leaves every
               rank with the summed rows it owns.  H is never replicated again.
  dominance    sum|diag|, sum|all|, sum|kernel blocks| of H + tau I (hessian/utils.py:4-23) from each rank's rows
               (bk_dominance_rows) + one all-reduce of three doubles.
  variance     |J (H + tau I)^-1 J^T| = ||Y||^2 with Y = J L^-T, H + tau I = L L^T, by a right-looking blocked
               Cholesky on the SHARDED rows: per step the owner factorises and inverts the nb x nb diagonal block
               in fp64 (bk_chol_trinv_f64), broadcasts W = L_kk^-1 (50 KB), every rank forms its panel rows
               X = A[:, k] W^T, the panel is all-gathered ((P - k nb) x nb values), and every rank updates its own
               trailing rows A -= X X_all^T on the tensor cores.  The test Jacobians J are sharded too and ride
               along as extra local rows: the same panel / trailing GEMMs turn them into Y (a bordered
               factorisation), so neither an inverse nor a distributed triangular solve is ever formed.
               The pseudo-inverse of the script is the inverse because H + tau I is positive definite (tau > 0).

Host logic (ownership, packing order, per-step choreography) is exercised on CPU with gloo and world size 2
(tests/test_distributed_cpu.py) with torch stand-ins for the four device operations (`ops=`); on a GPU the
operations are the CUDA library and nothing else.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor

from .distributed import gather_rows, rank, world_size

NB_DEFAULT = 112     # diagonal blocks are factorised by one CTA in fp64 (BK_SMALL64_MAX_DIM)


class CudaDenseOps:
    """The four device operations of the sharded dense Fisher, on libbk_kfac.so."""

    def __init__(self, precision: str = "bf16x3"):
        from . import _lib
        self._lib = _lib
        self.lib = _lib.load()
        _lib.require_device()
        self.precision = precision

    def syrk(self, grads: Tensor, normalise: float) -> Tensor:
        from .dense import dense_fisher
        return dense_fisher(grads, precision=self.precision, normalise=normalise)

    def chol_trinv(self, blk: Tensor, add: float, status: Tensor) -> Tensor:
        nb = blk.shape[0]
        w = torch.empty(nb, nb, device=blk.device, dtype=torch.float32)
        self._lib.check(self.lib.bk_chol_trinv_f64(blk.data_ptr(), blk.stride(0), nb, float(add), w.data_ptr(), nb,
                                                   status.data_ptr(), self._lib.stream_ptr()), "bk_chol_trinv_f64")
        return w

    def _staged(self, slot: int, x: Tensor):
        """bf16 (hi, lo) K-major copies of an fp32 (possibly strided) matrix in a grow-only staging buffer: two
        operands per GEMM, 270 GEMMs per factorisation - no allocation, no zero-fill, no contiguous() copy."""
        _lib = self._lib
        rows, cols = x.shape
        ld = (cols + 7) // 8 * 8
        need = rows * ld
        pool = getattr(self, "_pool", None)
        if pool is None:
            pool = self._pool = {}
        buf = pool.get(slot)
        if buf is None or buf.numel() < 2 * need or buf.device != x.device:
            buf = pool[slot] = torch.zeros(2 * max(need, 1 << 20), dtype=torch.bfloat16, device=x.device)
        if ld != cols:
            buf[:2 * need].zero_()      # padding columns take part in the TMA boxes
        hi, lo = buf.data_ptr(), buf.data_ptr() + 2 * need
        assert x.stride(1) == 1 and x.dtype == torch.float32
        _lib.check(self.lib.bk_convert_split(x.data_ptr(), x.stride(0), rows, cols, 1.0, 0, hi, lo, ld,
                                             _lib.stream_ptr()), "bk_convert_split")
        return hi, lo, ld

    def gemm_nt(self, a: Tensor, b: Tensor, out: Tensor, alpha: float, beta: float) -> None:
        """out = alpha * a b^T + beta * out (fp32 views, split-bf16 tensor-core passes).  `out` may alias `a`:
        the operands are staged copies."""
        _lib = self._lib
        m, k = a.shape
        n = b.shape[0]
        if m == 0 or n == 0:
            return
        a_hi, a_lo, lda = self._staged(0, a)
        b_hi, b_lo, ldb = self._staged(1, b)
        assert out.stride(1) == 1
        _lib.check(self.lib.bk_gemm_nt(a_hi, a_lo, lda, 0, b_hi, b_lo, ldb, 0, m, n, k, 1, _lib.BK_PREC_BF16X3, 0,
                                       float(alpha), float(beta), out.data_ptr(), out.stride(0), 0, 0, 0, 0, 0, 0, 0,
                                       _lib.stream_ptr()), "bk_gemm_nt")

    def rownorm2(self, y: Tensor) -> Tensor:
        _lib = self._lib
        y = y.contiguous()
        out = torch.empty(y.shape[0], device=y.device, dtype=torch.float32)
        if y.shape[0]:
            _lib.check(self.lib.bk_frob_dot(out.data_ptr(), y.data_ptr(), y.stride(0), y.data_ptr(), y.stride(0),
                                            y.shape[1], y.shape[0], 0, 0, _lib.stream_ptr()), "bk_frob_dot")
        return out

    def dominance_rows(self, rows: Tensor, row0: int, P: int, tau: float, coords) -> Tensor:
        _lib = self._lib
        dev = rows.device
        coords = sorted(coords)
        bb = torch.tensor([c[0] for c in coords], dtype=torch.int32, device=dev)
        be = torch.tensor([c[1] for c in coords], dtype=torch.int32, device=dev)
        out = torch.empty(3, dtype=torch.float64, device=dev)
        _lib.check(self.lib.bk_dominance_rows(rows.data_ptr(), rows.stride(0), int(row0), rows.shape[0], P,
                                              float(tau), bb.data_ptr(), be.data_ptr(), len(coords),
                                              out.data_ptr(), _lib.stream_ptr()), "bk_dominance_rows")
        return out


def owned_blocks(nblk: int, world: int, r: int) -> List[int]:
    """Global block rows of rank r (block-cyclic: g % world == r)."""
    return list(range(r, nblk, world))


def _reduce_scatter(send: Tensor, group=None) -> Tensor:
    """send [world, chunk...] -> the summed chunk of this rank.  NCCL: one reduce_scatter_tensor; backends without
    it (gloo, used by the CPU tests of the host logic): all-reduce and slice."""
    w, me = world_size(group), rank(group)
    if w == 1:
        return send[0]
    if send.is_cuda:
        out = torch.empty_like(send[0])
        dist.reduce_scatter_tensor(out, send, op=dist.ReduceOp.SUM, group=group)
        return out
    dist.all_reduce(send, op=dist.ReduceOp.SUM, group=group)
    return send[me].clone()


class ShardedDenseFisher:
    """Row-block-cyclic shard of the dense Fisher H [P, P] (see module docstring).  `rows` holds this rank's block
    rows of the identity-padded H ([n_owned * nb, P_pad], global block g = l * world + rank at local slot l)."""

    def __init__(self, rows: Tensor, P: int, nb: int, group=None, ops=None):
        self.rows, self.P, self.nb, self.group = rows, int(P), int(nb), group
        self.ops = ops if ops is not None else CudaDenseOps()
        self.world, self.rank = world_size(group), rank(group)
        self.nblk = (self.P + nb - 1) // nb
        self.P_pad = self.nblk * nb
        self.mine = owned_blocks(self.nblk, self.world, self.rank)
        assert rows.shape == (len(self.mine) * nb, self.P_pad)

    # ------------------------------------------------------------------------------------------ dominance
    def dominance(self, coords: Sequence[Tuple[int, int]], tau: float = 1e-5) -> Tuple[float, float]:
        """(sum|diag| / sum|all|, sum|kernel blocks| / sum|all|) of H + tau I (hessian/utils.py:4-23)."""
        acc = torch.zeros(3, dtype=torch.float64, device=self.rows.device)
        nb = self.nb
        for l, g in enumerate(self.mine):
            nrows = min(nb, self.P - g * nb)
            if nrows > 0:
                acc += self.ops.dominance_rows(self.rows[l * nb:l * nb + nrows], g * nb, self.P, tau, coords)
        if self.world > 1:
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        s_diag, s_all, s_blk = acc.tolist()
        return s_diag / s_all, s_blk / s_all

    # ------------------------------------------------------------------------------------------- variance
    def variance(self, J_local: Tensor, tau: float, n_rows: Optional[int] = None) -> Tensor:
        """|J_b (H + tau I)^-1 J_b^T| for the test Jacobian rows of ALL ranks ([n_rows], identical everywhere);
        J_local [B_local, P] are this rank's rows (`distributed.row_slice(n_rows, world, rank)` of the test set).
        classification_ll_dense.py:108-109, 160-161."""
        ops, nb, w, me = self.ops, self.nb, self.world, self.rank
        dev = self.rows.device
        nloc, nblk, Pp = len(self.mine), self.nblk, self.P_pad
        Bl = J_local.shape[0]
        A = torch.zeros(nloc * nb + Bl, Pp, device=dev, dtype=self.rows.dtype)
        A[:nloc * nb].copy_(self.rows)
        A[nloc * nb:, :self.P].copy_(J_local)
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        owned = [len(range(r, nblk, w)) for r in range(w)]

        def counts_after(k, r):      # blocks g = r, r + w, ... with k < g < nblk
            return owned[r] - (0 if k < r else (k - r) // w + 1)

        slots_max = (nblk - 1) // w + 1
        send_full = torch.zeros(slots_max * nb, nb, device=dev, dtype=A.dtype) if w > 1 else None
        recv_full = torch.empty(w * slots_max * nb, nb, device=dev, dtype=A.dtype) if w > 1 else None
        for k in range(nblk):
            owner, lk = k % w, k // w
            c0, c1 = k * nb, (k + 1) * nb
            if me == owner:
                W = ops.chol_trinv(A[lk * nb:(lk + 1) * nb, c0:c1], tau, status).contiguous()  # the wire layout
            else:
                W = torch.empty(nb, nb, device=dev, dtype=A.dtype)
            if w > 1:
                dist.broadcast(W, src=dist.get_global_rank(self.group, owner) if self.group is not None else owner,
                               group=self.group)
            l0 = nloc - counts_after(k, me)              # first local block row below block k
            mine_rows = A[l0 * nb:]                        # block rows below k, then the Jacobian rows
            panel = mine_rows[:, c0:c1]
            ops.gemm_nt(panel, W, panel, 1.0, 0.0)        # X = A[:, k] W^T (operands are staged copies: in place)
            if k == nblk - 1:
                break
            if w > 1:
                # all-gather of the panel's block rows.  Every rank places block g = l * w + r at slot l - base
                # (base = the local index of block k + 1 on its owner), so the gathered [rank][slot] buffer read in
                # [slot][rank] order IS the global block order: one strided copy, no index tensors.  The slots in
                # front of block k + 1 and behind the last block hold stale data and are sliced away.
                base = (k + 1) // w
                cm = slots_max - base
                send = send_full[:cm * nb]
                if nloc > l0:
                    send[(l0 - base) * nb:(nloc - base) * nb].copy_(A[l0 * nb:nloc * nb, c0:c1])
                recv = recv_full[:w * cm * nb]
                dist.all_gather_into_tensor(recv, send, group=self.group)
                ordered = recv.view(w, cm, nb * nb).permute(1, 0, 2).reshape(cm * w * nb, nb)
                x_all = ordered[(k + 1 - base * w) * nb:(nblk - base * w) * nb]
            else:
                x_all = A[l0 * nb:nloc * nb, c0:c1]
            # trailing update of my rows (block rows and Jacobian rows): A[:, k+1:] -= X X_all^T
            ops.gemm_nt(panel, x_all, mine_rows[:, c1:], -1.0, 1.0)
        code = int(status.item())
        if self.world > 1:
            t = torch.tensor([code], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
            code = int(t.item())
        if code:
            raise RuntimeError(f"H + tau I is not positive definite (pivot {code} of a diagonal block, tau={tau})")
        var_local = ops.rownorm2(A[nloc * nb:, :self.P])
        if w == 1:
            return var_local
        if n_rows is None:
            raise ValueError("n_rows (the global number of test rows) is needed when world > 1")
        return gather_rows(var_local, n_rows, self.group)


def dense_fisher_sharded(grads_local: Tensor, n_total: int, nb: int = NB_DEFAULT, group=None,
                         ops=None) -> ShardedDenseFisher:
    """H = sum over ALL ranks' gradient rows g g^T / n_total, returned as a row-block-cyclic shard.
    grads_local [n_local, P]: this rank's stacked flat gradients (dense.flat_gradient per batch)."""
    ops = ops if ops is not None else CudaDenseOps()
    w, me = world_size(group), rank(group)
    P = grads_local.shape[1]
    nblk = (P + nb - 1) // nb
    Pp = nblk * nb
    H = ops.syrk(grads_local, float(n_total))                     # [P, P] partial sum of this rank
    nloc_max = (nblk + w - 1) // w
    # send buffer in owner order: chunk r = the block rows r, r + w, ... (identity-padded to P_pad columns/rows)
    send = torch.zeros(w, nloc_max * nb, Pp, device=H.device, dtype=H.dtype)
    full_blocks = P // nb
    for r in range(w):
        blocks = owned_blocks(nblk, w, r)
        whole = [g for g in blocks if g < full_blocks]
        if whole:
            src = H[:full_blocks * nb].view(full_blocks, nb, -1)[r::w]        # strided view of rank r's blocks
            send[r, :len(whole) * nb, :P].copy_(src.reshape(len(whole) * nb, -1)[:, :P])
        if blocks and blocks[-1] == full_blocks and P % nb:
            l = len(blocks) - 1
            send[r, l * nb:l * nb + P % nb, :P].copy_(H[full_blocks * nb:P, :P])
    if P % nb and me == 0:
        # identity padding of the ragged last block (added once: the reduction sums the ranks' buffers)
        g = nblk - 1
        r, l = g % w, g // w
        pad = torch.arange(P, Pp, device=H.device)
        send[r, l * nb + (pad - g * nb), pad] = 1.0
    del H
    mine = _reduce_scatter(send, group)
    nloc = len(owned_blocks(nblk, w, me))
    return ShardedDenseFisher(mine[:nloc * nb].contiguous(), P, nb, group, ops)
