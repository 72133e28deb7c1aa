"""GPU tests of the peer-memory factor exchange (csrc/bk_peer.cu, distributed.PeerExchange): the tile-packed layout
and the fused sum + unpack kernel against torch on ONE GPU (sources in local memory), the flag kernels, and - when
two GPUs are visible - the whole exchange under torchrun against the NCCL route (tools/gpu_peer_check.py)."""
import ctypes as C
import os
import subprocess
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from bnn_kfac_b200 import _lib
    _lib.require_device()
    return torch.device("cuda:0")


def _arrays(ts):
    n = len(ts)
    return ((C.c_void_p * n)(*[t.data_ptr() for t in ts]), (C.c_longlong * n)(*[t.stride(0) for t in ts]),
            (C.c_int * n)(*[t.shape[0] for t in ts]))


def test_tile_pack_and_fused_sum_unpack_vs_torch(dev):
    """bk_tile_pack -> bk_peer_tile_unpack over 1, 3 and 8 source buffers (ragged sizes, a pitched factor, explicit
    chunk offsets): out = scale * (sum of the sources' lower triangles), mirrored or with a zero upper triangle;
    bit-exact against the same sum formed in rank order by torch."""
    from bnn_kfac_b200 import _lib
    from bnn_kfac_b200.curvatures import _alloc_factor
    lib = _lib.load()
    st = _lib.stream_ptr()
    g = torch.Generator().manual_seed(3)
    dims = [1, 5, 32, 33, 177, 1025, 10]
    n = len(dims)
    cd = (C.c_int * n)(*dims)
    total = int(lib.bk_tile_packed_floats(cd, n))
    assert total == sum(((d + 31) // 32) * ((d + 31) // 32 + 1) // 2 * 1024 for d in dims)
    for nsrc in (1, 3, 8):
        srcs, packed = [], []
        for s in range(nsrc):
            mats = []
            for d in dims:
                x = torch.randn(d, d, generator=g)
                f = _alloc_factor(d, dev)
                f.copy_((x + x.t()).to(dev))
                mats.append(f)
            buf = torch.full((total + 8,), 7.0, device=dev)
            p, ld, _ = _arrays(mats)
            assert lib.bk_tile_pack(p, ld, cd, None, n, buf.data_ptr(), st) == 0
            assert torch.all(buf[total:] == 7.0)
            srcs.append(mats)
            packed.append(buf)
        # per-factor offsets inside a source buffer (back to back)
        offs, off = [], 0
        for d in dims:
            offs.append(off)
            T = (d + 31) // 32
            off += T * (T + 1) // 2 * 1024
        outs = [torch.full((d, d), -1.0, device=dev) for d in dims]
        po, ldo, _ = _arrays(outs)
        ptrs = (C.c_void_p * (n * nsrc))(*[packed[s].data_ptr() + 4 * offs[k] for k in range(n) for s in range(nsrc)])
        for mirror in (1, 0):
            assert lib.bk_peer_tile_unpack(po, ldo, cd, n, ptrs, nsrc, 0.25, mirror, st) == 0
            for k in range(n):
                acc = srcs[0][k].clone()
                for s in range(1, nsrc):
                    acc = acc + srcs[s][k]
                want = 0.25 * acc
                assert torch.equal(outs[k], want if mirror else torch.tril(want)), (nsrc, mirror, dims[k])
    # explicit offsets on the packing side: two factors swapped inside the buffer
    a, b = srcs[0][4], srcs[0][5]
    ta, tb = [((d + 31) // 32) * ((d + 31) // 32 + 1) // 2 * 1024 for d in (177, 1025)]
    buf = torch.zeros(ta + tb, device=dev)
    p, ld, dd = _arrays([a, b])
    assert lib.bk_tile_pack(p, ld, dd, (C.c_longlong * 2)(tb, 0), 2, buf.data_ptr(), st) == 0
    outs = [torch.empty(177, 177, device=dev), torch.empty(1025, 1025, device=dev)]
    po, ldo, _ = _arrays(outs)
    ptrs = (C.c_void_p * 2)(buf.data_ptr() + 4 * tb, buf.data_ptr())
    assert lib.bk_peer_tile_unpack(po, ldo, dd, 2, ptrs, 1, 1.0, 1, st) == 0
    assert torch.equal(outs[0], a) and torch.equal(outs[1], b)


def test_peer_flags_and_timeout(dev):
    """bk_peer_signal / bk_peer_wait on one device: a wait passes once every slot has reached the epoch and raises the
    error word (1 + the missing rank) instead of hanging when one never does."""
    from bnn_kfac_b200 import _lib
    lib = _lib.load()
    st = _lib.stream_ptr()
    flags = torch.zeros(8, dtype=torch.int32, device=dev)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    world = 3
    for me in range(world):                      # the three "ranks" share one flag array here
        arr = (C.c_void_p * world)(*[flags.data_ptr()] * world)
        assert lib.bk_peer_signal(arr, world, me, 5, st) == 0
    assert lib.bk_peer_wait(flags.data_ptr(), world, 5, C.c_double(1.0), err.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert flags[:3].tolist() == [5, 5, 5] and err.item() == 0
    assert lib.bk_peer_wait(flags.data_ptr(), world, 4, C.c_double(1.0), err.data_ptr(), st) == 0   # older epoch
    flags[1] = 2
    assert lib.bk_peer_wait(flags.data_ptr(), world, 5, C.c_double(0.05), err.data_ptr(), st) == 0
    torch.cuda.synchronize()
    assert err.item() == 2                       # rank 1 never arrived


def test_peer_alloc_export_roundtrip(dev):
    from bnn_kfac_b200 import _lib
    lib = _lib.load()
    p = C.c_void_p()
    assert lib.bk_peer_alloc(1 << 20, C.byref(p)) == 0 and p.value
    h = (C.c_ubyte * 64)()
    assert lib.bk_peer_export(p, h) == 0 and any(bytes(h))
    out = C.c_uint(99)
    assert lib.bk_peer_read_u32(p, C.byref(out)) == 0 and out.value == 0     # zero-filled
    assert lib.bk_peer_free(p) == 0


def test_peer_exchange_two_ranks_vs_nccl():
    """The whole peer-memory exchange (reduce-scatter to owners, return of the Cholesky factors, invert_sharded) on 2
    GPUs against the NCCL route; skipped on a 1-GPU box."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", BK_WIDTHS="300,520,260,10")
    proc = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29733",
                           str(ROOT / "tools" / "gpu_peer_check.py")], capture_output=True, text=True, env=env,
                          timeout=600)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-3000:]
    assert "peer check ok" in proc.stdout
