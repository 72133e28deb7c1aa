"""Raw peer-memory bandwidth under torchrun: every rank copies 256 MB between its buffer and a peer's
((rank + 1) % world, or from / to ALL peers in slices), pull and push, 4- and 16-byte accesses, several grid sizes."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, ".")
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from bnn_kfac_b200 import _lib
from bnn_kfac_b200 import distributed as D
L = _lib.load()
NB = 256 << 20
ctx = D.PeerExchange.get(2 * NB, dev)
assert ctx is not None
st = _lib.stream_ptr()


def timed(fn, reps=5):
    ms = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms.append(t.item())
    return min(ms)


peer = (rank + 1) % world
for vec in (4, 16):
    for ctas in (148 * 2, 148 * 8, 148 * 32):
        pull = timed(lambda: L.bk_peer_copy(ctx.data(rank, NB // 4), ctx.data(peer, 0), NB, vec, ctas, st))
        push = timed(lambda: L.bk_peer_copy(ctx.data(peer, NB // 4), ctx.data(rank, 0), NB, vec, ctas, st))
        loc = timed(lambda: L.bk_peer_copy(ctx.data(rank, NB // 4), ctx.data(rank, 0), NB, vec, ctas, st))
        if rank == 0:
            print(f"world={world} one peer, {vec:2d} B accesses, {ctas:5d} CTAs: pull {NB / pull / 1e6:7.1f} GB/s  "
                  f"push {NB / push / 1e6:7.1f} GB/s  (local copy {NB / loc / 1e6:7.1f} GB/s)", flush=True)
if world > 2:
    sl = NB // (world - 1) // 16 * 16
    for vec, ctas in ((16, 148 * 8), (16, 148 * 32)):
        def all_peers(push):
            for k in range(1, world):
                p = (rank + k) % world
                if push:
                    L.bk_peer_copy(ctx.data(p, NB // 4 + rank * (sl // 4)), ctx.data(rank, (k - 1) * (sl // 4)), sl, vec, ctas // (world - 1), st)
                else:
                    L.bk_peer_copy(ctx.data(rank, NB // 4 + (k - 1) * (sl // 4)), ctx.data(p, rank * (sl // 4)), sl, vec, ctas // (world - 1), st)
        # the (world - 1) copies are queued on one stream: sequential, one peer at a time, every rank a different one
        pull = timed(lambda: all_peers(False)); push = timed(lambda: all_peers(True))
        if rank == 0:
            print(f"world={world} all peers in turn (staggered), 16 B, {ctas // (world - 1)} CTAs each: pull "
                  f"{sl * (world - 1) / pull / 1e6:7.1f} GB/s  push {sl * (world - 1) / push / 1e6:7.1f} GB/s", flush=True)
src = torch.empty(NB // 4, device=dev); dst = torch.empty(NB // 4, device=dev)
dist.destroy_process_group()
