"""Times distributed.invert_sharded on the cfg5 factor set under torchrun (barrier before every repetition,
max over ranks), with the Cholesky factors returned by ONE all-gather of packed triangles (default) or by one
broadcast per factor (BK_SHARDED_BCAST=1)."""
import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, ".")
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from bnn_kfac_b200.curvatures import KFAC
from bnn_kfac_b200.distributed import invert_sharded, reduce_state_copy
from bnn_kfac_b200.wrapper import MLP
widths = [4096, 4096, 4096, 4096, 10]
model = MLP(widths).to(dev)
est = KFAC(model, precision="bf16")
layers = [l for _, l in est._selected_layers()]
g = torch.Generator().manual_seed(1 + rank)
for l, (i, o) in zip(layers, zip(widths[:-1], widths[1:])):
    est.record[l] = [torch.randn(4096, i, generator=g).to(dev), (torch.randn(4096, o, generator=g) / 4096).to(dev)]
est.update(4096)
def timed(fn, reps=5):
    best = []
    for _ in range(reps):
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX)
        best.append(t.item())
    return min(best), sorted(best)[len(best) // 2]
invert_sharded(est, 1.0, 200.0)
a = timed(lambda: invert_sharded(est, 1.0, 200.0))
b = timed(lambda: reduce_state_copy(est))
if rank == 0:
    mode = "broadcast per factor" if os.environ.get("BK_SHARDED_BCAST") else "one all-gather of packed triangles"
    print(f"world={world} [{mode}]: invert_sharded best {a[0]:.2f} ms / median {a[1]:.2f} ms; exchange alone best {b[0]:.2f} ms", flush=True)
dist.destroy_process_group()
