#!/usr/bin/env python
"""bench.py — KFAC factor-update throughput (+ inversion latency and posterior-predictive
throughput) on BASELINE.json's wide-MLP configuration.

    python bench.py --gpus 1 --steps 20 --warmup 3            # this repo's CUDA path
    python bench.py --impl reference --steps 20 --warmup 3    # the reference algorithm on host cores
    torchrun --nproc-per-node N ... bench.py --gpus N ...     # one rank per GPU (weak scaling)

Workload (config.workload = "cfg5_wide_mlp"): MLP 4096-4096-4096-4096-10, batch 4096 per GPU,
synthetic bf16-representable activations N(0,1) and output gradients N(0,1)/N (BASELINE.md §3,
config 5).  A *step* is one `KFAC.update` over one batch: first and second Kronecker factor of all
four Linear layers (A 4097^2 x4, G 4096^2 x3 + 10^2), accumulated into the running state
(models/curvatures.py:325-365 of the reference).  Model forward/backward is not part of the step
(the metric is "factor-update samples/s": samples whose (a, g) are folded into all factors per second).

  value : samples/s with the step's (a, g) already resident in HBM (device-timed, max over ranks)
  e2e   : same metric through the public API with HOST buffers: per step the (a, g) tensors are copied
          from pinned host memory, `KFAC.update` runs, and a per-factor checksum is read back (the copy of
          step i + 1 is issued on a copy stream before the kernels of step i: every step still moves its own
          0.47 GB, the link and the GPU work concurrently)
  roofline     : the tcgen05 SYRK kernel, algorithmic flops d'(d'+1)N per launch / event-timed launch
  cpu_baseline : oracle/ (CPU restatement of the reference's update) on a bounded sample, rank 0 only

Multi-GPU: the batch axis shards (each rank owns its own 4096-sample batches, weak scaling); the
factors are plain sums (curvatures.py:359-361), so the only exchange is ONE reduction of the factor
states after the K accumulation steps (what `distributed.invert_sharded` needs); it is inside the
timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WIDTHS = [4096, 4096, 4096, 4096, 10]
BATCH = 4096


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0, "source": "fallback"}


def synth_batch(gen, batch, widths):
    """Per layer (a [batch, d_in], g [batch, d_out]) fp32, bf16-representable (BASELINE config 5)."""
    out = []
    for d_in, d_out in zip(widths[:-1], widths[1:]):
        a = torch.randn(batch, d_in, generator=gen).bfloat16().float()
        g = (torch.randn(batch, d_out, generator=gen) / batch).bfloat16().float()
        out.append((a, g))
    return out


def algorithmic_flops_per_sample(widths):
    """d_in'(d_in'+1) + d_out(d_out+1) per sample per Linear layer."""
    return sum((a + 1) * (a + 2) + b * (b + 1) for a, b in zip(widths[:-1], widths[1:]))


class ClockSampler:
    """Samples SM clock / throttle reasons of one GPU during the timed region (pynvml)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40,
                 "sw_thermal_slowdown": 0x20, "hw_power_brake_slowdown": 0x80, "sync_boost": 0x10,
                 "applications_clocks_setting": 0x2}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


# ----------------------------------------------------------------------------------------------------
class _HostMLP(torch.nn.Module):
    """The cfg5 MLP as plain torch (Linear + ReLU), the model the reference estimator is attached to."""

    def __init__(self, widths):
        super().__init__()
        self.layers = torch.nn.ModuleList(torch.nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))

    def forward(self, x):
        for i, l in enumerate(self.layers):
            x = l(x)
            if i + 1 < len(self.layers):
                x = torch.relu(x)
        return x


def reference_stepper(rows):
    """One step of the REFERENCE's own factor update on the host cores: the unmodified `KFAC.update`
    (models/curvatures.py:325-365, staged under oracle/_ref by oracle/make_ref.py) fed the same (a, g) the GPU arm
    gets — `record[layer] = [a, g * N]` is exactly what its hooks store (:319-323).  Falls back to the oracle port
    (same arithmetic, restated) only if oracle/_ref was not staged.  Returns (step_fn, kind, description)."""
    from oracle import make_ref
    gen = torch.Generator().manual_seed(1234)
    data = synth_batch(gen, rows, WIDTHS)
    if make_ref.available():
        ref = make_ref.load()
        torch.manual_seed(0)
        model = _HostMLP(WIDTHS)
        est = ref.KFAC(model)
        layers = list(model.layers)
        recs = [[a, g * a.shape[0]] for a, g in data]

        def step():
            for layer, rec in zip(layers, recs):
                est.record[layer] = rec
            est.update(rows)
        return step, "reference", "unmodified reference KFAC.update (oracle/_ref/models/curvatures.py), fp32 torch CPU"
    from oracle import kfac_oracle as O
    state = [None] * len(data)

    def step_port():
        for i, (a, g) in enumerate(data):
            f1, f2 = O.kfac_linear_factors(a, g * a.shape[0], True)
            if state[i] is None:
                state[i] = [f1, f2]
            else:  # models/curvatures.py:359-361
                state[i][0] += f1
                state[i][1] += f2
    return step_port, "port", "oracle port of KFAC.update (oracle/_ref not staged), fp32 torch CPU"


def bench_config(world):
    """The `config` object shared by both arms (the driver compares them key by key)."""
    return {"workload": "cfg5_wide_mlp", "widths": WIDTHS, "batch_per_gpu": BATCH,
            "parallelism": f"batch-sharded x{world}, one factor exchange per timed region",
            "l2": "inputs_larger_than_l2 (0.47 GB of activations + 0.4 GB of factor state per step)"}


def run_reference(args):
    """`--impl reference`: the reference's own CPU implementation at the FULL batch (same config as the GPU arm),
    all host threads, median of the timed steps."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, what = reference_stepper(BATCH)
    for _ in range(max(args.warmup, 1)):
        step()
    times = []
    for _ in range(max(args.steps, 1)):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    value = BATCH / med
    sample = f"{len(times)} steps of the full {BATCH}-row batch, all 4 layers; median step; {what}"
    line = {"impl": "reference", "metric": "kfac_factor_update_samples_per_s", "value": value,
            "unit": "samples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": med * 1e3, "ms_per_step_min_max": [times[0] * 1e3, times[-1] * 1e3],
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": bench_config(max(args.gpus, 1)),
            "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind,
                             "sample": sample},
            "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)
    return 0


def cpu_baseline_leg(max_seconds=20.0):
    """Reported beside the GPU line (rank 0, N = 1): the same reference arm on a bounded number of full-batch
    steps (1 warm-up + up to 5 timed within ~20 s)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind, what = reference_stepper(BATCH)
    step()
    times, t_start = [], time.perf_counter()
    while len(times) < 5 and (not times or time.perf_counter() - t_start < max_seconds):
        t0 = time.perf_counter()
        step()
        times.append(time.perf_counter() - t0)
    times.sort()
    med = times[len(times) // 2]
    return {"value": BATCH / med, "unit": "samples/s", "cores": cores, "kind": kind,
            "sample": f"{len(times)} steps of the full {BATCH}-row batch, all 4 layers; median step; {what}"}


def cpu_predictive_leg(n_inputs=256):
    """Oracle port of ONE posterior weight sample of every layer (curvatures.py:400-405: z, L_A z L_G^T, transpose;
    :67-82 _replace) + one forward pass (wrapper.py:35-44) at cfg5 on the host cores.  The Cholesky factors are
    random lower-triangular matrices: the arithmetic (two d^3 products per layer and sample) does not depend on
    their values, and inverting 4097-wide factors on the CPU first would take longer than the whole bench."""
    from oracle import kfac_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    gen = torch.Generator().manual_seed(4321)
    dims = list(zip(WIDTHS[:-1], WIDTHS[1:]))
    chol = [(torch.tril(torch.randn(i + 1, i + 1, generator=gen)) * 1e-2,
             torch.tril(torch.randn(o, o, generator=gen)) * 1e-2) for i, o in dims]
    weights = [(torch.randn(o, i, generator=gen) / i ** 0.5, torch.zeros(o)) for i, o in dims]
    x = torch.randn(n_inputs, WIDTHS[0], generator=gen)
    t0 = time.perf_counter()
    h = x
    for li, ((la, lg), (w, b)) in enumerate(zip(chol, weights)):
        z = torch.randn(la.shape[0], lg.shape[0], generator=gen)
        w_s, b_s = O.replace(O.kfac_sample(la, lg, z), w, b)
        h = torch.nn.functional.linear(h, w_s, b_s)
        if li + 1 < len(chol):
            h = torch.relu(h)
    torch.softmax(h, 1)
    dt = time.perf_counter() - t0
    return {"value": n_inputs / dt, "unit": "(weight samples x test inputs)/s", "weight_samples_per_s": 1.0 / dt,
            "cores": cores, "kind": "port",
            "sample": f"1 weight sample of all 4 layers + forward of {n_inputs} inputs, fp32 torch CPU "
                      f"(per-sample cost is independent of the input count up to the forward GEMMs)"}


# ----------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip inversion / predictive extras")
    ap.add_argument("--no-sustained", action="store_true", help="skip the >= 2 s back-to-back leg")
    ap.add_argument("--inputs", default="fp32", choices=["fp32", "bf16"],
                    help="dtype of the (a, g) tensors handed to KFAC.update: fp32 (what the reference's hooks see on an "
                         "fp32 model; staged to bf16 by the library) or bf16 (a model under bf16 autocast; consumed "
                         "directly).  BASELINE config 5's values are bf16-representable: both carry the same numbers")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from bnn_kfac_b200 import _lib
    from bnn_kfac_b200.curvatures import KFAC
    from bnn_kfac_b200.distributed import (bind_to_gpu_numa_node, plan_owners, reduce_scatter_to_owners,
                                           reduce_state_copy)
    # N > 1: each rank runs on (and allocates its pinned e2e buffers from) the NUMA node next to its GPU
    numa_cores = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    from bnn_kfac_b200.wrapper import MLP
    L = _lib.load()
    _lib.require_device()

    # N > 1: the sharded path must reproduce the single-GPU result before anything is timed (same check as
    # tests/test_gpu_parity_r2.py::test_nccl_two_rank_parity; its outcome travels in the JSON line)
    parity = None
    if world > 1:
        sys.path.insert(0, os.path.join(ROOT, "tools"))
        import gpu_dist_check
        parity = gpu_dist_check.run_check(rank, world, dev)
        # SURVEY 8(e) rows 4-6 (linearised predictive, Diagonal, dense Fisher at BASELINE config 3's P = 15 080)
        # under NCCL: parity against one GPU + a device-timed figure each
        sharded_rows = gpu_dist_check.run_check_rows(rank, world, dev, dense_p=15080)
        parity["ok"] = parity["ok"] and sharded_rows["ok"]

    torch.manual_seed(0)
    model = MLP(WIDTHS).to(dev)
    est = KFAC(model, precision=args.precision)
    layers = [l for _, l in est._selected_layers()]
    synth = synth_batch(torch.Generator().manual_seed(1234 + rank), BATCH, WIDTHS)

    def make_legs(in_dtype):
        """Device-resident and end-to-end step closures for (a, g) tensors of one container dtype."""
        host = [(a.to(in_dtype).pin_memory(), g.to(in_dtype).pin_memory()) for a, g in synth]
        resident = [(a.to(dev), g.to(dev)) for a, g in host]
        # e2e: two sets of device input buffers; the H2D copy of step i + 1 (copy stream) overlaps the kernels of step i
        staging2 = [[(torch.empty_like(a), torch.empty_like(g)) for a, g in resident] for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=dev)
        copied = [torch.cuda.Event(), torch.cuda.Event()]
        consumed = [torch.cuda.Event(), torch.cuda.Event()]
        e2e_state = {"next": 0, "primed": False}
        h2d_bytes = sum(a.numel() * a.element_size() + g.numel() * g.element_size() for a, g in host)
        checksum_host = torch.empty(2 * len(layers), dtype=torch.float32).pin_memory()

        def step_device(bufs):
            for layer, (a, g) in zip(layers, bufs):
                est.record[layer] = [a, g]
            est.update(BATCH)

        def enqueue_h2d(slot):
            """this step's inputs: pinned host buffers -> device set `slot`, on the copy stream"""
            copy_stream.wait_event(consumed[slot])            # the kernels that last read this set are done
            with torch.cuda.stream(copy_stream):
                for (ha, hg), (da, dg) in zip(host, staging2[slot]):
                    da.copy_(ha, non_blocking=True)
                    dg.copy_(hg, non_blocking=True)
                copied[slot].record(copy_stream)

        def step_e2e():
            """One end-to-end step: H2D of the step's host inputs, KFAC.update, D2H of a checksum of the result.
            Every step copies its own inputs; the copy of the NEXT step is issued before this step's kernels so
            that the two overlap (software pipelining over the timed steps; the first call primes the pipeline)."""
            cur = e2e_state["next"]
            if not e2e_state["primed"]:
                enqueue_h2d(cur)
                e2e_state["primed"] = True
            enqueue_h2d(cur ^ 1)                               # prefetch the next step's inputs
            torch.cuda.current_stream().wait_event(copied[cur])
            step_device(staging2[cur])
            consumed[cur].record(torch.cuda.current_stream())
            # checksum of the accumulators: trace of every factor (the diagonal is valid in the lower-only
            # accumulators, so this read does not trigger the mirror pass a public `state` read would run)
            cs = torch.stack([est._raw(l)[k].diagonal().sum() for l in layers for k in range(2)])
            checksum_host.copy_(cs, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            e2e_state["next"] = cur ^ 1
            return checksum_host

        return {"resident": resident, "step_device": step_device, "step_e2e": step_e2e, "h2d_bytes": h2d_bytes,
                "checksum_host": checksum_host}

    legs = make_legs(torch.bfloat16 if args.inputs == "bf16" else torch.float32)
    resident, step_device, step_e2e = legs["resident"], legs["step_device"], legs["step_e2e"]
    h2d_bytes, checksum_host = legs["h2d_bytes"], legs["checksum_host"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, after=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        if after is not None:
            after()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # the exchange invert_sharded() performs: one all-reduce of (copies of) the accumulated factors
    # (a reduce-scatter of the packed lower triangles to the owning ranks, as invert_sharded issues it)
    def exchange():
        owners = plan_owners([t.shape[0] for l in layers for t in est._raw(l)], world)
        return reduce_scatter_to_owners(est, owners)
    reduce_after = exchange if world > 1 else None
    for _ in range(args.warmup):
        step_device(resident)
    if reduce_after is not None:
        reduce_after()          # warm the NCCL channels: the timed region holds exactly one reduction
    with ClockSampler(local_rank) as clocks:
        ms_dev = timed(lambda: step_device(resident), args.steps, reduce_after)
    reduce_ms = scatter_ms = None
    if world > 1:
        def best_of3(fn):
            best = None
            for _ in range(3):
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                fn()
                e1.record()
                barrier()
                t = torch.tensor([e0.elapsed_time(e1)], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                best = t.item() if best is None else min(best, t.item())
            return best
        reduce_state_copy(est)
        reduce_ms = best_of3(lambda: reduce_state_copy(est))
        scatter_ms = best_of3(exchange)
        # a denser exchange schedule: 2 steps per exchange (the headline region holds `steps` steps + ONE exchange)
        def two_steps_and_exchange():
            step_device(resident)
            step_device(resident)
            exchange()
        two_steps_and_exchange()
        every2_ms = timed(two_steps_and_exchange, 10) / 10.0
        # the same exchange through NCCL (pack -> reduce_scatter_tensor -> unpack), for comparison
        os.environ["BK_NO_PEER"] = "1"
        exchange()
        scatter_nccl_ms = best_of3(exchange)
        os.environ.pop("BK_NO_PEER", None)

    for _ in range(args.warmup):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps, reduce_after)

    total_samples = BATCH * args.steps * world
    value = total_samples / (ms_dev * 1e-3)
    e2e_value = total_samples / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel: the grouped tcgen05 SYRK launch of one step (all 7 wide
    # factors: 4 x A on the 4096 x 4096 block + 3 x G; the bias row / column of A comes from column
    # sums in the staging kernel), timed on its own launches with the operand shapes and state
    # pitches the step uses
    import ctypes as C
    peaks = load_peaks()
    prec = {"bf16": 1, "bf16x3": 3, "fp32": 1}[args.precision]
    n = BATCH
    grp = []      # (state, ld, hi, lo, d)
    for layer, (a, g) in zip(layers, resident):
        for x, st_t in ((a, est.state[layer][0]), (g, est.state[layer][1])):
            d = x.shape[1]
            if d <= 176:
                continue
            if args.inputs == "bf16":      # the launch reads the activations themselves
                hi, lo = x, x
            else:
                hi = torch.empty(d, n, dtype=torch.bfloat16, device=dev)
                lo = torch.empty_like(hi)
                L.bk_transpose_split(x.data_ptr(), d, n, d, 1.0, 0, hi.data_ptr(), lo.data_ptr(), n,
                                     _lib.stream_ptr())
            scratch = torch.zeros(st_t.shape[0], st_t.stride(0), device=dev)
            grp.append((scratch, st_t.stride(0), hi, lo, d))
    cnt = len(grp)
    a_states = (C.c_void_p * cnt)(*[t[0].data_ptr() for t in grp])
    a_lds = (C.c_longlong * cnt)(*[t[1] for t in grp])
    a_hi = (C.c_void_p * cnt)(*[t[2].data_ptr() for t in grp])
    a_lo = (C.c_void_p * cnt)(*[t[3].data_ptr() for t in grp])
    a_ldt = (C.c_longlong * cnt)(*[t[2].stride(0) for t in grp])
    a_ns = (C.c_int * cnt)(*[n] * cnt)
    a_ds = (C.c_int * cnt)(*[t[4] for t in grp])
    a_al = (C.c_float * cnt)(*[1.0 / n] * cnt)
    a_be = (C.c_float * cnt)(*[1.0] * cnt)
    reps = 20
    syrk_flags = (_lib.SYRK_LOWER_ONLY if est.lower_only else 0) | (_lib.SYRK_ROW_MAJOR if args.inputs == "bf16" else 0)
    if args.inputs == "bf16":
        prec = 1

    def syrk():
        _lib.check(L.bk_syrk_accum_staged_grouped(a_states, a_lds, a_hi, a_lo, a_ldt, a_ns, a_ds, a_al, a_be,
                                                  cnt, prec, syrk_flags, _lib.stream_ptr()),
                   "bk_syrk_accum_staged_grouped")
    for _ in range(3):
        syrk()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        syrk()
    e1.record()
    torch.cuda.synchronize()
    syrk_ms = e0.elapsed_time(e1) / reps
    syrk_flops = sum(t[4] * (t[4] + 1) * n for t in grp)   # one multiply-add per lower-triangle entry per sample
    achieved = syrk_flops / (syrk_ms * 1e-3) / 1e12
    roofline = {"bound": "tensor",
                "kernel": f"umma_syrk_grouped_kernel (cta_group::2, TMA-reduce epilogue; {cnt} SYRKs 4096x4096x4096, "
                          f"{'lower triangle only' if est.lower_only else 'lower + mirror'}, in one launch)",
                "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "peak_source": peaks["source"] + " (burst)",
                "us_per_launch": syrk_ms * 1e3, "flops_per_launch": syrk_flops, "traffic": None}
    prof = os.path.join(ROOT, "profiles", "syrk_traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get("dram_bytes_per_launch")
        except Exception:
            pass

    line = {"metric": "kfac_factor_update_samples_per_s", "value": value, "unit": "samples/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "bf16x3": "bf16x3", "fp32": "f32"}[args.precision],
            "data": "synthetic", "inputs": args.inputs,
            "config": bench_config(world),
            "algorithmic_tflops": value * algorithmic_flops_per_sample(WIDTHS) / 1e12,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": checksum_host.numel() * 4, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": None, "roofline": roofline, "clocks": clocks.summary()}
    if numa_cores is not None:
        line["e2e"]["host_cores_bound"] = len(numa_cores)
    line["e2e"]["d2h_note"] = ("the result of a step is the device-resident factor state (0.47 GB, consumed on the device "
                               "by invert()); the per-step read-back is a 32-byte checksum (trace of every factor)")
    line["step_frac_of_burst_peak"] = line["algorithmic_tflops"] / world / peaks["bf16_tflops"]
    if parity is not None:
        line["parity"] = parity
        line["sharded_rows"] = sharded_rows
    if reduce_ms is not None:
        state_bytes = sum(t.numel() * 4 for l in layers for t in est.state[l])
        wire_bytes = sum(t.shape[0] * (t.shape[0] + 1) // 2 * 4 for l in layers for t in est.state[l])
        line["factor_allreduce"] = {
            "ms": reduce_ms, "state_bytes": state_bytes, "wire_bytes": wire_bytes,
            "algbw_GBps": wire_bytes / (reduce_ms * 1e-3) / 1e9,
            "note": "the replicated variant (callers that read .state afterwards): bk_tri_pack -> ONE NCCL "
                    "all-reduce of the packed lower triangles -> bk_tri_unpack (mirror, 1/world); NOT in the timed region"}
        sent = wire_bytes * (world - 1) / world
        from bnn_kfac_b200.distributed import PeerExchange
        peer = PeerExchange.get(0, dev) is not None
        line["factor_exchange"] = {
            "ms": scatter_ms, "wire_bytes": wire_bytes, "bytes_sent_per_rank": sent,
            "busbw_GBps": sent / (scatter_ms * 1e-3) / 1e9,
            "route": "peer memory" if peer else "nccl",
            "nccl_route_ms": scatter_nccl_ms,
            "exchange_every_2_steps": {"ms_per_2_steps_and_exchange": every2_ms,
                                       "samples_per_s": 2 * BATCH * world / (every2_ms * 1e-3),
                                       "note": "10 x (2 steps + 1 exchange), device-timed, max over ranks; divide by "
                                               "world x the N = 1 value for the efficiency at this schedule"},
            "note": "one per timed region (deferred: state is a plain sum of batch means); what invert_sharded issues. "
                    "peer memory: bk_tile_pack into the rank's CUDA-IPC buffer -> flags -> ONE bk_peer_tile_unpack launch "
                    "on every owner that pulls its chunk from all ranks over NVLink, adds in rank order and writes the "
                    "mirrored dense factors (pack and flags included in ms).  nccl route: packed lower triangles in owner "
                    "order -> reduce_scatter_tensor -> bk_tri_unpack"}

    # kernels launched by this library inside the device-timed region (counted by the library itself)
    c0 = L.bk_launch_count()
    step_device(resident)
    torch.cuda.synchronize()
    line["gpu_launches"] = int(L.bk_launch_count() - c0) * args.steps

    if not args.no_extras:
        line["extras"] = extras(est, model, layers, dev, world, rank)
    # ---- the same workload with bf16 (a, g) containers (a model under bf16 autocast; BASELINE config 5's values are
    # bf16-representable, so the numbers are identical): the tensor cores read the activations as they are (no
    # staging pass) and an end-to-end step moves half the bytes over the host link
    if args.inputs == "fp32" and not args.no_extras:
        alt = make_legs(torch.bfloat16)
        for _ in range(args.warmup):
            alt["step_device"](alt["resident"])
        ms_alt = timed(lambda: alt["step_device"](alt["resident"]), args.steps)
        for _ in range(args.warmup):
            alt["step_e2e"]()
        ms_alt_e2e = timed(alt["step_e2e"], args.steps)
        line.setdefault("extras", {})["bf16_inputs"] = {
            "value": total_samples / (ms_alt * 1e-3), "ms_per_step": ms_alt / args.steps,
            "e2e": {"value": total_samples / (ms_alt_e2e * 1e-3), "ms_per_step": ms_alt_e2e / args.steps,
                    "h2d_bytes_per_step": alt["h2d_bytes"]},
            "unit": "samples/s",
            "note": "(a, g) handed to KFAC.update as bf16 tensors: direct MN-major tcgen05 operands, no staging pass; "
                    "no factor exchange in this timed region"}
        del alt
    # ---- sustained legs LAST: seconds of full load leave the part power-capped (SM clock ~1.5 GHz) for a while,
    # which would colour every figure measured after them (the predictive extras read 35 % low in r02a)
    # ---- sustained leg: the same device-resident step back to back for >= 2 s (the headline's timed region is
    # ~10 ms, a burst-clock figure; under seconds of load the part sits at its 1 kW power cap and the SM clock
    # settles near 1.3 - 1.5 GHz).  Own clock record; compared with the SUSTAINED cuBLAS peak.
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(2.2 / (ms_dev / args.steps * 1e-3)))
        with ClockSampler(local_rank) as clocks_sus:
            ms_sus = timed(lambda: step_device(resident), n_sus)
        sustained = {"seconds": ms_sus * 1e-3, "steps": n_sus, "ms_per_step": ms_sus / n_sus,
                     "value": BATCH * n_sus * world / (ms_sus * 1e-3), "unit": "samples/s",
                     "clocks": clocks_sus.summary()}

    syrk_sus = None
    if not args.no_sustained:      # the kernel alone, back to back for >= 2 s
        n_k = int(2.2 / (syrk_ms * 1e-3))
        e0.record()
        for _ in range(n_k):
            syrk()
        e1.record()
        torch.cuda.synchronize()
        syrk_sus = syrk_flops / (e0.elapsed_time(e1) / n_k * 1e-3) / 1e12
    if syrk_sus is not None and peaks.get("bf16_tflops_sustained"):
        line["roofline"]["sustained"] = {"achieved": syrk_sus, "peak": peaks["bf16_tflops_sustained"],
                                 "frac": syrk_sus / peaks["bf16_tflops_sustained"],
                                 "note": "same launch back to back for >= 2 s, vs the sustained cuBLAS figure"}
    if sustained is not None:
        sustained["algorithmic_tflops"] = sustained["value"] * algorithmic_flops_per_sample(WIDTHS) / 1e12
        if peaks.get("bf16_tflops_sustained"):
            sustained["frac_of_sustained_peak"] = sustained["algorithmic_tflops"] / world / peaks["bf16_tflops_sustained"]
        line["sustained"] = sustained
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg()
        if "extras" in line and "posterior_predictive" in line["extras"]:
            line["extras"]["posterior_predictive"]["cpu_baseline"] = cpu_predictive_leg()
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def extras(est, model, layers, dev, world, rank):
    """Inversion latency and posterior-predictive throughput on the same workload (device-timed)."""
    import torch.distributed as dist
    from bnn_kfac_b200.predictive import mc_moments
    out = {}

    def ev_ms(fn, reps=1):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # second figure of SURVEY 8(d): the update INCLUDING the model's own forward / backward (torch fp32)
    xb = torch.randn(BATCH, WIDTHS[0], device=dev)
    yb = torch.randint(0, WIDTHS[-1], (BATCH,), device=dev)

    def full_step():
        loss = torch.nn.functional.cross_entropy(model(xb), yb)
        model.zero_grad()
        loss.backward()
        est.update(BATCH)
    for _ in range(2):
        full_step()
    ms_full = ev_ms(full_step, reps=5)
    out["update_incl_model_fwd_bwd"] = {"ms_per_step": ms_full, "samples_per_s": BATCH * world / (ms_full * 1e-3),
                                        "note": "torch fp32 forward + backward of the MLP, hooks, then KFAC.update"}
    if world > 1:
        from bnn_kfac_b200.distributed import invert_sharded
        inv = lambda: invert_sharded(est, 1.0, 200.0)     # noqa: E731
        cfg = f"8 factors sharded over {world} ranks: reduce-scatter to owners over peer memory (NCCL when unavailable), owner inverts, Cholesky factors pulled back over peer memory"
    else:
        inv = lambda: est.invert(1.0, 200.0)              # noqa: E731
        cfg = "one batched launch sequence"
    inv()      # warm-up (workspace allocation)
    if world > 1:
        dist.barrier()     # all ranks enter the collective phase together: the figure is not rank skew
    # ~2000 dependent launches in ~10 ms: this phase is as fast as the host thread can issue them, and the box is
    # shared (another tenant's CPU leg shows up as a 2-8x slower sample) -> best of 3
    t_inv = torch.tensor([min(ev_ms(inv) for _ in range(3))], device=dev)
    if world > 1:
        dist.all_reduce(t_inv, op=dist.ReduceOp.MAX)
    out["invert_ms_all_layers"] = t_inv.item()
    out["invert_config"] = "8 factors (4 x 4097^2, 3 x 4096^2, 10^2), add=1, multiply=200; best of 3; " + cfg
    # posterior predictive: S weight samples per rank (sample ids sharded over ranks), 1024 test inputs
    S, B = 16, 1024
    x = torch.randn(B, WIDTHS[0], device=dev)
    for _ in range(2):
        mc_moments(est, x, S, sample0=rank * S)
    # best of 7 x (3 repetitions): 19 launches + allocator calls per 5.6 ms call - a host thread that loses its core for
    # a few ms (shared box) shows up directly; all samples are kept in the line
    with ClockSampler(dev.index or 0) as pclk:
        samples_ms = [ev_ms(lambda: mc_moments(est, x, S, sample0=rank * S), reps=3) for _ in range(7)]
        ms = min(samples_ms)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out["posterior_predictive"] = {"value": S * world * B / (t.item() * 1e-3),
                                   "unit": "(weight samples x test inputs)/s",
                                   "weight_samples_per_s": S * world / (t.item() * 1e-3),
                                   "config": {"samples_per_gpu": S, "test_inputs": B},
                                   "ms_per_call_samples": samples_ms, "clocks": pclk.summary()}
    return out


if __name__ == "__main__":
    sys.exit(main())
