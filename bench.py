#!/usr/bin/env python
"""Benchmark of the hot path: JopDBlock mul! + adjoint on device-resident DBArrays.

    python bench.py --gpus N --steps K --warmup W            # this backend (CUDA, sm_100a)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU algorithm (oracle port)

A "step" is one forward apply ``mul!(d, A, m)`` plus one adjoint apply ``mul!(m, A', d)`` of the
workload BASELINE.json's metric is quoted on: configs[1], the block-diagonal operator of 64 dense
Float32 4096x4096 blocks (strong scaling: the 64 blocks are sharded over the N ranks, grid [N,1],
isdiag, no data-path collective).  Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement".
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "JopDBlock mul!+adjoint effective HBM GB/s"
NBLK, BM, BN = 64, 4096, 4096
DTYPE = np.float32
SEED = 20242


def block_payload(i: int, m: int = BM, n: int = BN, dtype=DTYPE) -> np.ndarray:
    """Block (i,i), U[0,1), identical on every arm (seeded per block)."""
    rng = np.random.default_rng([SEED, i])
    return np.asfortranarray(rng.random((n, m), dtype=dtype).T)


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the GPU is under load."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,utilization.gpu,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.device)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None
            return
        threading.Thread(target=self._read, daemon=True).start()

    def _read(self):
        for line in self.proc.stdout:
            f = [v.strip() for v in line.split(",")]
            if len(f) >= 9:
                self.rows.append(f)

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        rows = self.rows
        def num(v):
            try:
                return float(v)
            except ValueError:
                return None
        loaded = [r for r in rows if (num(r[4]) or 0) >= 50] or rows
        sm = sorted(v for v in (num(r[1]) for r in loaded) if v is not None)
        reasons = []
        for k, name in ((5, "hw_slowdown"), (6, "hw_thermal_slowdown"), (7, "sw_thermal_slowdown"), (8, "sw_power_cap")):
            if any(r[k].lower().startswith("active") for r in loaded):
                reasons.append(name)
        pw = [v for v in (num(r[3]) for r in loaded) if v is not None]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": num(rows[0][2]) if rows else None,
                "reasons": reasons, "samples": len(rows), "samples_under_load": len(loaded),
                "power_w_max": max(pw) if pw else None}


# ------------------------------------------------------------------------------------------------------
# reference arm: the restated reference (oracle port) on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_run(nblocks_sample: int, steps: int, warmup: int):
    """Block-diagonal forward + adjoint as the reference runs it (src:656-661, 703-708): W workers, one
    core each with single-threaded BLAS (what `addprocs` workers are), every worker applying its own
    blocks with per-block gemv (`A*m`, `A'*d`: test/runtests.jl:30-36).  Uses the oracle's block
    operator classes; this is the only place outside tests/ that executes oracle/ code."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import djets_oracle as O
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:                                  # pragma: no cover
        threadpool_limits = None
    cores = max(1, min(os.cpu_count() or 1, nblocks_sample))
    blocks = [O.ODense(block_payload(i)) for i in range(1, nblocks_sample + 1)]
    rng = np.random.default_rng(SEED + 1)
    xs = [rng.random(BN, dtype=DTYPE) for _ in blocks]
    cuts = np.linspace(0, nblocks_sample, cores + 1).astype(int)

    def worker(w):
        for b in range(cuts[w], cuts[w + 1]):            # a worker's local block rows, sequentially
            d = blocks[b].mul(xs[b])                     # mul!(getblock(d,i), A[i,i], getblock(m,i))
            blocks[b].mul_adjoint(d)                     # mul!(getblock(m,i), A[i,i]', getblock(d,i))

    def step(pool):
        list(pool.map(worker, range(cores)))

    limiter = threadpool_limits(limits=1) if threadpool_limits else None
    with ThreadPoolExecutor(cores) as pool:
        for _ in range(warmup):
            step(pool)
        t0 = time.perf_counter()
        for _ in range(steps):
            step(pool)
        dt = (time.perf_counter() - t0) / steps
    if limiter:
        limiter.restore_original_limits() if hasattr(limiter, "restore_original_limits") else None
    nbytes = 2 * nblocks_sample * (BM * BN + BM + BN) * np.dtype(DTYPE).itemsize
    return nbytes / dt / 1e9, dt, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    sample = max(16, min(NBLK, os.cpu_count() or 1))     # at least one block per host core
    steps = max(1, min(args.steps, 10))
    gbs, dt, cores = cpu_reference_run(sample, steps, max(1, min(args.warmup, 2)))
    sample_txt = (f"{sample} of the {NBLK} dense Float32 {BM}x{BN} diagonal blocks, forward+adjoint, "
                  f"{cores} worker threads x 1 BLAS thread, {steps} steps")
    line = {
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.gpus),
        "host_cores": os.cpu_count(),
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample_txt},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "restated reference (NumPy/OpenBLAS port of src/DistributedJets.jl:656-661,703-708), not Julia: "
                "Julia is absent from the image; omits Distributed.jl RPC cost, so it flatters the reference",
    }
    print(json.dumps(line), file=args.out, flush=True)
    return 0


def workload_config(ngpus: int, nblk: int = NBLK) -> dict:
    return {"workload": f"BASELINE configs[1]: block-diagonal JopDBlock, {nblk} dense Float32 {BM}x{BN} blocks, "
                        f"isdiag, grid [{ngpus},1], forward mul! + adjoint mul! per step",
            "blocks": nblk, "block_shape": [BM, BN], "grid": [ngpus, 1],
            "l2": "inputs larger than L2 (operator payload per GPU >= 512 MiB vs 126 MB L2), no flush needed"}


# ------------------------------------------------------------------------------------------------------
# this backend
# ------------------------------------------------------------------------------------------------------
def run_cuda(args):
    import djets_loader
    dj = djets_loader.load()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        box = [dj.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        ctx = dj.Context.spmd(rank, world, local_rank, box[0])
    elif args.single_process and args.gpus > 1:
        world = args.gpus
        ctx = dj.Context(world=world, devices=list(range(world)))
    else:
        ctx = dj.Context(world=1, devices=[0])
    dj.set_default_context(ctx)
    for name in ("fwd_tx", "fwd_tc", "adj_ru", "adj_tc", "ld_policy", "pdl", "graphs", "p2p", "auto_tile"):
        v = getattr(args, name)
        if v is not None:
            ctx.set_param(name, v)

    def barrier():
        ctx.sync()
        if dist is not None:
            dist.barrier()
            import torch
            torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v: float) -> float:
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nblk = args.blocks * (world if args.scaling == "weak" else 1)
    shapes = ([(BM,)] * nblk, [(BN,)] * nblk)
    t_build = time.perf_counter()
    A = dj.JopDBlock(lambda i, j: dj.JopDense(block_payload(i)) if i == j else dj.JopZeroBlock(DTYPE, BM, BN),
                     dims=(nblk, nblk), pids=list(range(world)), dist=[world, 1], isdiag=True, shapes=shapes,
                     dtype=DTYPE, ctx=ctx)
    t_build = time.perf_counter() - t_build
    m, d, m2 = A.domain().zeros(), A.range().zeros(), A.domain().zeros()
    rng = np.random.default_rng(SEED + 1)
    hx = dj.PinnedBuffer(len(m), DTYPE)
    hout = dj.PinnedBuffer(len(m), DTYPE)
    hx.array[:] = rng.random(len(m), dtype=DTYPE)
    m.scatter(hx.array)
    bytes_local = 2 * A.algorithmic_bytes()              # forward + adjoint on the ranks this process drives
    bytes_total = sum_over_ranks(float(bytes_local))

    def step():
        A.mul(d, m)
        A.mul_adjoint(m2, d)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    t_load0 = time.perf_counter()
    for _ in range(args.warmup):
        step()
    barrier()
    ctx.reset_stats()
    ctx.timer_start()
    for _ in range(args.steps):
        step()
    ms = ctx.timer_stop()
    launches = sum(v["launches"] for v in ctx.kernel_stats().values())
    barrier()
    ms = max_over_ranks(ms)
    ms_per_step = ms / args.steps
    value = bytes_total / (ms_per_step * 1e-3) / 1e9

    # per-kernel pass: every launch bracketed by CUDA events on its own stream
    ctx.reset_stats()
    ctx.kernel_timing(True)
    for _ in range(args.steps):
        step()
    ctx.kernel_timing(False)
    kstats = ctx.kernel_stats()
    barrier()

    # end to end through the public API with pinned host buffers: setblock!/scatter of m (H2D),
    # mul!, adjoint mul!, collect of the result (D2H), every step
    def e2e_step():
        m.scatter(hx.array)
        A.mul(d, m)
        A.mul_adjoint(m2, d)
        m2.to_array(hout.array)

    e2e_steps = 0 if args.profile else args.steps
    for _ in range(0 if args.profile else max(3, args.warmup)):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = max_over_ranks((time.perf_counter() - t0) / max(e2e_steps, 1))
    local_elems = sum(len(r) for r, p in zip(m.indices, m.pids) if ctx.is_local(p))
    h2d = int(sum_over_ranks(float(local_elems * 4)))
    d2h = h2d

    # keep the GPU busy long enough for nvidia-smi to have seen it under load
    if rank == 0 and not args.profile:
        while time.perf_counter() - t_load0 < 1.5:
            for _ in range(50):
                step()
            ctx.sync()
    barrier()
    clocks = sampler.stop() if rank == 0 else None

    # sanity spot check of the timed configuration on rank 0's first block against a Float64 product
    # computed here with NumPy (the parity tests proper live in tests/)
    parity = None
    if rank == 0:
        x1 = hx.array[:BN].astype(np.float64)
        want = block_payload(1).astype(np.float64) @ x1
        A.mul(d, m)
        got = d.getblock(1).astype(np.float64)
        parity = float(np.linalg.norm(got - want) / np.linalg.norm(want))

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.profile:
        sample = 8
        gbs, dt, cores = cpu_reference_run(sample, 3, 1)
        cpu = {"value": gbs, "unit": "GB/s", "cores": cores, "kind": "port",
               "sample": f"{sample} of the {NBLK} blocks, forward+adjoint, {cores} worker threads x 1 BLAS thread, "
                         f"3 steps ({dt * 1e3:.1f} ms/step)"}

    if rank == 0:
        peak, peak_src = peaks()
        es = np.dtype(DTYPE).itemsize
        local_blocks = nblk // world + (1 if rank < nblk % world else 0)     # rank 0's shard
        if args.single_process:
            for name in ("gemv_n", "gemv_t"):                               # stats hold the max over local ranks' totals
                kstats[name]["launches"] //= world
        per_launch = local_blocks * (BM * BN + BM + BN) * es        # algorithmic bytes of one gemv launch
        kn, kt = kstats["gemv_n"], kstats["gemv_t"]
        dom = "gemv_t" if kt["ms"] >= kn["ms"] else "gemv_n"
        kd = kstats[dom]
        ach = per_launch / (kd["ms"] / max(kd["launches"], 1) * 1e-3) / 1e9 if kd["launches"] else None
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "traffic_current.json")
        if world == 1 and nblk == NBLK and os.path.exists(tpath):     # the ncu capture is of exactly this launch
            with open(tpath) as f:
                tj = json.load(f).get(dom + "_kernel")
            if tj:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_src = "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch (profiles/traffic_current.json)"
        kernels = {}
        for name, st in kstats.items():
            if st["launches"]:
                kernels[name] = {"launches": st["launches"], "avg_us": st["ms"] / st["launches"] * 1e3}
                if name in ("gemv_n", "gemv_t"):
                    kernels[name]["GBps"] = per_launch / (st["ms"] / st["launches"] * 1e-3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload_config(world, nblk),
            "applies_per_s": 2.0 / (ms_per_step * 1e-3),
            "frac_of_aggregate_hbm_peak": value / (peak * world),
            "roofline": {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s",
                         "frac": (ach / peak) if ach else None, "traffic": traffic, "traffic_source": traffic_src,
                         "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": per_launch, "kernels": kernels},
            "cpu_baseline": cpu,
            "e2e": {"value": bytes_total / e2e_s / 1e9, "unit": "GB/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_s * 1e3},
            "gpu_launches": launches if (world == 1 or args.single_process) else int(launches * world),
            "mode": "one host process" if (args.single_process or world == 1) else "one process per GPU (NCCL)",
            "clocks": clocks, "host_cores": os.cpu_count(), "parity_relerr_block1": parity, "build_s": t_build,
            "schedule": {"fwd_tiles_items": A.schedule_info(False), "adj_tiles_items": A.schedule_info(True)},
        }
        print(json.dumps(line), file=args.out, flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def quiet_stdout():
    """Keep stdout for the ONE JSON line: libraries that print to fd 1 (NCCL's version banner, for
    instance) are sent to stderr; returns a file object on the original stdout."""
    sys.stdout.flush()
    keep = os.dup(1)
    os.dup2(2, 1)
    return os.fdopen(keep, "w")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--blocks", type=int, default=NBLK, help="number of diagonal blocks (default: the BASELINE config)")
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"],
                    help="strong (default): the BASELINE operator's 64 blocks are sharded over the N ranks; "
                         "weak: 64 blocks per rank")
    ap.add_argument("--single-process", action="store_true",
                    help="drive all --gpus N ranks from this one process (the Julia-master model) instead of one "
                         "process per GPU under torchrun")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile", action="store_true",
                    help="short run for ncu: skips the e2e pass, the clock-sampling loop and the CPU baseline")
    for name in ("fwd_tx", "fwd_tc", "adj_ru", "adj_tc", "ld_policy", "pdl", "graphs", "p2p", "auto_tile"):
        ap.add_argument(f"--{name}", type=int, default=None)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    args.out = quiet_stdout()
    if args.impl == "reference":
        return run_reference(args)
    return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
