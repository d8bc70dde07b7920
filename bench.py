#!/usr/bin/env python
"""bench.py — kriged target locations per second (mean + variance), BASELINE.json's metric.

    python bench.py --gpus N --steps K --warmup W [--config C2] [--impl reference]

A "step" is one pass of the Kriging hot path (neighbour search + assemble/factor/solve + mean/variance
for the local configs; RHS + triangular GEMM for the global ones) over one rank's slab of targets.
N = 1 runs the configuration the metric is quoted on (C2: local OK, Spherical, k = 20, 1e4 samples →
1000×1000 grid). N > 1 is weak scaling: the grid grows along its slowest axis and the sample count with
it (same density), every rank keeps all samples resident, owns a contiguous slab of one config-worth of
targets, and the per-rank results are all-gathered with NCCL inside the timed region.

`value` is timed on the device (CUDA events on the stream the kernels run on) with samples/bins already
resident in HBM; `e2e` is the same metric through the C ABI call a host binding makes (`gsk_krige`:
host buffers in, host buffers out — H2D of the samples, bin build, kernels, D2H of both fields inside
the timed region). The CPU oracle is used only for the `cpu_baseline` / `--impl reference` legs.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "oracle"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import numpy as np  # noqa: E402

METRIC = "kriged locations/sec (mean+variance)"
UNIT = "locations/s"


def weak_spec(gsk, name, world):
    """The named config with its slowest grid axis and its sample count multiplied by `world`."""
    cfg = gsk.synth.CONFIGS[name]
    grid = list(cfg["grid"])
    grid[-1] *= world
    if world == 1:
        return gsk.synth.config_spec(name)
    return gsk.synth.config_spec(name, grid=tuple(grid), n=cfg["n"] * world)


def peaks():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        d = json.loads(f.read_text())
        return d.get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(device)], stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        rows = [r.split(",") for r in Path(self.tmp.name).read_text().strip().splitlines() if r.count(",") >= 8]
        os.unlink(self.tmp.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except ValueError:
                continue
            for nme, val in zip(names, r[5:9]):
                if val.strip().lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        top = sorted(sm)[len(sm) // 2:]            # the loaded half of the samples
        return {"sm_mhz": statistics.median(top), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def run_reference(args):
    """--impl reference: the reference's CPU path. The reference is pure Julia and no julia binary
    exists in this image, so this times the oracle port (oracle/gsk_oracle.c, KD-tree search, OpenMP
    over targets on all host cores) on the same config/metric. Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import gskrige
    import oracle_py as O
    spec = weak_spec(gskrige, args.config, 1)
    T = spec.n_targets
    sample = min(T, args.cpu_sample)
    first = (T - sample) // 2
    slab = spec.with_slab(first, sample)
    cores = O.threads()
    for _ in range(max(1, min(args.warmup, 1))):
        O.krige(slab)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.krige(slab)
    dt = (time.perf_counter() - t0) / args.steps
    value = sample / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config, spec), "sample_targets": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} consecutive targets of the {args.config} grid per step (KD-tree search, OpenMP)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference is pure Julia (no julia binary in this image); timed: the C oracle port on all host cores",
    }
    print(json.dumps(line), flush=True)


def workload_name(cfg, spec):
    p = spec.params
    est = {0: "SimpleKriging", 1: "OrdinaryKriging", 2: f"UniversalKriging(degree={p['uk_degree']})"}[p["estimator"]]
    vg = {0: "Gaussian", 1: "Spherical", 2: "Exponential"}[p["vario_kind"]]
    k = p["max_neighbors"]
    grid = "x".join(str(g) for g in spec.grid_dims)
    return (f"{cfg}: {'local' if k else 'global'} {est} {vg}Variogram(range={p['vario_range']:g})"
            f"{f' maxneighbors={k}' if k else ''}, {spec.n_samples} {spec.dim}D samples -> {grid} grid, "
            f"block support q={spec.support[0].shape[0]}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C2", choices=["C1", "C2", "C3a", "C3b", "C4", "C5"])
    ap.add_argument("--cpu-sample", type=int, default=1_000_000, help="targets per step of the CPU legs")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--targets", type=int, default=0, help="limit the per-rank slab (debug)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: per-GPU work fixed (grid and samples grow with N, default); strong: the named config's grid is sharded over N")
    ap.add_argument("--gather", default="multicast", choices=["peer", "multicast", "nccl"],
                    help="N>1 result gather: stores into every rank's symmetric-memory buffer fused in the solve kernel "
                         "(peer), the same through one NVLS multicast store (multicast), or an NCCL all-gather (nccl)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    import gskrige

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: the Kriging path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    spec = weak_spec(gskrige, args.config, world if args.scaling == "weak" else 1)
    T = spec.n_targets
    first, count = gskrige.slab_bounds(T, rank, world)
    if args.targets:
        count = min(count, args.targets)
    k = spec.params["max_neighbors"]

    ctx = gskrige.Context(local_rank)
    stream = torch.cuda.current_stream()
    ctx.set_stream(stream.cuda_stream)
    ctx.plan(spec)                                   # samples + bins (or the global factor) resident in HBM
    plan_ms = ctx.timing()["ms_plan"]
    d_mean = torch.empty(count, dtype=torch.float64, device=dev)
    d_var = torch.empty(count, dtype=torch.float64, device=dev)
    gather_mode = "single GPU"
    hdl = None
    if world > 1:
        Tall = count * world
        if args.gather != "nccl":
            # fused gather: the solve kernel stores every result into all ranks' buffers over NVLink
            try:
                import torch.distributed._symmetric_memory as symm_mem
                sbuf = symm_mem.empty(2 * Tall, dtype=torch.float64, device=dev)
                hdl = symm_mem.rendezvous(sbuf, dist.group.WORLD)
                ptrs = [int(p) for p in hdl.buffer_ptrs]
                mean_ptrs, var_ptrs, use_mc = ptrs, [p + 8 * Tall for p in ptrs], False
                if args.gather == "multicast":
                    mc = int(getattr(hdl, "multicast_ptr", 0) or 0)
                    if mc:
                        mean_ptrs, var_ptrs, use_mc = [mc], [mc + 8 * Tall], True
                gather_mode = ("results stored by the compute kernels into every rank's symmetric-memory buffer over NVLink "
                               + ("(one NVLS multicast store per value)" if use_mc else "(P2P stores to each peer)")
                               + ", device-side barrier per step")
            except Exception as exc:  # noqa: BLE001 - fall back to NCCL, say so in the JSON line
                hdl = None
                gather_mode = f"NCCL all-gather (symmetric memory unavailable: {type(exc).__name__})"
        if hdl is None:
            g_mean = torch.empty(Tall, dtype=torch.float64, device=dev)
            g_var = torch.empty(Tall, dtype=torch.float64, device=dev)
            if args.gather == "nccl":
                gather_mode = "NCCL all-gather of mean and variance inside the timed step"
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # 256 MB > 126 MB L2

    def step():
        if hdl is not None:
            ctx.execute_peers(first, count, mean_ptrs, var_ptrs, out_offset=rank * count, multicast=use_mc)
            hdl.barrier(channel=0)                   # every rank's stores have landed everywhere
            return
        ctx.execute(first, count, d_mean.data_ptr(), d_var.data_ptr())
        if world > 1:                                # result gather: NCCL all-gather over NVLink
            dist.all_gather_into_tensor(g_mean, d_mean)
            dist.all_gather_into_tensor(g_var, d_var)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    launches_per_step = ctx.timing()["launches"]

    # ---- timed region: K steps, each bracketed by CUDA events on the launching stream; L2 flushed between ----
    sampler = ClockSampler(local_rank) if rank == 0 else None
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for a, b in evs:
        flush.zero_()
        a.record(stream)
        step()
        b.record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    clocks = sampler.stop() if sampler else None
    t = torch.tensor([dev_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms = float(t.item())
    ms_per_step = dev_ms / args.steps
    value = count * world / (ms_per_step * 1e-3)

    if hdl is not None:
        # the gathered field must be identical on every rank and hold this rank's slab at its place
        ctx.execute(first, count, d_mean.data_ptr(), d_var.data_ptr())
        torch.cuda.synchronize()
        full_mean = sbuf[:count * world]
        assert torch.equal(full_mean[rank * count:(rank + 1) * count], d_mean), "fused gather: own slab differs"
        chk = torch.stack([full_mean.sum(), sbuf[count * world:].sum()])
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "fused gather: ranks hold different gathered fields"

    # ---- dominant kernel, timed live with events around each launch (extra steps, not part of `value`) ----
    ctx.set_phase_timing(True)
    ps, pv = [], []
    for _ in range(min(args.steps, 10)):
        flush.zero_()
        ctx.execute(first, count, d_mean.data_ptr(), d_var.data_ptr())
        tt = ctx.timing()
        ps.append(tt["ms_search"]); pv.append(tt["ms_solve"])
    ctx.set_phase_timing(False)
    torch.cuda.synchronize()
    dfma, dmma = ctx.measure_fp64_peak()

    # ---- e2e: the C-ABI call with host buffers (pinned), H2D + plan + kernels + D2H inside the timed region ----
    h_mean = torch.empty(count, dtype=torch.float64).pin_memory().numpy()
    h_var = torch.empty(count, dtype=torch.float64).pin_memory().numpy()
    slab = spec.with_slab(first, count)
    e2e_ctx = gskrige.Context(local_rank)
    for _ in range(2):
        e2e_ctx.krige_into(slab, h_mean, h_var)
    barrier()
    e0 = time.perf_counter()
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(e2e_steps):
        e2e_ctx.krige_into(slab, h_mean, h_var)
    barrier()
    e2e_s = (time.perf_counter() - e0) / e2e_steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d = spec.n_samples * (spec.dim + 1) * 8 + 3 * spec.support[0].shape[0] * 8
    d2h = 16 * count
    if hdl is None:
        ctx.execute(first, count, d_mean.data_ptr(), d_var.data_ptr())
        torch.cuda.synchronize()
    assert np.array_equal(h_mean, d_mean.cpu().numpy()), "e2e and resident paths disagree"

    if rank == 0:
        flops_t = gskrige.synth.algorithmic_flops_per_target(spec)
        solve_ms = statistics.median(pv) if k else statistics.median(pv) or ms_per_step
        nlaunch_solve = max(1, int(launches_per_step) // 2) if k else 1   # one search + one solve launch per chunk
        extra = {}
        if k:
            achieved = flops_t * count / (solve_ms * 1e-3) / 1e12          # all solve launches of a step together
            kernel = "local_solve_small_kernel" if (k <= 20 and spec.params["estimator"] != 2) else "local_solve_kernel"
            bound, peak, frac = "fp64", dfma, achieved / dfma
        else:
            # global path: the Gram formulation needs only the forward triangular solve, i.e. n² flop per target
            # instead of the canonical 2(n+c)² — both are reported; frac uses the EXECUTED flops (conservative)
            solve_ms = ms_per_step
            achieved = flops_t * count / (ms_per_step * 1e-3) / 1e12
            np_ = -(-spec.n_samples // 128) * 128
            executed = float(np_) * np_ * count / (ms_per_step * 1e-3) / 1e12
            kernel, bound, peak, frac = "ygemm_dmma_kernel (+rhs_kernel, epilogue)", "tensor", dmma, executed / dmma
            extra = {"achieved_executed": executed, "frac_canonical": achieved / dmma,
                     "note": "tensor = FP64 mma.sync (DMMA); tcgen05 has no FP64 kind. frac = executed flops / measured DMMA peak"}
        hbm_peak, how = peaks()
        traffic = None
        tf = ROOT / "profiles" / "roofline_traffic.json"
        if tf.exists():
            traffic = json.loads(tf.read_text()).get(args.config)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, spec), "targets_per_gpu": count, "l2": "flushed between timed steps (256 MB write)",
                       "multi_gpu": ("slabs of the slowest axis, samples replicated; " + gather_mode) if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": count * world / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s * 1e3, "api": "gsk_krige (C ABI, pinned host buffers; includes sample upload + bin build)"},
            "gpu_launches": int(launches_per_step * args.steps),
            "roofline": {"bound": bound, "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": frac, "traffic": traffic, **extra,
                         "peak_source": "measured in this run (gsk_measure_fp64_peak): DFMA loop %.1f TFLOP/s, DMMA m16n8k16 loop %.1f TFLOP/s" % (dfma, dmma),
                         "algorithmic_flops_per_target": flops_t, "kernel_ms_per_step": solve_ms, "launches_per_step": nlaunch_solve,
                         "hbm": {"achieved_gbs": 16.0 * count / (solve_ms * 1e-3) / 1e9, "peak_gbs": hbm_peak, "peak_source": how,
                                 "algorithmic_bytes_per_target": 16}},
            "phases_ms": {"plan": plan_ms, "search": statistics.median(ps), "solve": statistics.median(pv)},
            "wall_s_timed_region": wall,
        }
        if not args.no_cpu_baseline and world == 1:
            import oracle_py as O
            T1 = spec.n_targets
            sample = min(T1, args.cpu_sample)
            sl = spec.with_slab((T1 - sample) // 2, sample)
            O.krige(sl)
            best = 1e30
            for _ in range(5):
                c0 = time.perf_counter(); O.krige(sl); best = min(best, time.perf_counter() - c0)
            cores = O.threads()   # read before the 1-thread pass below, which lowers the OpenMP thread count
            # the reference's own loop is serial (krig.jl:180,205 use no threads): the 1-thread figure is its analogue
            s1 = max(1, sample // 8)
            sl1 = spec.with_slab((T1 - s1) // 2, s1)
            best1 = 1e30
            for _ in range(2):
                c0 = time.perf_counter(); O.krige(sl1, nthreads=1); best1 = min(best1, time.perf_counter() - c0)
            line["cpu_baseline"] = {"value": sample / best, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{sample} consecutive targets of the grid, best of 5 passes (C oracle, KD-tree, OpenMP all cores)",
                                    "value_1_thread": s1 / best1,
                                    "sample_1_thread": f"{s1} consecutive targets, best of 2 passes, 1 thread"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
