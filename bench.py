#!/usr/bin/env python
"""bench.py — GP train+predict test-points/sec on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c4|c1|c2|c3|c3n256|c3n484|c5|spgp|...]

Default workload = BASELINE.json configs[3] ("c4"): the batched small-GP stream, 50 000 independent
GPs per GPU (n = 128, 3-D inputs, Matern32, 128 test points each, float32) — the configuration
the metric's "at 1/2/4/8 B200" is quoted on.  A "step" is one pass of the hot path (fused
Gram + Cholesky + alpha + predict, one kernel launch) over the whole batch.  GPUs shard the GP
stream (weak scaling: 50k GPs per rank, no data-path collective; torch.distributed is used
only for the barrier and the max-over-ranks of the device time).

`value`  : test points / s with all inputs resident in HBM (CUDA events on the launch stream).
`e2e`    : the same metric through the C-ABI host-buffer call erl_gp_batch_train_predict_f32
           (pinned host buffers; H2D of the training sets and queries and D2H of mean /
           variance / valid / info inside the timed region).
`--workload c1|c2|c3|c3n256|c3n484|c5|spgp`: the other BASELINE.json configurations (bench_workloads.py), same line schema: `value` with
           device-resident inputs, `e2e` through the C-ABI calls with pinned host buffers, `roofline` of the dominant kernel
           (HBM for the partitioned sensor GPs, FP64 tensor peak for the dense / SPGP paths), `cpu_baseline`, `clocks`.
`--impl reference`: the reference's CPU path (the OpenMP oracle restatement — the reference
           itself cannot be built here) on all host cores, bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import bench_workloads  # noqa: E402  (numpy only at import time)

SEED = 6
WORKLOADS = {
    # name: (description, num_gps per GPU, n, x_dim, queries per GP, kernel, scale, dtype)
    "c4": dict(desc="batched small-GP stream: 50k independent GPs per GPU, n=128, x_dim=3, Matern32(0.3), 128 test points per GP, f32 (BASELINE.json configs[3])",
               num_gps=50_000, n=128, x_dim=3, q_per_gp=128, kernel="matern32", scale=0.3, dtype="f32"),
    # diagnostics: the larger partitions of RangeSensorGaussianProcess3D (default grouping n <= 192; (16,2)^2 grouping n = 256)
    "n192": dict(desc="diagnostic: 8k independent GPs per GPU, n=192, x_dim=2, Matern32(0.3), 192 test points per GP, f32",
                 num_gps=8_000, n=192, x_dim=2, q_per_gp=192, kernel="matern32", scale=0.3, dtype="f32"),
    "n256": dict(desc="diagnostic: 5k independent GPs per GPU, n=256, x_dim=2, Matern32(0.3), 256 test points per GP, f32",
                 num_gps=5_000, n=256, x_dim=2, q_per_gp=256, kernel="matern32", scale=0.3, dtype="f32"),
    "n192f64": dict(desc="diagnostic: 8k independent GPs per GPU, n=192, x_dim=2, Matern32(0.3), 192 test points per GP, f64",
                    num_gps=8_000, n=192, x_dim=2, q_per_gp=192, kernel="matern32", scale=0.3, dtype="f64"),
    # diagnostic: the same stream in double (DMMA row-GP kernel); not the bench line
    "c4f64": dict(desc="batched small-GP stream in double: 50k independent GPs per GPU, n=128, x_dim=3, Matern32(0.3), 128 test points per GP, f64 (diagnostic)",
                  num_gps=50_000, n=128, x_dim=3, q_per_gp=128, kernel="matern32", scale=0.3, dtype="f64"),
}


def algorithmic_bytes_per_gp(n, d, t, s):
    """n(d+2)s in, (n^2+n)s out (L, alpha), t*d*s queries in, t(2s+1) out."""
    return n * (d + 2) * s + (n * n + n) * s + t * d * s + t * (2 * s + 1)


def flops_per_gp(n, d, t):
    """useful flops: Cholesky n^3/3 + alpha 2n^2 + predict t(n^2 + 2n); kernel evaluations not counted."""
    return n ** 3 / 3 + 2 * n * n + t * (n * n + 2 * n)


def synth_batch(w, rank):
    rng = np.random.default_rng(SEED + 1000 * rank)
    b, n, d, t = w["num_gps"], w["n"], w["x_dim"], w["q_per_gp"]
    dt = np.float32 if w["dtype"] == "f32" else np.float64
    x = rng.random((b, n, d), dtype=np.float32).astype(dt)
    wv = rng.uniform(1, 4, (b, 1, d)).astype(dt)
    y = (0.5 * np.sin(wv * x * 3.0).sum(axis=2)).astype(dt)
    var = np.full((b, n), 0.01, dtype=dt)
    n_train = np.full(b, n, dtype=np.int32)
    q_offsets = (np.arange(b + 1, dtype=np.int64) * t)
    q_x = rng.random((b * t, d), dtype=np.float32).astype(dt)
    return n_train, x, y, var, q_offsets, q_x


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def wait_first_sample(self, timeout=5.0):
        """nvidia-smi needs ~1 s to start: block until it delivers so that short timed regions are still sampled."""
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.05)
        self.lines.clear()  # keep only samples taken under load

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                smax.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(smax) if smax else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return json.load(open(path)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


def profile_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu --set full capture, or None."""
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return None
    return None


def cpu_baseline(w, sample_gps, threads=None, min_seconds=10.0):
    """Time the CPU reference path (oracle port: OpenMP over GPs, each Reset/Train/Test) on a bounded sample:
    `sample_gps` GPs of the workload, repeated until at least `min_seconds` of CPU work have been timed."""
    import oracle

    # all host threads: torchrun exports OMP_NUM_THREADS=1 to its workers, which would silently time one core
    oracle.set_num_threads(threads or len(os.sched_getaffinity(0)))
    cores = oracle.num_threads()
    ww = dict(w, num_gps=sample_gps)
    n_train, x, y, var, q_offsets, q_x = synth_batch(ww, 0)
    kid = oracle.KERNELS[w["kernel"]]
    oracle.batched_train_predict(kid, w["scale"], n_train[:64], x[:64], y[:64], var[:64], q_offsets[:65], q_x[: 64 * w["q_per_gp"]])  # warm
    reps = 0
    t0 = time.perf_counter()
    while True:
        oracle.batched_train_predict(kid, w["scale"], n_train, x, y, var, q_offsets, q_x)
        reps += 1
        dt = time.perf_counter() - t0
        if dt >= min_seconds:
            break
    return {"value": reps * sample_gps * w["q_per_gp"] / dt, "unit": "test-points/s", "cores": cores, "kind": "port",
            "sample": f"{reps} x {sample_gps} of the workload's GPs (n={w['n']}, {w['q_per_gp']} test points each), {dt:.2f} s on {cores} OpenMP threads", "seconds": dt,
            "points": reps * sample_gps * w["q_per_gp"]}


def run_reference(args, w, rank, world):
    if rank != 0:
        return
    times = []
    base = None
    for i in range(args.warmup + args.steps):
        base = cpu_baseline(w, args.ref_sample, min_seconds=args.ref_seconds)
        if i >= args.warmup:
            times.append(base["seconds"] * 1e3 / base["points"])  # ms per test point
    ms_per_point = sum(times) / len(times)
    value = 1e3 / ms_per_point
    base["value"] = value
    ms = base["seconds"] * 1e3
    line = {"impl": "reference", "metric": "gp_train_predict_test_points_per_sec", "value": value, "unit": "test-points/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["desc"], "sample": base["sample"], "note": "CPU reference path = OpenMP oracle port (the reference cannot be built in this image)"},
            "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": value, "unit": "test-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def run_generic(args):
    """c1 / c2 / c3 / c5 / spgp: same contract as the c4 line (see the module docstring of bench_workloads.py)."""
    wl = bench_workloads.make(args.workload)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    metric = "gp_train_predict_test_points_per_sec"
    if args.impl == "reference":
        if rank != 0:
            return
        vals, base = [], None
        for i in range(args.warmup + args.steps):
            base = wl.cpu_baseline(args.ref_seconds)
            if i >= args.warmup:
                vals.append(base["value"])
        value = sum(vals) / len(vals)
        base["value"] = value
        print(json.dumps({"impl": "reference", "metric": metric, "value": value, "unit": "test-points/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": base["seconds"] * 1e3, "higher_is_better": True, "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
                          "config": {"workload": wl.desc, "sample": base["sample"], "note": "CPU reference path = OpenMP oracle port (the reference cannot be built in this image)"},
                          "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")},
                          "e2e": {"value": value, "unit": "test-points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return

    import torch
    import torch.distributed as dist

    import erl_gaussian_process_b200 as gp

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = gp.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    wl.setup(gp, torch, ctx, dev, stream, rank, world)
    for _ in range(args.warmup):
        wl.step_dev()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first_sample()
    wl.step_dev()
    barrier()
    wl.dom_events.clear()
    if hasattr(wl, "train_events"):
        wl.train_events.clear()
    launches0 = ctx.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        wl.step_dev()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - launches0
    ms_dom = wl.dominant_ms()
    if len(sampler.lines) < 3:
        t_s = time.perf_counter()
        while len(sampler.lines) < 3 and time.perf_counter() - t_s < 3.0:
            wl.step_dev()
            torch.cuda.synchronize()
    clocks = sampler.stop()
    wl.check()
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    units = torch.tensor([float(wl.units)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(units, op=dist.ReduceOp.SUM)
    total_units = float(units.item())
    value = total_units / (ms_step * 1e-3)

    e2e = None
    if not args.no_e2e:
        for _ in range(2):
            wl.step_e2e()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            wl.step_e2e()
        barrier()
        dt = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt = float(tt.item())
        wl.check_e2e()
        e2e = {"value": total_units / dt, "unit": "test-points/s", "h2d_bytes_per_step": int(wl.h2d), "d2h_bytes_per_step": int(wl.d2h), "ms_per_step": dt * 1e3,
               "api": "C-ABI train + test calls with pinned host buffers (copies inside the timed region)"}
    if rank == 0:
        peaks, peak_kind = measured_peaks()
        roof = wl.roofline(ms_dom, peaks)
        roof.setdefault("peak_source", peak_kind)
        line = {"metric": metric, "value": value, "unit": "test-points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
                "scaling": wl.scaling, "vs_baseline": None, "dtype": wl.dtype, "data": "synthetic",
                "config": {"workload": wl.desc, "test_points_per_step": total_units, "sharding": wl.sharding,
                           "l2": "every step re-reads its inputs from HBM-resident buffers and rewrites K / L / slabs larger than L2 for c5; c1 / c2 / c3 / spgp working sets (8 MB - 0.5 GB) are "
                                 "not flushed between steps: they are latency / launch bound, not bandwidth bound (see roofline.note)"},
                "roofline": roof, "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e}
        if world == 1 and not args.no_cpu_baseline:
            base = wl.cpu_baseline(10.0)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs NVML reports as local to GPU `index` (several ranks per box: every rank uploads 205 MB per
    step from pinned host memory; first-touch on the GPU's own NUMA node keeps the copies off the inter-socket link)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(mask) for b in range(64) if (word >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
    except Exception:  # no NVML / no permission: run unbound
        pass


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c4", choices=sorted(WORKLOADS) + list(bench_workloads.NAMES))
    ap.add_argument("--num-gps", type=int, default=None, help="override GPs per GPU (smoke runs)")
    ap.add_argument("--ref-sample", type=int, default=4000, help="GPs in the CPU reference sample")
    ap.add_argument("--ref-seconds", type=float, default=2.0, help="--impl reference: CPU seconds per step (the sample is repeated)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-variants", action="store_true", help="N=1: also time the host-buffer call with alpha and with L + alpha written back to the host")
    ap.add_argument("--rowgp-tc", type=int, default=-1, choices=[-1, 0, 1],
                    help="fused FP32 kernel for n <= 128: 1 = tcgen05 / TMEM kernel, 0 = mma.sync kernel, -1 = library default (mma.sync)")
    ap.add_argument("--no-tc-variant", action="store_true", help="N=1: do not time the tcgen05 kernel in a child process beside the default kernel")
    ap.add_argument("--no-other-workloads", action="store_true",
                    help="default workload, N=1: do not append the short child runs of the other BASELINE.json configurations (`other_workloads` in the JSON line)")
    ap.add_argument("--phase", default="fused", choices=["fused", "train", "predict", "split"],
                    help="diagnostics only: time the train kernel, the predict kernel or both as separate launches (the reported metric is always the fused step)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.workload in bench_workloads.NAMES:
        run_generic(args)
        return
    w = dict(WORKLOADS[args.workload])
    if args.num_gps:
        w["num_gps"] = args.num_gps
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return

    import torch
    import torch.distributed as dist

    import erl_gaussian_process_b200 as gp

    torch.cuda.set_device(local_rank)
    if world > 1:
        bind_to_gpu_numa_node(local_rank)  # before any pinned allocation: the host buffers should sit next to this rank's GPU
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    b, n, d, t = w["num_gps"], w["n"], w["x_dim"], w["q_per_gp"]
    np_dt = np.float32 if w["dtype"] == "f32" else np.float64
    s = np.dtype(np_dt).itemsize
    ctx = gp.Context(local_rank)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)  # our kernels run on torch's current stream: its events time them
    ctx.set_rowgp_tc(args.rowgp_tc)
    batch = gp.BatchGp(b, n, d, w["kernel"], w["scale"], np_dt, ctx)

    # pinned host buffers (the e2e path reads / writes these)
    n_train, x, y, var, q_offsets, q_x = synth_batch(w, rank)

    def pinned(a):
        tns = torch.from_numpy(a).pin_memory()
        return tns, tns.numpy()

    keep = []
    host = {}
    for name, arr in (("n_train", n_train), ("x", x), ("y", y), ("var", var), ("q_offsets", q_offsets), ("q_x", q_x)):
        tns, view = pinned(arr)
        keep.append(tns)
        host[name] = view
    tq = b * t
    h_mean_t, h_mean = pinned(np.zeros(tq, dtype=np_dt))
    h_var_t, h_var = pinned(np.zeros(tq, dtype=np_dt))
    h_valid_t, h_valid = pinned(np.zeros(tq, dtype=np.uint8))
    h_info_t, h_info = pinned(np.zeros(b, dtype=np.int32))

    # resident inputs for the kernel-only measurement
    batch.upload(host["n_train"], host["x"], host["y"], host["var"])
    d_off = torch.from_numpy(q_offsets).to(dev)
    d_qx = torch.from_numpy(q_x).to(dev)
    d_mean = torch.empty(tq, dtype=d_qx.dtype, device=dev)
    d_var = torch.empty(tq, dtype=d_qx.dtype, device=dev)
    d_valid = torch.empty(tq, dtype=torch.uint8, device=dev)

    def step():
        if args.phase == "fused":
            batch.train_predict_dev(d_off, d_qx, tq, d_mean, d_var, d_valid, min_num_samples=0, write_l=True)
        if args.phase in ("train", "split"):
            batch.train_dev(min_num_samples=0, write_l=True)
        if args.phase in ("predict", "split"):
            batch.predict_dev(d_off, d_qx, tq, d_mean, d_var, d_valid)

    if args.phase == "predict":
        batch.train_dev(min_num_samples=0, write_l=True)

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first_sample()
    for _ in range(2):  # the GPU is busy again when the sampled region starts
        step()
    barrier()
    launches0 = ctx.kernel_launches
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step()
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.kernel_launches - launches0
    if len(sampler.lines) < 3:
        # the K timed steps were shorter than a few sampler periods: keep the same load running (untimed) until
        # nvidia-smi has reported the clocks it runs at
        t_s = time.perf_counter()
        while len(sampler.lines) < 3 and time.perf_counter() - t_s < 3.0:
            step()
            torch.cuda.synchronize()
    clocks = sampler.stop()
    tms = torch.tensor([ms_total], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_step = float(tms.item()) / args.steps
    value = world * tq / (ms_step * 1e-3)

    # sanity: the timed kernel really produced finite predictions
    assert args.phase == "train" or (bool(torch.isfinite(d_mean).all()) and bool(d_valid.all())), "kernel output invalid"

    # ---- e2e: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        import ctypes as C

        from erl_gaussian_process_b200.host import _p

        fn = ctx.fn("erl_gp_batch_train_predict", np_dt)

        def e2e_step():
            rc = fn(batch.handle, C.c_long(0), _p(host["n_train"]), _p(host["x"]), _p(host["y"]), _p(host["var"]), _p(host["q_offsets"]), _p(host["q_x"]), C.c_long(tq), None, None,
                    _p(h_info), _p(h_mean), _p(h_var), _p(h_valid))
            assert rc == 0, rc

        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(3, min(args.steps, 10))
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt_e2e = (time.perf_counter() - t0) / e2e_steps
        tt = torch.tensor([dt_e2e], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dt_e2e = float(tt.item())
        assert np.isfinite(h_mean).all() and h_valid.all() and (h_info == 0).all()
        h2d = host["n_train"].nbytes + host["x"].nbytes + host["y"].nbytes + host["var"].nbytes + host["q_offsets"].nbytes + host["q_x"].nbytes
        d2h = h_mean.nbytes + h_var.nbytes + h_valid.nbytes + h_info.nbytes
        e2e = {"value": world * tq / dt_e2e, "unit": "test-points/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "ms_per_step": dt_e2e * 1e3,
               "api": "erl_gp_batch_train_predict_f32 (C ABI, pinned host buffers; L stays device-resident, materialised on demand by erl_gp_batch_download)"}
        if world == 1 and args.e2e_variants:
            # the reference's Train() materialises alpha (and L) on the host (BatchGaussianProcessUpdateTorch::GetGpResult,
            # src/batch_gp_update_torch.cpp:86-98): the same call with alpha / L + alpha written back, so that the PCIe cost of
            # those semantics is on record beside the predict-only e2e
            variants = {}
            h_alpha_t, h_alpha = pinned(np.zeros((b, n), dtype=np_dt))
            h_l_t, h_l = pinned(np.zeros((b, n, n), dtype=np_dt))
            for name, lp, ap_ in (("alpha_back", None, h_alpha), ("l_alpha_back", h_l, h_alpha)):
                def var_step():
                    rc = fn(batch.handle, C.c_long(0), _p(host["n_train"]), _p(host["x"]), _p(host["y"]), _p(host["var"]), _p(host["q_offsets"]), _p(host["q_x"]), C.c_long(tq), _p(lp), _p(ap_),
                            _p(h_info), _p(h_mean), _p(h_var), _p(h_valid))
                    assert rc == 0, rc
                var_step()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(3):
                    var_step()
                torch.cuda.synchronize()
                dtv = (time.perf_counter() - t0) / 3
                extra = h_alpha.nbytes + (h_l.nbytes if lp is not None else 0)
                variants[name] = {"value": tq / dtv, "unit": "test-points/s", "ms_per_step": dtv * 1e3, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h + extra)}
            assert np.isfinite(h_alpha).all() and np.isfinite(h_l[:: max(1, b // 64)]).all()
            e2e["variants"] = variants
            del h_l_t, h_alpha_t

    # ---- strong scaling: ONE stream of b GPs cut into `world` contiguous ranges (BASELINE configs[3]: "50k independent GPs ...
    #      sharded across 1/2/4/8 GPUs").  (a) one process per GPU (this launch): rank r takes range r, device-timed and from host
    #      buffers, max over ranks; (b) ONE process driving all the GPUs through erl_gp_batch_train_predict_multi_f32 (rank 0,
    #      the other ranks wait): the host gather is the devices' copies into the caller's arrays, inside the timed region.
    strong = None
    if world > 1 and args.phase == "fused" and not args.no_e2e:
        import ctypes as C

        from erl_gaussian_process_b200.host import _p

        bs = b // world
        g0 = rank * bs
        sl = slice(g0, g0 + bs)
        q0 = int(q_offsets[g0])
        off_s = np.ascontiguousarray(q_offsets[g0:g0 + bs + 1] - q0)
        tqs = int(off_s[-1])
        part = gp.BatchGp(bs, n, d, w["kernel"], w["scale"], np_dt, ctx)
        part.upload(host["n_train"][sl], host["x"][sl], host["y"][sl], host["var"][sl])
        d_off_s = torch.from_numpy(off_s).to(dev)

        def strong_step():
            part.train_predict_dev(d_off_s, d_qx[q0:q0 + tqs], tqs, d_mean[q0:q0 + tqs], d_var[q0:q0 + tqs], d_valid[q0:q0 + tqs], min_num_samples=0, write_l=True)

        for _ in range(3):
            strong_step()
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            strong_step()
        e1.record(stream)
        barrier()
        tms_s = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device=dev)
        dist.all_reduce(tms_s, op=dist.ReduceOp.MAX)
        off_host_t, off_host = pinned(off_s)
        fn_s = ctx.fn("erl_gp_batch_train_predict", np_dt)

        def strong_e2e_step():
            rc = fn_s(part.handle, C.c_long(0), _p(host["n_train"][sl]), _p(host["x"][sl]), _p(host["y"][sl]), _p(host["var"][sl]), _p(off_host), _p(host["q_x"][q0:q0 + tqs]), C.c_long(tqs),
                      None, None, _p(h_info[sl]), _p(h_mean[q0:q0 + tqs]), _p(h_var[q0:q0 + tqs]), _p(h_valid[q0:q0 + tqs]))
            assert rc == 0, rc

        for _ in range(2):
            strong_e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            strong_e2e_step()
        barrier()
        tt = torch.tensor([(time.perf_counter() - t0) / 10], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        strong = {"scaling": "strong", "gps_total": bs * world, "gps_per_gpu": bs,
                  "kernel": {"ms_per_step": float(tms_s.item()), "value": bs * world * t / (float(tms_s.item()) * 1e-3), "unit": "test-points/s"},
                  "e2e": {"ms_per_step": float(tt.item()) * 1e3, "value": bs * world * t / float(tt.item()), "unit": "test-points/s",
                          "api": "one process per GPU, erl_gp_batch_train_predict_f32 on the rank's contiguous range (pinned host buffers)"}}
        del part
        # (b) one process, all the GPUs.  The other ranks wait on a CPU (gloo) barrier: inside an NCCL barrier their GPUs would run
        # a spinning collective kernel, and rank 0's kernels on those GPUs would be time-sliced against it.
        barrier()
        cpu_group = dist.new_group(backend="gloo")
        if rank == 0:
            try:
                multi = gp.MultiDeviceBatchGp(bs * world, n, d, w["kernel"], w["scale"], np_dt, devices=list(range(world)))
                nb = bs * world
                tqm = int(q_offsets[nb])
                off_m_t, off_m = pinned(np.ascontiguousarray(q_offsets[:nb + 1]))
                out = {"mean": h_mean[:tqm], "var": h_var[:tqm], "valid": h_valid[:tqm], "info": h_info[:nb]}

                def multi_step():
                    multi.train_predict(host["n_train"][:nb], host["x"][:nb], host["y"][:nb], host["var"][:nb], off_m, host["q_x"][:tqm], want_l=False, want_alpha=False, out=out)

                for _ in range(2):
                    multi_step()
                t0 = time.perf_counter()
                for _ in range(10):
                    multi_step()
                dtm = (time.perf_counter() - t0) / 10
                assert np.isfinite(h_mean[:tqm]).all() and h_valid[:tqm].all() and (h_info[:nb] == 0).all()
                strong["single_process_multi_device"] = {"ms_per_step": dtm * 1e3, "value": nb * t / dtm, "unit": "test-points/s", "devices": world,
                                                         "api": "erl_gp_batch_train_predict_multi_f32: one host thread + three streams per device, results written straight into the caller's pinned arrays"}
                del multi
            except Exception as exc:  # e.g. the launcher restricted this rank to one visible device
                strong["single_process_multi_device"] = {"unavailable": repr(exc)[:200]}
        dist.barrier(group=cpu_group)
        barrier()

    if rank == 0:
        peaks, peak_kind = measured_peaks()
        alg_bytes = algorithmic_bytes_per_gp(n, d, t, s) * b
        achieved = alg_bytes / (ms_step * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"])
        traffic = profile_traffic()
        fl = flops_per_gp(n, d, t) * b
        line = {
            "metric": "gp_train_predict_test_points_per_sec", "value": value, "unit": "test-points/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": w["dtype"], "data": "synthetic",
            "config": {"workload": w["desc"], "gps_per_gpu": b, "n_train": n, "x_dim": d, "test_points_per_gp": t, "kernel": w["kernel"], "scale": w["scale"], "seed": SEED,
                       "sharding": f"{world} x {b} GPs, contiguous GP ranges per rank, no data-path collective",
                       "l2": f"per-step inputs+outputs {alg_bytes / 1e9:.2f} GB >> 126 MB L2, no flush needed"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_kind,
                         "traffic": None if traffic is None else traffic.get("dram_bytes_per_launch"),
                         "kernel": (f"rowgp::RowGpKernel<x_dim={d}, NBLK={(n + 15) // 16}, train+predict>" if w["dtype"] == "f32" else (f"rowgp64::RowGp64Kernel<x_dim={d}, NBLK={(n + 15) // 16}, train+predict> (mma.sync.m8n8k4.f64)" if n <= 128 and not os.environ.get("ERL_GP_BATCH_LEGACY")
                                          else f"BatchedGpKernel<double, x_dim={d}, n<={n}, train+predict>")) + " (one launch per step)", "algorithmic_bytes_per_launch": alg_bytes,
                         # FP64: the DMMA pipe (37.03 TFLOP/s measured, tools/mma_rate.cu) bounds the double-precision stream
                         "fp64_tensor_pipe": {"useful_tflops": fl / (ms_step * 1e-3) / 1e12, "peak_tflops": 37.03, "frac": fl / (ms_step * 1e-3) / 1e12 / 37.03} if w["dtype"] == "f64" else None,
                         # the kernel's real bound: 3xTF32 mma.sync work (3 HMMA products per FP32 product; 276 TFLOP/s TF32 measured
                         # => 92 TFLOP/s FP32-equivalent); 590 HMMA.1688 per 16 queries and GP at n = t = 128
                         "tensor_pipe": {"useful_tflops_fp32_equiv": fl / (ms_step * 1e-3) / 1e12, "peak_tflops_fp32_equiv": 276.46 / 3,
                                         "frac": fl / (ms_step * 1e-3) / 1e12 / (276.46 / 3), "source": "tools/mma_rate.cu (profiles/r01_mma_rate.jsonl)"} if w["dtype"] == "f32" else None,
                         "fp32_pipe": {"useful_tflops": fl / (ms_step * 1e-3) / 1e12, "peak_tflops": 71.05, "note": "FFMA peak measured with tools/mma_rate.cu (FFMA2 with fresh operands sustains ~55, tools/fma_lds_rate.cu); factorisation and predict run on the tensor pipe as 3xTF32 mma.sync (276 TFLOP/s TF32 peak = 92 FP32-equivalent), pivot blocks / back-substitution / covariance entries on the FP32 pipe; the kernel is latency / issue bound, not HBM bound, see DESIGN.md"}},
            "clocks": clocks, "gpu_launches": int(launches), "e2e": e2e,
        }
        if strong is not None:
            line["strong"] = strong
        tc_now = args.rowgp_tc == 1 or (args.rowgp_tc < 0 and os.environ.get("ERL_GP_ROWGP_TC", "0") not in ("", "0"))
        if w["dtype"] == "f32" and n <= 128 and tc_now:
            line["roofline"]["kernel"] = f"rowgp_tc::RowGpTcKernel<x_dim={d}> (tcgen05.mma kind::tf32, TMEM accumulators; one launch per step)"
        if world == 1 and w["dtype"] == "f32" and n <= 128 and not tc_now and not args.no_tc_variant and args.phase == "fused":
            # the same step on the tcgen05 / TMEM kernel, in a child process (a kernel under development must not be able to
            # take the measured process down with it); device-timed, inputs resident, like `value`
            import subprocess

            cmd = [sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--rowgp-tc", "1", "--steps", str(args.steps), "--warmup", str(args.warmup), "--no-e2e",
                   "--no-cpu-baseline"] + (["--num-gps", str(args.num_gps)] if args.num_gps else [])
            try:
                res = subprocess.run(cmd, capture_output=True, text=True, timeout=180)
                child = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
                line["tcgen05_variant"] = {"kernel": child["roofline"]["kernel"], "ms_per_step": child["ms_per_step"], "value": child["value"], "unit": child["unit"],
                                           "roofline_frac": child["roofline"]["frac"], "select": "erl_gp_context_set_rowgp_tc(ctx, 1) or ERL_GP_ROWGP_TC=1"}
            except Exception as exc:
                line["tcgen05_variant"] = {"unavailable": repr(exc)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            base = cpu_baseline(w, args.ref_sample)
            line["cpu_baseline"] = {k: base[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if world == 1 and args.workload == "c4" and args.phase == "fused" and not args.no_other_workloads and not tc_now and not args.num_gps:
            # The other BASELINE.json configurations, each as a short child run of this same script (same timing rules, clocks
            # sampled inside the child): the headline stays C4, these put a number next to every configuration in the same record.
            import subprocess

            others = {}
            for name, extra in (("c4f64", ["--steps", "5"]), ("c1", ["--steps", "10"]), ("c2", ["--steps", "10"]), ("c3", ["--steps", "10"]), ("c3n256", ["--steps", "10"]),
                                ("spgp", ["--steps", "5"]), ("c5", ["--steps", "1", "--no-e2e"])):
                cmd = [sys.executable, os.path.abspath(__file__), "--workload", name, "--warmup", "3" if name != "c5" else "1", "--no-cpu-baseline", "--no-other-workloads"] + extra
                try:
                    res = subprocess.run(cmd, capture_output=True, text=True, timeout=150)
                    child = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][-1])
                    others[name] = {"workload": child["config"]["workload"], "ms_per_step": child["ms_per_step"], "value": child["value"], "unit": child["unit"], "dtype": child["dtype"],
                                    "steps": child["steps"], "warmup": child["warmup"], "e2e_ms_per_step": (child.get("e2e") or {}).get("ms_per_step"),
                                    "roofline": {k: child["roofline"].get(k) for k in ("bound", "achieved", "peak", "unit", "frac", "kernel", "train") if k in child["roofline"]},
                                    "clocks": child.get("clocks")}
                    if "phases" in child:
                        others[name]["phases"] = child["phases"]
                except Exception as exc:
                    others[name] = {"unavailable": repr(exc)[:200]}
            line["other_workloads"] = others
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
