"""bench.py workloads other than the batched stream (c4): BASELINE.json configs[0] (c1), configs[1] (c2), configs[2] (c3) and
configs[4] (c5, spgp).  Every workload offers the same four things to bench.py's timing loop:

    step_dev()   one pass of the hot path with every input and output resident in HBM (torch device tensors handed to the
                 C ABI, which accepts host or device pointers); the dominant sub-call is bracketed by CUDA events on the
                 launch stream (self.dom_events) so that the roofline uses that kernel's own duration
    step_e2e()   the same pass through the reference-facing C-ABI call with pinned HOST buffers (H2D / D2H inside)
    roofline(ms_dominant)   SURVEY.md 8(d) algorithmic bytes / flops of the dominant kernel over its measured duration
    cpu_baseline(seconds)   the CPU port (oracle) on a bounded sample of the same workload
"""
from __future__ import annotations

import ctypes as C
import os
import time

import numpy as np

FP64_TENSOR_PEAK_TFLOPS = 37.03  # mma.sync.m8n8k4.f64 rate measured on this pool's B200 (tools/mma_rate.cu, profiles/r01_mma_rate.jsonl)
FP64_PEAK_NOTE = "measured DMMA (mma.sync m8n8k4 f64) issue rate, tools/mma_rate.cu -> profiles/r01_mma_rate.jsonl; MEASURED_PEAKS.json carries no FP64 figure (cuBLAS DGEMM reaches 35.5)"


def _pin(torch, a):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


class Base:
    scaling = "weak"  # N > 1: every rank runs its own replica of the workload (the path does not shard below one GP / one scan)
    sharding = "replicas: every rank runs the whole workload on its own GPU (no data-path collective)"

    def __init__(self):
        self.dom_events = []
        self.keep = []

    def pinned(self, a):
        t, v = _pin(self.torch, a)
        self.keep.append(t)
        return v

    def dev(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).to(self.device)

    def bracket(self, fn):
        e0 = self.torch.cuda.Event(enable_timing=True)
        e1 = self.torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        fn()
        e1.record(self.stream)
        self.dom_events.append((e0, e1))

    def dominant_ms(self):
        ts = [a.elapsed_time(b) for a, b in self.dom_events]
        return sum(ts) / max(1, len(ts))


# ----------------------------------------------------------------------------------------------------------------------
class DenseVanilla(Base):
    """c1 / c5: one dense VanillaGaussianProcess<double> (train = Gram + blocked Cholesky + alpha; test = mean + variance)."""

    def __init__(self, name, n, t, scale, desc):
        super().__init__()
        self.name, self.n, self.t, self.kscale, self.desc = name, n, t, scale, desc
        self.dtype = "f64"

    def setup(self, gp, torch, ctx, device, stream, rank, world):
        self.gp, self.torch, self.ctx, self.device, self.stream = gp, torch, ctx, device, stream
        n, t = self.n, self.t
        x, y, var, xt = self.data()
        if world > 1 and self.name == "c5":
            # the factorisation is "replicas only" (SURVEY.md 8e); the test points shard in contiguous ranges
            from erl_gaussian_process_b200 import sharding

            t0, t1 = sharding.shard_range(t, rank, world)
            xt = np.ascontiguousarray(xt[t0:t1])
            self.scaling = "strong"
            self.sharding = f"{world} ranks: every rank factors its own replica (n = {n}), contiguous ranges of the {t} test points per rank, host gather of mean / variance"
        self.t_local = len(xt)
        self.units = self.t_local
        self.h_x, self.h_y, self.h_var, self.h_xt = self.pinned(x), self.pinned(y), self.pinned(var), self.pinned(xt)
        self.h_mean, self.h_v = self.pinned(np.zeros(self.t_local)), self.pinned(np.zeros(self.t_local))
        self.d_x, self.d_y, self.d_var, self.d_xt = self.dev(x), self.dev(y), self.dev(var), self.dev(xt)
        self.d_mean = torch.empty(self.t_local, dtype=torch.float64, device=device)
        self.d_v = torch.empty(self.t_local, dtype=torch.float64, device=device)
        s = gp.VanillaGaussianProcess.Setting("matern32", self.kscale, max_num_samples=-1)
        self.g = gp.VanillaGaussianProcess(s, np.float64, ctx)
        self.h2d = self.h_x.nbytes + self.h_y.nbytes + self.h_var.nbytes + self.h_xt.nbytes
        self.d2h = self.h_mean.nbytes + self.h_v.nbytes
        self.train_events = []

    def data(self):
        rng = np.random.default_rng(1)
        x = rng.uniform(-1, 1, (self.n, 2))
        y = 2 * np.sin(10 * x[:, 0]) * np.cos(10 * x[:, 1])
        var = np.full(self.n, 1e-3)
        xt = np.random.default_rng(2).uniform(-1, 1, (self.t, 2))
        return x, y, var, xt

    def step_dev(self):
        g = self.g
        e0 = self.torch.cuda.Event(enable_timing=True)
        e1 = self.torch.cuda.Event(enable_timing=True)
        e0.record(self.stream)
        g.train_dev(self.d_x, self.d_y, self.d_var, self.n, 2, 1)
        e1.record(self.stream)
        self.train_events.append((e0, e1))
        self.bracket(lambda: g.test_dev(self.d_xt, self.t_local, 2, self.d_mean, self.d_v))

    def step_e2e(self):
        g = self.g
        fn_tr, fn_te = self.ctx.fn("erl_gp_vanilla_train", np.float64), self.ctx.fn("erl_gp_vanilla_test", np.float64)
        info = C.c_int(0)
        p = lambda a: a.ctypes.data_as(C.c_void_p)
        rc = fn_tr(g.handle, C.c_int(1), C.c_double(self.kscale), C.c_long(2), C.c_long(1), C.c_long(self.n), p(self.h_x), C.c_long(2), p(self.h_y), C.c_long(self.n), p(self.h_var), C.byref(info))
        assert rc == 0 and info.value == 0, (rc, info.value)
        rc = fn_te(g.handle, C.c_long(self.t_local), p(self.h_xt), C.c_long(2), p(self.h_mean), p(self.h_v))
        assert rc == 0, rc

    def check(self):
        self.ctx.synchronize()
        assert self.g.vanilla_info() == 0
        assert bool(self.torch.isfinite(self.d_mean).all()) and bool(self.torch.isfinite(self.d_v).all())

    def check_e2e(self):
        assert np.isfinite(self.h_mean).all() and np.isfinite(self.h_v).all()

    def roofline(self, ms_dom, peaks):
        n, t = self.n, self.t_local
        flops = t * (n * n + 2 * n)  # SURVEY.md 8(d) "Large predict": n^2 per column for the solve + 2n for the mean
        ach = flops / (ms_dom * 1e-3) / 1e12
        tr = [a.elapsed_time(b) for a, b in self.train_events]
        ms_train = sum(tr) / max(1, len(tr))
        potrf = n ** 3 / 3 / (ms_train * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_TENSOR_PEAK_TFLOPS, "traffic": None,
                "peak_source": FP64_PEAK_NOTE, "kernel": "PredictVarianceKernelDmma16<double> (+ PredictMeanKernel, VarianceFinalize): erl_gp_vanilla_test_dev", "ms_dominant": ms_dom,
                "algorithmic_flops_per_launch": flops,
                "train": {"what": "erl_gp_vanilla_train_dev = Gram + blocked look-ahead Cholesky (DMMA) + alpha", "ms": ms_train, "flops": n ** 3 / 3, "achieved": potrf,
                          "peak": FP64_TENSOR_PEAK_TFLOPS, "frac": potrf / FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s"}}

    def cpu_baseline(self, min_seconds, threads=None):
        import oracle

        oracle.set_num_threads(threads or len(os.sched_getaffinity(0)))
        cores = oracle.num_threads()
        n, t = self.n, self.t
        x, y, var, xt = self.data()
        if n <= 2048:
            o = oracle.VanillaGp(oracle.MATERN32, self.kscale, np.float64, max_num_samples=n)
            reps, t0 = 0, time.perf_counter()
            while True:
                assert o.train(x, y, var) == 0
                o.test(xt)
                reps += 1
                dt = time.perf_counter() - t0
                if dt >= min_seconds:
                    break
            return {"value": reps * t / dt, "unit": "test-points/s", "cores": cores, "kind": "port", "seconds": dt, "points": reps * t,
                    "sample": f"{reps} x the whole workload (train n={n} + test {t} points, parallel=true path), {dt:.2f} s on {cores} OpenMP threads"}
        # c5: the 1M-point variance (2.7e14 flop) is infeasible on the CPU in the time budget: time the training
        # once and a slice of the test points, extrapolate the predict linearly in T
        ts = 1024
        o = oracle.VanillaGp(oracle.MATERN32, self.kscale, np.float64, max_num_samples=-1)
        t0 = time.perf_counter()
        assert o.train(x, y, var) == 0
        t_train = time.perf_counter() - t0
        xt = xt[:ts]
        t0 = time.perf_counter()
        o.test(xt)
        t_test = time.perf_counter() - t0
        est = t_train + t_test * (t / ts)
        return {"value": t / est, "unit": "test-points/s", "cores": cores, "kind": "port", "seconds": t_train + t_test, "points": ts,
                "sample": f"train n={n} once ({t_train:.1f} s) + test of {ts} of the {t} points ({t_test:.1f} s) on {cores} OpenMP threads; whole-job time extrapolated linearly in T: {est:.0f} s"}


# ----------------------------------------------------------------------------------------------------------------------
def _batch_bytes(n_train, counts, d, s):
    """SURVEY.md 8(d), batched small-GP train+predict: per GP n(d+2)s + (n^2+n)s + t d s + t(2s+1)."""
    n = np.asarray(n_train, dtype=np.float64)
    t = np.asarray(counts, dtype=np.float64)
    return float((n * (d + 2) * s + (n * n + n) * s + t * d * s + t * (2 * s + 1)).sum())


class Lidar(Base):
    """c2: LidarGaussianProcess2D<float>, one 1080-beam scan, group 64 / overlap 18 => 24 partitions, OU(0.05), 100k test rays."""

    name, dtype = "c2", "f32"
    desc = "LidarGp2D<float> single 1080-beam synthetic scan, 24 partitions (43, 22 x 64, 43), OU(0.05), mapping 1/sqrt(r), 100000 test rays (BASELINE.json configs[1])"

    def setup(self, gp, torch, ctx, device, stream, rank, world):
        self.gp, self.torch, self.ctx, self.device, self.stream = gp, torch, ctx, device, stream
        dtype = np.float32
        n, t = 1080, 100_000
        ang, ranges, self.q = self.data()
        s = gp.LidarGaussianProcess2D.Setting()
        s.group_size, s.overlap_size, s.margin, s.symmetric_partitions = 64, 18, 1, True
        s.sensor_range_var, s.discontinuity_var = 0.01, 100.0
        s.sensor_frame.angle_min, s.sensor_frame.angle_max, s.sensor_frame.num_rays = float(ang[0]), float(ang[-1]), n
        s.sensor_frame.valid_range_min, s.sensor_frame.valid_range_max = 0.1, 30.0
        s.gp.kernel_type, s.gp.scale = "ou", 0.05
        s.mapping_type = 2
        self.lg = lg = gp.LidarGaussianProcess2D(s, dtype, ctx)
        lg.sensor_frame.angles = ang
        assert lg.num_partitions == 24
        frame = lg.sensor_frame
        frame.update_ranges(np.eye(2), np.zeros(2), ranges)  # erl_geometry's part (masks), outside the path
        self.t = self.units = t
        rot = np.asfortranarray(frame.rotation.astype(dtype))
        self.rot = rot
        hit, con = np.ascontiguousarray(frame.mask_hit, dtype=np.uint8), np.ascontiguousarray(frame.mask_continuous, dtype=np.uint8)
        self.h_ranges, self.h_hit, self.h_con, self.h_q = self.pinned(frame.ranges.astype(dtype)), self.pinned(hit), self.pinned(con), self.pinned(self.q)
        self.h_mean, self.h_var, self.h_valid = self.pinned(np.zeros(t, dtype)), self.pinned(np.zeros(t, dtype)), self.pinned(np.zeros(t, np.uint8))
        self.d_ranges, self.d_hit, self.d_con, self.d_q = self.dev(self.h_ranges), self.dev(hit), self.dev(con), self.dev(self.q)
        self.d_mean, self.d_var = torch.empty(t, dtype=torch.float32, device=device), torch.empty(t, dtype=torch.float32, device=device)
        self.d_valid = torch.empty(t, dtype=torch.uint8, device=device)
        self.fn_train, self.fn_test = ctx.fn("erl_gp_lidar2d_train", dtype), ctx.fn("erl_gp_lidar2d_test", dtype)
        self.h2d = self.h_ranges.nbytes + hit.nbytes + con.nbytes + self.h_q.nbytes
        self.d2h = self.h_mean.nbytes + self.h_var.nbytes + self.h_valid.nbytes
        self.frame_arrays = (frame.ranges.astype(dtype), frame.mask_hit.copy(), frame.mask_continuous.copy())
        self.ang = ang

    @staticmethod
    def data():
        n, t, dtype = 1080, 100_000, np.float32
        rng = np.random.default_rng(3)
        ang = np.linspace(-3 * np.pi / 4, 3 * np.pi / 4, n).astype(dtype)
        ranges = (5 + 2 * np.sin(3 * ang) + 0.5 * np.sign(np.sin(7 * ang))).astype(dtype)
        ranges[rng.random(n) < 0.02] = 1e3
        q = np.random.default_rng(4).uniform(-3 * np.pi / 4, 3 * np.pi / 4, t).astype(dtype)
        return ang, ranges, q

    def _run(self, ranges, hit, con, q, mean, var, valid):
        from erl_gaussian_process_b200.host import _p

        h = self.lg.handle
        rc = self.fn_train(h, _p(self.rot), _p(ranges), _p(hit), _p(con))
        assert rc == 0, rc
        self.bracket_or_call(lambda: self.fn_test(h, _p(q), C.c_long(self.t), C.c_int(1), C.c_int(1), _p(mean), _p(var), _p(valid)))

    def step_dev(self):
        self.bracket_or_call = self.bracket
        self._run(self.d_ranges, self.d_hit, self.d_con, self.d_q, self.d_mean, self.d_var, self.d_valid)

    def step_e2e(self):
        self.bracket_or_call = lambda f: f()
        self._run(self.h_ranges, self.h_hit, self.h_con, self.h_q, self.h_mean, self.h_var, self.h_valid)

    def check(self):
        v = self.d_valid.bool()
        assert int(v.sum()) > 0.9 * self.t and bool(self.torch.isfinite(self.d_mean[v]).all())

    def check_e2e(self):
        v = self.h_valid.astype(bool)
        assert v.sum() > 0.9 * self.t and np.isfinite(self.h_mean[v]).all() and np.isfinite(self.h_var[v]).all()

    def _sizes(self):
        parts = self.lg.angle_partitions
        hit = self.frame_arrays[1]
        n_train = [int(hit[a:b].sum()) for a, b, _, _ in parts]
        # rays per partition: first matching closed interval (src/lidar_gp_2d.cpp:398-411)
        cl = np.array([p[2] for p in parts]); cr = np.array([p[3] for p in parts])
        inside = (self.q[:, None] >= cl[None]) & (self.q[:, None] <= cr[None])
        first = np.where(inside.any(axis=1), inside.argmax(axis=1), -1)
        counts = np.bincount(first[first >= 0], minlength=len(parts))
        return n_train, counts

    def roofline(self, ms_dom, peaks):
        n_train, counts = self._sizes()
        by = _batch_bytes(n_train, counts, 1, 4)
        ach = by / (ms_dom * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"])
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "ms_dominant": ms_dom, "algorithmic_bytes_per_launch": by,
                "kernel": "erl_gp_lidar2d_test: LidarAssign + counting sort + rowgp::RowGpKernel<x_dim=1, predict> over 24 GPs",
                "note": "24 partition GPs (<= 64 points) put at most 24 x tiles CTAs on 148 SMs: this configuration cannot approach the HBM roofline whatever the kernel; it is latency bound (SURVEY.md 8e)"}

    def cpu_baseline(self, min_seconds, threads=None):
        import oracle

        oracle.set_num_threads(threads or len(os.sched_getaffinity(0)))
        cores = oracle.num_threads()
        ang, ranges, q = self.data()
        hit = np.isfinite(ranges) & (ranges >= 0.1) & (ranges <= 30.0)  # LidarFrame2D::UpdateRanges hit mask (stand-in, host.py)
        con = np.ones(len(ranges), dtype=bool)  # discontinuity detection off
        self.t = len(q)
        og = oracle.LidarGp2D(ang, oracle.KERNELS["ou"], 0.05, 64, 18, 1, True, 0.01, 100.0, False, 2, 1.0, 0.1, 30.0, np.float32)
        reps, t0 = 0, time.perf_counter()
        while True:
            assert og.train(ranges, hit, con)
            og.test(q, True, True)
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
        return {"value": reps * self.t / dt, "unit": "test-points/s", "cores": cores, "kind": "port", "seconds": dt, "points": reps * self.t,
                "sample": f"{reps} x the whole workload (Train of the scan + Test of {self.t} rays: serial per-ray TestResult construction, omp mean / variance), {dt:.2f} s on {cores} OpenMP threads"}


class Range3d(Base):
    """c3: RangeSensorGaussianProcess3D<float> on a 480 x 640 range image, full-image predict."""

    dtype = "f32"

    def __init__(self, name, rg, ro, cg, co):
        super().__init__()
        self.name, self.group = name, (rg, ro, cg, co)
        self.desc = (f"RangeSensorGp3D<float> synthetic 480x640 range image, grouping ({rg},{ro})x({cg},{co}), Matern32(0.05), full-image predict (T = 307200) (BASELINE.json configs[2]; "
                     "its '32x24 grid' is not reachable with the reference's partition formula, SURVEY.md 8d: the grid actually produced is reported)")

    def setup(self, gp, torch, ctx, device, stream, rank, world):
        self.gp, self.torch, self.ctx, self.device, self.stream = gp, torch, ctx, device, stream
        dtype = np.float32
        rows, cols = 480, 640
        rg, ro, cg, co = self.group
        s3 = gp.RangeSensorGaussianProcess3D.Setting()
        s3.row_group_size, s3.row_overlap_size, s3.col_group_size, s3.col_overlap_size = rg, ro, cg, co
        s3.sensor_frame.azimuth_min, s3.sensor_frame.azimuth_max, s3.sensor_frame.num_azimuth_lines = -0.6, 0.6, rows
        s3.sensor_frame.elevation_min, s3.sensor_frame.elevation_max, s3.sensor_frame.num_elevation_lines = -0.8, 0.8, cols
        s3.sensor_frame.valid_range_min, s3.sensor_frame.valid_range_max = 0.1, 30.0
        s3.gp.kernel_type, s3.gp.scale = "matern32", 0.05
        self.s3 = s3
        self.rg3 = rg3 = gp.RangeSensorGaussianProcess3D(s3, dtype, ctx)
        self.fc = fc = rg3.sensor_frame.frame_coords
        img = self.image()
        frame = rg3.sensor_frame
        frame.update_ranges(np.eye(3), np.zeros(3), img.astype(dtype))
        coords = fc.reshape(-1, 2).astype(dtype).copy()
        self.t = self.units = t = len(coords)
        ranges = np.asfortranarray(frame.ranges.astype(dtype))
        hit = np.asfortranarray(frame.mask_hit.astype(np.uint8))
        self.frame_arrays = (frame.ranges.copy(), frame.mask_hit.copy())
        self.coords = coords
        # the C ABI reads the range image column-major (rows x cols), as the reference's Eigen matrix
        self.h_ranges, self.h_hit, self.h_coords = self.pinned(ranges.T), self.pinned(hit.T), self.pinned(coords)
        self.h_mean, self.h_var, self.h_valid = self.pinned(np.zeros(t, dtype)), self.pinned(np.zeros(t, dtype)), self.pinned(np.zeros(t, np.uint8))
        self.d_ranges, self.d_hit, self.d_coords = self.dev(self.h_ranges), self.dev(self.h_hit), self.dev(coords)
        self.d_mean, self.d_var = torch.empty(t, dtype=torch.float32, device=device), torch.empty(t, dtype=torch.float32, device=device)
        self.d_valid = torch.empty(t, dtype=torch.uint8, device=device)
        self.fn_train, self.fn_test = ctx.fn("erl_gp_range3d_train", dtype), ctx.fn("erl_gp_range3d_test", dtype)
        self.h2d = self.h_ranges.nbytes + self.h_hit.nbytes + self.h_coords.nbytes
        self.d2h = self.h_mean.nbytes + self.h_var.nbytes + self.h_valid.nbytes
        nr, nc = rg3.grid
        self.grid = (int(nr), int(nc))
        self.desc += f"; grid produced: {nr} x {nc} = {nr * nc} GPs, n <= {rg * cg}"

    @staticmethod
    def image():
        rows, cols = 480, 640
        r, c = np.meshgrid(np.arange(rows), np.arange(cols), indexing="ij")
        img = 4.0 + 0.8 * np.sin(r / 9.0) * np.cos(c / 13.0) + 0.002 * c
        img[np.random.default_rng(5).random((rows, cols)) < 0.05] = np.inf
        return img.astype(np.float32)

    def _run(self, ranges, hit, coords, mean, var, valid, bracket):
        from erl_gaussian_process_b200.host import _p

        h = self.rg3.handle
        rc = self.fn_train(h, _p(ranges), _p(hit))
        assert rc == 0, rc
        bracket(lambda: self.fn_test(h, _p(coords), None, C.c_long(self.t), C.c_int(1), _p(mean), _p(var), _p(valid)))

    def step_dev(self):
        self._run(self.d_ranges, self.d_hit, self.d_coords, self.d_mean, self.d_var, self.d_valid, self.bracket)

    def step_e2e(self):
        self._run(self.h_ranges, self.h_hit, self.h_coords, self.h_mean, self.h_var, self.h_valid, lambda f: f())

    def check(self):
        v = self.d_valid.bool()
        assert int(v.sum()) > 0.9 * self.t and bool(self.torch.isfinite(self.d_mean[v]).all())

    def check_e2e(self):
        v = self.h_valid.astype(bool)
        assert v.sum() > 0.9 * self.t and np.isfinite(self.h_mean[v]).all() and np.isfinite(self.h_var[v]).all()

    def roofline(self, ms_dom, peaks):
        rg3 = self.rg3
        hit = self.frame_arrays[1]
        rows_p, cols_p = rg3.partitions(0), rg3.partitions(1)
        n_train = np.array([[int(hit[ra:rb, ca:cb].sum()) for (ra, rb, _, _) in rows_p] for (ca, cb, _, _) in cols_p]).ravel()  # (col, row) order as the GP grid
        n_train = np.where(n_train > 32, n_train, 0)  # train iff cnt > min_num_samples_per_group (src/range_sensor_gp_3d.cpp:358)
        counts = np.full(len(n_train), self.t / len(n_train))
        by = _batch_bytes(n_train, counts, 2, 4)
        ach = by / (ms_dom * 1e-3) / 1e9
        peak = float(peaks["hbm_gbs"])
        return {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "ms_dominant": ms_dom, "algorithmic_bytes_per_launch": by,
                "kernel": ("erl_gp_range3d_test: Range3dAssign + counting sort + " + (f"rowgp::RowGpKernel<x_dim=2, NBLK={(self.group[0] * self.group[2] + 15) // 16}, predict>"
                                                                                       if self.group[0] * self.group[2] <= 256 else "largegp::PredictKernel<float> (L resident in HBM / L2)")),
                "note": "the L write-back of Train() belongs to the other launch of the step; bytes here are the whole step's algorithmic bytes over the predict launch only when "
                        "'ms_dominant' is the predict; see 'step_bytes_over_step_ms' for the whole step", "grid": list(self.grid)}

    def cpu_baseline(self, min_seconds, threads=None):
        import oracle

        oracle.set_num_threads(threads or len(os.sched_getaffinity(0)))
        cores = oracle.num_threads()
        rg, ro, cg, co = self.group
        az, el = np.linspace(-0.6, 0.6, 480), np.linspace(-0.8, 0.8, 640)
        fc = np.stack(np.meshgrid(az, el, indexing="ij"), axis=-1).astype(np.float32)  # LidarFrame3D stand-in (host.py)
        ranges = self.image()
        hit = np.isfinite(ranges) & (ranges >= 0.1) & (ranges <= 30.0)
        coords = fc.reshape(-1, 2).copy()
        self.t = len(coords)
        og = oracle.RangeSensorGp3D(fc, oracle.KERNELS["matern32"], 0.05, rg, ro, 0, cg, co, 0, 32, 0.01, 2, 1.0, np.float32)
        reps, t0 = 0, time.perf_counter()
        while True:
            assert og.train(ranges, hit)
            og.test(coords, None, True)
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
        return {"value": reps * self.t / dt, "unit": "test-points/s", "cores": cores, "kind": "port", "seconds": dt, "points": reps * self.t,
                "sample": f"{reps} x the whole workload (Train of the 480x640 image + Test of {self.t} directions), {dt:.2f} s on {cores} OpenMP threads"}


class Spgp(Base):
    """spgp: SparsePseudoInputGaussianProcess<double>, M = 2048, one Update of 2000 samples + Test of a 100 x 100 grid per step."""

    name, dtype = "spgp", "f64"
    desc = ("SPGP occupancy map shape: M=2048 pseudo-inputs (64x32 grid on [-3,3]^2), Matern32(0.18), noise 1e-4, per step one Update() of 2000 samples + Test() of a 100x100 grid "
            "(L_QM refactorised because Q_M changed), f64 (BASELINE.json configs[4], second half; config/spgp_occupancy_map_2d.yaml)")

    def setup(self, gp, torch, ctx, device, stream, rank, world):
        self.gp, self.torch, self.ctx, self.device, self.stream = gp, torch, ctx, device, stream
        z, x, y, var, xt = self.data()
        self.t = self.units = len(xt)
        self.g = gp.SparsePseudoInputGaussianProcess("matern32", 0.18, z, np.float64, ctx)
        self.h_x, self.h_y, self.h_var, self.h_xt = self.pinned(x), self.pinned(y), self.pinned(var), self.pinned(xt)
        self.h_mean, self.h_v = self.pinned(np.zeros(self.t)), self.pinned(np.zeros(self.t))
        self.d_x, self.d_y, self.d_var, self.d_xt = self.dev(x), self.dev(y), self.dev(var), self.dev(xt)
        self.d_mean, self.d_v = torch.empty(self.t, dtype=torch.float64, device=device), torch.empty(self.t, dtype=torch.float64, device=device)
        self.fn_up, self.fn_te = ctx.fn("erl_gp_spgp_update", np.float64), ctx.fn("erl_gp_spgp_test", np.float64)
        self.h2d = self.h_x.nbytes + self.h_y.nbytes + self.h_var.nbytes + self.h_xt.nbytes
        self.d2h = self.h_mean.nbytes + self.h_v.nbytes

    def data(self):
        gx, gy = np.linspace(-3, 3, 64), np.linspace(-3, 3, 32)
        self.z = z = np.array([[a, b] for a in gx for b in gy])
        rng = np.random.default_rng(7)
        self.m, self.ns = len(z), 2000
        x = rng.uniform(-3, 3, (self.ns, 2))
        y = np.tanh(x[:, 0] * x[:, 1])
        var = np.full(self.ns, 1e-4)
        gt = np.linspace(-3, 3, 100)
        xt = np.array([[a, b] for a in gt for b in gt])
        return z, x, y, var, xt

    def _run(self, x, y, var, xt, mean, v, bracket):
        from erl_gaussian_process_b200.host import _p

        h = self.g.handle
        bracket(lambda: self.fn_up(h, C.c_long(self.ns), _p(x), C.c_long(2), _p(y), _p(var)))
        rc = self.fn_te(h, C.c_long(self.t), _p(xt), C.c_long(2), _p(mean), _p(v))
        assert rc == 0, rc

    def step_dev(self):
        self._run(self.d_x, self.d_y, self.d_var, self.d_xt, self.d_mean, self.d_v, self.bracket)

    def step_e2e(self):
        self._run(self.h_x, self.h_y, self.h_var, self.h_xt, self.h_mean, self.h_v, lambda f: f())

    def check(self):
        assert bool(self.torch.isfinite(self.d_mean).all()) and bool(self.torch.isfinite(self.d_v).all())

    def check_e2e(self):
        assert np.isfinite(self.h_mean).all() and np.isfinite(self.h_v).all()

    def roofline(self, ms_dom, peaks):
        m, n = self.m, self.ns
        flops = 3.0 * m * m * n + 2.0 * m * n  # SURVEY.md 8(d) "SPGP update": beta trsm M^2 N + rank-N update 2 M^2 N + alpha 2MN
        ach = flops / (ms_dom * 1e-3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": ach / FP64_TENSOR_PEAK_TFLOPS, "traffic": None, "peak_source": FP64_PEAK_NOTE,
                "kernel": "erl_gp_spgp_update: Gram K_MN + GEMM-based TRSM (DMMA) + Q_M rank-N update (DMMA)", "ms_dominant": ms_dom, "algorithmic_flops_per_launch": flops}

    def cpu_baseline(self, min_seconds, threads=None):
        import oracle

        oracle.set_num_threads(threads or len(os.sched_getaffinity(0)))
        cores = oracle.num_threads()
        z, x, y, var, xt = self.data()
        self.t = len(xt)
        o = oracle.Spgp(oracle.MATERN32, 0.18, z, np.float64)
        reps, t0 = 0, time.perf_counter()
        while True:
            assert o.update(x, y, var)
            o.test(xt)
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= min_seconds:
                break
        return {"value": reps * self.t / dt, "unit": "test-points/s", "cores": cores, "kind": "port", "seconds": dt, "points": reps * self.t,
                "sample": f"{reps} x one step (Update of {self.ns} samples + Test of {self.t} points, M = {self.m}), {dt:.2f} s on {cores} OpenMP threads"}


def make(name):
    if name == "c1":
        return DenseVanilla("c1", 1024, 8192, 0.25, "VanillaGp<double> Matern32(0.25) on synthetic 2-D data, N=1024 train / 8192 test points, train + mean + variance per step (BASELINE.json configs[0])")
    if name == "c5":
        return DenseVanilla("c5", 16384, 1_000_000, 0.1, "Large dense VanillaGp<double> N=16384 Matern32(0.1), blocked Cholesky + 1M-point predict (mean + variance) per step (BASELINE.json configs[4], first half)")
    if name == "c2":
        return Lidar()
    if name == "c3":
        return Range3d("c3", 24, 6, 8, 2)
    if name == "c3n256":
        return Range3d("c3n256", 16, 2, 16, 2)
    if name == "c3n484":  # the nearest reachable grid to BASELINE's literal "32 x 24": 33 x 25 partitions of 22 x 22 samples (large-GP path)
        return Range3d("c3n484", 22, 2, 22, 2)
    if name == "spgp":
        return Spgp()
    raise KeyError(name)


NAMES = ("c1", "c2", "c3", "c3n256", "c3n484", "c5", "spgp")
