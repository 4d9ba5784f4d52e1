#!/usr/bin/env python
"""Benchmark of the hw_matrix SpMV hot path on B200 (metric of BASELINE.json: SpMV GFLOP/s + effective GB/s
against the HBM roofline, reference CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload laplacian|rmat|uniform|band] [--impl reference]

A "step" is one SpMV y = A x over the whole (per-rank) matrix: zero y, one kernel launch.  At N>1 (torchrun, one rank
per GPU) the rows are sharded over the ranks (the reference's CU dimension), x is replicated, there is no data-path
collective; per-GPU work is fixed as N grows (weak scaling).  Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 when it is unset.  The host side of the path is OpenMP (layout build, y_host += y
# in spmv_hw, the all-cores CPU baseline): give every rank its share of the cores instead - before anything loads an
# OpenMP runtime.  The reference arm works on rank 0 alone and may use them all.
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
    _share = 1 if "reference" in sys.argv else int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _share))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="laplacian", choices=["laplacian", "rmat", "uniform", "band", "poweriter"])
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32"])
    ap.add_argument("--scale", type=int, default=0, help="log2(rows per GPU) for rmat/uniform (default 24 / 23)")
    ap.add_argument("--variant", type=int, default=0)
    ap.add_argument("--cu", type=int, default=0, help="0 = 1, or enough row tiles that a tile of y is <= 16 MB when y exceeds 256 MB; "
                    "compute units of the hw_matrix layout per GPU (row tiles: >1 keeps y L2-resident on very tall matrices)")
    ap.add_argument("--cols-div-blocks", type=int, default=0, help="column block width (0 = reference default 32768)")
    ap.add_argument("--flush-l2", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-build", action="store_true", help="skip the GPU layout-builder measurement (setup_s.gpu_layout_build)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (default min(steps, 20))")
    return ap.parse_args()


def workload_spec(args, world):
    """Global shape of the synthetic matrix and this job's config dict (weak scaling: fixed rows per GPU)."""
    if args.workload == "laplacian":
        nx, ny = 2048, 2048 * world
        return dict(kind="laplacian", nx=nx, ny=ny, rows=nx * ny, cols=nx * ny,
                    name="2D 5-point Laplacian %dx%d grid (BASELINE configs[1]: 4M rows/GPU, ~21M nnz/GPU)" % (nx, ny))
    if args.workload == "rmat":
        s = (args.scale or 24) + int(np.log2(world))
        return dict(kind="rmat", scale=s, rows=1 << s, cols=1 << s,
                    name="R-MAT scale %d ef16 (0.57,0.19,0.19,0.05), many empty rows (BASELINE configs[2])" % s)
    if args.workload == "uniform":
        s = (args.scale or 23) + int(np.log2(world))
        return dict(kind="uniform", rows=1 << s, cols=1 << s, k=16,
                    name="uniform random %d rows x 16 nnz/row (BASELINE configs[3] shape)" % (1 << s))
    if args.workload == "poweriter":  # reference arm: one gold SpMV per iteration on the whole matrix
        s = args.scale or 24
        return dict(kind="rmat", scale=s, rows=1 << s, cols=1 << s, name="R-MAT scale %d ef16 (power-iteration matrix)" % s)
    return dict(kind="band", rows=10000 * world, cols=10000 * world, name="band 10k rows (BASELINE configs[0])")


def make_matrix(spmvb, spec, is_double, rank, world):
    rows = spec["rows"]
    per = rows // world
    rb, re = rank * per, (rank + 1) * per if rank < world - 1 else rows
    if spec["kind"] == "laplacian":
        return spmvb.Csr.laplacian2d(spec["nx"], spec["ny"], rb, re, is_double), rb, re
    if spec["kind"] == "rmat":
        return spmvb.Csr.rmat(spec["scale"], 16, 0.57, 0.19, 0.19, 1, rb, re, is_double), rb, re
    if spec["kind"] == "uniform":
        return spmvb.Csr.uniform(rows, spec["cols"], spec["k"], 1, rb, re, is_double), rb, re
    assert world == 1
    return spmvb.Csr.band(rows, 5, 1, is_double), 0, rows


class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def sample(self):
        if not self.nv:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
                     0x80: "hw_power_brake_slowdown"}
            for bit, name in names.items():
                if r & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def start(self):
        def loop():
            while not self._stop.is_set():
                self.sample()
                time.sleep(0.002)
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()
        self.sample()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(args, world, variant):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the main kernel, from the committed ncu --set full
    capture of this workload (profiles/traffic.json); None when no capture exists for it."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "%s_%s_v%d" % (args.workload, args.dtype, variant)
        return t.get(key) if world == 1 else None
    except Exception:
        return None


def cpu_reference_spmv(csr, is_double, budget_s=12.0, min_reps=3, max_reps=50):
    """Times the reference's CPU SpMV (spmv_gold, csr.cpp:184-194) single-threaded, exactly as the reference runs it:
    oracle/_ref (the unmodified reference compiled here) when present, else the oracle port."""
    import ctypes
    import oracle_api as oa
    vt = np.float64 if is_double else np.float32
    rows, cols, nnz = csr.rows, csr.cols, csr.nnz
    x = np.random.default_rng(1).random(cols).astype(vt)
    y = np.zeros(rows, vt)
    ci, va = csr.col_ind, csr.values
    kind = "port"
    fn = None
    if nnz < 2 ** 32 and oa.have_ref(1, 1, is_double):
        try:
            R = oa.RefLib(1, 1, is_double)
            rp32 = csr.row_ptr.astype(np.uint32)
            args = (rows, cols, nnz, oa._ptr(rp32), oa._ptr(ci), oa._ptr(va), oa._ptr(x), oa._ptr(y))
            fn = lambda: R.L.ref_spmv_gold(*args)
            kind = "reference"
        except OSError:
            fn = None
    if fn is None:
        O = oa.OracleLib()
        rp = csr.row_ptr
        args = (rows, oa._ptr(rp), oa._ptr(ci), oa._ptr(va), oa._ptr(x), oa._ptr(y), int(is_double))
        fn = lambda: O.L.orc_spmv_gold(*args)
    fn()  # warm
    times = []
    t_end = time.perf_counter() + budget_s
    while len(times) < max_reps and (len(times) < min_reps or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return kind, times, y, x


def rmat_row_bounds(scale, world, p_one=0.24):
    """Row ranges with (almost) equal expected non-zero count for an R-MAT matrix without vertex permutation: every row
    bit is 1 with probability c + d independently, so the row CDF has a closed form (no need to generate the matrix on
    every rank just to balance it)."""
    n = 1 << scale

    def cdf(r):  # P(row < r)
        acc, pref = 0.0, 1.0
        for k in range(scale - 1, -1, -1):
            if (r >> k) & 1:
                acc += pref * (1.0 - p_one)
                pref *= p_one
            else:
                pref *= (1.0 - p_one)
        return acc

    bounds = [0]
    for j in range(1, world):
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) // 2
            if cdf(mid) < j / world:
                lo = mid + 1
            else:
                hi = mid
        bounds.append(max(bounds[-1], lo // 4 * 4))
    bounds.append(n)
    return bounds


def run_poweriter(args, world, rank, local_rank):
    """BASELINE configs[4]: fp32 power iteration x <- A x / ||A x|| on an R-MAT matrix row-sharded over the GPUs; one
    step = one iteration = local SpMV + norm all-reduce + all-gather of the y slices into every rank's x (NCCL)."""
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    import torch
    import host_driver
    import spmvb
    spmvb.lib()
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    is_double = args.dtype == "f64"
    tdt = torch.float64 if is_double else torch.float32
    scale = args.scale or 24
    n = 1 << scale
    bounds = rmat_row_bounds(scale, world)
    lo, hi = bounds[rank], bounds[rank + 1]
    t0 = time.perf_counter()
    csr = spmvb.Csr.rmat(scale, 16, 0.57, 0.19, 0.19, 1, lo, hi, is_double)
    t_gen = time.perf_counter() - t0
    lay = spmvb.Layout.from_csr(csr, 1, 1, args.cols_div_blocks)
    eng = spmvb.Engine(lay, local_rank, args.variant)
    x_len = lay.blocks * (args.cols_div_blocks or 32768)
    # a side stream of our own: the legacy default stream has handle 0, which the C ABI reads as "engine stream"
    side = torch.cuda.Stream()
    torch.cuda.set_stream(side)
    x = torch.zeros(x_len, dtype=tdt, device="cuda")
    x[:n] = 1.0 / np.sqrt(n)
    plan = host_driver.GatherPlan(bounds, mode=os.environ.get("SPMVB_EXCHANGE", "broadcast"))
    y = torch.zeros(plan.max_len, dtype=tdt, device="cuda")
    stream = torch.cuda.current_stream().cuda_stream

    def spmv_local(x_full, y_local):
        eng.spmv_dev(x_full.data_ptr(), y_local.data_ptr(), accumulate=False, stream=stream)

    def sync():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def sumsq(y_local, n_local, out):
        eng.sumsq(y_local.data_ptr(), n_local, out.data_ptr(), stream=stream)

    def scale_y(y_local, n_local, ss):
        eng.scale_rsqrt(y_local.data_ptr(), y_local.data_ptr(), n_local, ss.data_ptr(), stream=stream)

    host_driver.power_iteration(spmv_local, x, y, plan, max(args.warmup, 3), dist=dist, sumsq=sumsq, scale=scale_y)
    sync()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    nrm = host_driver.power_iteration(spmv_local, x, y, plan, args.steps, dist=dist, sumsq=sumsq, scale=scale_y)
    e1.record()
    sync()
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms, float(csr.nnz), float(eng.launches - l0)], dtype=torch.float64, device="cuda")
    if dist is not None:
        tm = t[:1].clone(); dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        ts = t[1:].clone(); dist.all_reduce(ts, op=dist.ReduceOp.SUM)
        ms, nnz_total, launches = float(tm.item()), int(ts[0].item()), int(ts[1].item())
    else:
        nnz_total, launches = int(csr.nnz), int(eng.launches - l0)
    if rank == 0:
        vb = 8 if is_double else 4
        per = ms / args.steps
        alg = nnz_total * (2 + vb) + n * vb + world * n * vb
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        print(json.dumps({
            "metric": "SpMV GFLOP/s (2*nnz/t)", "value": 2.0 * nnz_total / (per * 1e-3) / 1e9, "unit": "GFLOP/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": per,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": "power iteration on R-MAT scale %d ef16, rows sharded over %d GPU(s), y slices exchanged "
                                   "into x by NCCL every iteration (BASELINE configs[4])" % (scale, world),
                       "rows": n, "nnz": nnz_total, "variant": int(eng.variant), "row_bounds": bounds,
                       "step": "clear rows + SpMV kernel + sum of squares + all-reduce + scale kernel + exchange (%s)" % ("all-to-all into equal chunks + all-gather" if plan.mode == "chunks" else "one broadcast per row owner")},
            "effective_gbs": alg / (per * 1e-3) / 1e9, "last_norm": nrm, "gpu_launches": launches,
            "setup_s": {"generate": t_gen}}), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def run_reference(args, spec, world, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same config/metric."""
    if rank != 0:
        return
    import spmvb
    is_double = args.dtype == "f64"
    csr, _, _ = make_matrix(spmvb, spec, is_double, 0, 1)  # the whole job's matrix on rank 0
    kind, _, _, _ = cpu_reference_spmv(csr, is_double, budget_s=0.0, min_reps=max(args.warmup, 1), max_reps=max(args.warmup, 1))
    kind, times, _, _ = cpu_reference_spmv(csr, is_double, budget_s=0.0, min_reps=args.steps, max_reps=args.steps)
    total = float(np.sum(times))
    gflops = 2.0 * csr.nnz * len(times) / total / 1e9
    vb = 8 if is_double else 4
    alg = csr.nnz * (2 + vb) + csr.rows * vb + csr.cols * vb
    sample = "whole %s, %d nnz, %d passes of spmv_gold (csr.cpp:184-194), single thread as the reference runs it" % (
        spec["name"], csr.nnz, len(times))
    line = {
        "impl": "reference", "metric": "SpMV GFLOP/s (2*nnz/t)", "value": gflops, "unit": "GFLOP/s", "n_gpus": world,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": spec["name"], "rows": csr.rows, "cols": csr.cols, "nnz": int(csr.nnz)},
        "effective_gbs": alg * len(times) / total / 1e9,
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if spec.get("kind") == "band" and world == 1:
        # BASELINE configs[0] IS the CPU emulation run (make TARGET=emu CU=1 VF=1 DOUBLE=1 ./run.elf): time the
        # unmodified reference's create_csr_hw_matrix + spmv_hw (HLS functions compiled for the CPU) as well
        try:
            import oracle_api as oa
            if oa.have_ref(1, 1, is_double):
                R = oa.RefLib(1, 1, is_double)
                t0 = time.perf_counter()
                h = R.build(csr.rows, csr.cols, csr.row_ptr, csr.col_ind, csr.values)
                t_build = time.perf_counter() - t0
                xx = np.random.default_rng(1).random(csr.cols).astype(np.float64 if is_double else np.float32)
                ts = []
                for _ in range(max(args.steps, 1)):
                    yy = np.zeros(csr.rows, xx.dtype)
                    t0 = time.perf_counter()
                    rc = R.spmv_hw(h, xx, yy)
                    ts.append(time.perf_counter() - t0)
                    if rc:
                        raise RuntimeError("reference spmv_hw: FIFO under-run (SURVEY Q1)")
                R.free(h)
                line["emu_spmv_hw"] = {"value": 2.0 * csr.nnz / float(np.mean(ts)) / 1e9, "unit": "GFLOP/s", "cores": 1,
                                       "kind": "reference", "ms_per_call": float(np.mean(ts)) * 1e3,
                                       "create_csr_hw_matrix_ms": t_build * 1e3,
                                       "sample": "TARGET=emu path: create_csr_hw_matrix + spmv_hw of the unmodified reference"}
        except Exception as exc:
            line["emu_spmv_hw"] = {"error": str(exc)}
    try:  # for information: the same loop spread over all host cores (oracle port, OpenMP over rows); the reference
        # itself is single-threaded, so `value` above stays its own number
        import oracle_api as oa
        O = oa.OracleLib()
        x = np.random.default_rng(1).random(csr.cols).astype(np.float64 if is_double else np.float32)
        O.spmv_gold_omp(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x, is_double)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            _, threads = O.spmv_gold_omp(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x, is_double)
            ts.append(time.perf_counter() - t0)
        line["cpu_baseline_all_cores"] = {"value": 2.0 * csr.nnz / min(ts) / 1e9, "unit": "GFLOP/s", "cores": int(threads),
                                          "kind": "port", "sample": "same matrix, best of 3 passes, OpenMP over rows"}
    except Exception as exc:
        line["cpu_baseline_all_cores"] = {"error": str(exc)}
    print(json.dumps(line), flush=True)


def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.workload == "poweriter" and args.dtype == "f64" and "--dtype" not in sys.argv:
        args.dtype = "f32"  # BASELINE configs[4] is the DOUBLE=0 run, in both arms
    if args.workload == "poweriter" and args.impl != "reference":
        run_poweriter(args, world, rank, local_rank)
        return
    spec = workload_spec(args, world)
    if args.impl == "reference":
        run_reference(args, spec, world, rank)
        return

    # Rank 0 must print exactly ONE JSON line on stdout: libraries (NCCL's version banner, ...) write there too, so
    # stdout is pointed at stderr until the line is ready.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        sys.stdout.flush()
        os.dup2(_real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)

    import torch
    import spmvb
    spmvb.lib()  # fails loudly when the CUDA library is missing: there is no fallback
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a B200; no CUDA device is visible (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    is_double = args.dtype == "f64"
    vt = np.float64 if is_double else np.float32
    vb = 8 if is_double else 4

    t0 = time.perf_counter()
    csr, rb, re = make_matrix(spmvb, spec, is_double, rank, world)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    if args.cu <= 0:  # row tiles (the layout's compute units) so that the y range being updated stays in the L2 cache
        # (uniform columns have no y locality at all: tile as soon as y exceeds half the L2; power-law rows re-touch
        #  the same heavy rows block after block and prefer one tile up to a few hundred MB)
        ybytes = csr.rows * vb
        limit = (64 << 20) if args.workload == "uniform" else (256 << 20)
        args.cu = 1 if ybytes <= limit else min(64, 1 << int(np.ceil(np.log2(ybytes / (16 << 20)))))
    lay = spmvb.Layout.from_csr(csr, args.cu, 1, args.cols_div_blocks)
    t_layout = time.perf_counter() - t0
    t0 = time.perf_counter()
    eng = spmvb.Engine(lay, local_rank, args.variant)
    t_upload = time.perf_counter() - t0
    nnz_local = int(csr.nnz)
    alg_bytes_local = int(eng.algorithmic_bytes)
    x_upload_local = int(eng.x_upload_bytes)  # set_x copies only the column blocks this rank's rows touch

    # x replicated on every rank (pinned host copy for the e2e leg), y sharded by rows
    x_host = torch.empty(csr.cols, dtype=torch.float64 if is_double else torch.float32).pin_memory()
    x_np = x_host.numpy()
    x_np[:] = np.random.default_rng(1).random(csr.cols).astype(vt)
    y_host = torch.zeros(csr.rows, dtype=x_host.dtype).pin_memory()
    eng.set_x(x_np)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        eng.sync()

    # ---- warm-up
    eng.enqueue_steps(max(args.warmup, 3), args.flush_l2)
    eng.collect_steps()

    # ---- timed region: K steps, device-timed (events at both ends of the region), clocks sampled while it runs
    sampler = ClockSampler(local_rank)
    barrier()
    launches0 = eng.launches
    sampler.start()
    eng.enqueue_steps(args.steps, args.flush_l2, inner_events=args.flush_l2)
    total_ms, kernel_ms = eng.collect_steps()
    launches = eng.launches - launches0
    # ---- the same K steps again with events around every launch of the main kernel (roofline.achieved); the events
    #      sit between the row-clearing kernel and the SpMV kernel and cost ~1.5 us per step, hence a separate pass
    if not args.flush_l2:
        barrier()
        eng.enqueue_steps(args.steps, False, inner_events=True)
        _, kernel_ms = eng.collect_steps()
    sampler.stop()
    barrier()
    t_job_ms = total_ms
    if args.flush_l2:  # the flush kernels are not part of a step: count kernel + memset only
        t_job_ms = float(np.sum(kernel_ms))
    nnz_total = nnz_local
    if dist is not None:
        t = torch.tensor([t_job_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_job_ms = float(t.item())
        n = torch.tensor([nnz_local, alg_bytes_local, launches, x_upload_local], dtype=torch.float64, device="cuda")
        dist.all_reduce(n, op=dist.ReduceOp.SUM)
        nnz_total, alg_total, launches_total = int(n[0].item()), int(n[1].item()), int(n[2].item())
        x_upload_total = int(n[3].item())
    else:
        alg_total, launches_total, x_upload_total = alg_bytes_local, launches, x_upload_local
    ms_per_step = t_job_ms / args.steps
    gflops = 2.0 * nnz_total / (ms_per_step * 1e-3) / 1e9
    eff_gbs = alg_total / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C-ABI call with host buffers (spmv_hw semantics: H2D x, SpMV, D2H y, y_host += y)
    e2e_steps = args.e2e_steps or min(args.steps, 20)
    xp = (x_host.data_ptr(), csr.cols)
    yp = y_host.data_ptr()
    for _ in range(2):
        eng.spmv_host(xp, yp, accumulate=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.spmv_host(xp, yp, accumulate=True)
    barrier()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    e2e_gflops = 2.0 * nnz_total * e2e_steps / e2e_s / 1e9

    # ---- every rank checks the first and last rows of its shard against the CPU SpMV of the same rows (the oracle as
    #      the checker): at N > 1 these are the rows that read x across the neighbouring ranks' column ranges.  A wrong
    #      result is an error, not a number.
    shard_err = 0.0
    if world > 1:
        import oracle_api as oa
        if rank == 0:
            oa.build_port()  # (re)compiles the checker if it is missing or stale: once, not by every rank at a time
        dist.barrier()
        O = oa.OracleLib()
        eng.set_x(x_np)
        eng.spmv_dev()
        y_gpu = eng.get_y()
        rp_all = csr.row_ptr
        for lo_r, hi_r in ((0, min(csr.rows, 4096)), (max(0, csr.rows - 4096), csr.rows)):
            j0 = int(rp_all[lo_r])
            rp_s = (rp_all[lo_r:hi_r + 1] - rp_all[lo_r]).astype(np.uint64)
            ci_s, va_s = csr.col_ind[j0:int(rp_all[hi_r])], csr.values[j0:int(rp_all[hi_r])]
            gold = O.spmv_gold(hi_r - lo_r, rp_s, ci_s, va_s, x_np, is_double).astype(np.float64)
            bound = O.abs_ax(hi_r - lo_r, rp_s, ci_s, va_s, x_np, is_double) * (1e-12 if is_double else 1e-5) + 1e-300
            shard_err = max(shard_err, float(np.max(np.abs(y_gpu[lo_r:hi_r].astype(np.float64) - gold) / bound)))
        t = torch.tensor([shard_err], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        shard_err = float(t.item())
        if not shard_err <= 1.0:
            raise RuntimeError("multi-GPU result check failed: error = %g x the tolerance" % shard_err)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak()
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_bytes_local / (k_ms * 1e-3) / 1e9
    line = {
        "metric": "SpMV GFLOP/s (2*nnz/t)", "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
        "config": {"workload": spec["name"], "rows": spec["rows"], "cols": spec["cols"], "nnz": nnz_total,
                   "layout": "hw_matrix CU=%d VF=1 per GPU (rows sharded over GPUs, x replicated)" % args.cu,
                   "l2": "flushed between steps" if args.flush_l2 else "inputs larger than L2 (no flush)",
                   "variant": int(eng.variant), "variant_requested": int(args.variant), "cols_div_blocks": int(args.cols_div_blocks) or 32768,
                   "pairs": int(lay.pairs), "zero_rows": int(lay.zero_rows),
                   "step": "clear listed rows of y + one SpMV kernel launch"},
        "effective_gbs": eff_gbs,
        "roofline_nominal_frac": eff_gbs / (8000.0 * world),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": x_upload_total,
                "d2h_bytes_per_step": int(spec["rows"] * vb), "steps": e2e_steps,
                "what": "spmvb_engine_spmv_host: pinned x -> GPU, kernel, y -> pinned host, y_host += y (spmv_hw semantics)"},
        "gpu_launches": launches_total,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(args, world, eng.variant), "peak_source": peak_src, "kernel": {7: "spmv_occ_kernel<3 CTAs/SM>", 6: "spmv_occ_kernel<4 CTAs/SM>", 8: "spmv_xs_kernel",
                                1: "spmv_direct_kernel"}.get(eng.variant, "variant %d" % eng.variant),
                     "kernel_ms_avg": k_ms, "kernel_ms_min": float(np.min(kernel_ms)),
                     "algorithmic_bytes_per_launch": alg_bytes_local},
        "setup_s": {"generate": t_gen, "layout_build": t_layout, "upload": t_upload},
    }
    if world > 1:
        line["check_shard_edges_max_err_over_tolerance"] = shard_err
    if world == 1 and not args.no_gpu_build and nnz_local < (1 << 29):
        # the same layout built by CUDA kernels straight into a second engine's image; reported next
        # to the host builder's time.  The first call pays the one-time kernel loading, the second is the steady state.
        first = None
        for rep in range(2):
            lay2, eng2 = spmvb.Engine.from_csr(csr.rows, csr.cols, csr.row_ptr, csr.col_ind, csr.values, args.cu, 1,
                                               is_double, args.cols_div_blocks, local_rank)
            ms = eng2.build_ms()
            if rep == 0:
                first = ms["total_ms"]
                eng2.fetch_layout()
                same = lay.difference(lay2) == ""
            eng2.free(); lay2.free()
        line["setup_s"]["gpu_layout_build"] = {
            "csr_upload_ms": ms["h2d_ms"], "build_ms": ms["build_ms"], "call_ms": ms["total_ms"], "first_call_ms": first,
            "identical_to_host_build": bool(same),
            "what": "spmvb_engine_create_from_csr (host CSR in pageable memory -> device image + engine); build_ms = CUDA "
                    "events around the build, host builder = setup_s.layout_build"}
    if world == 1 and not args.no_cpu_baseline:
        kind, times, y_cpu, x_cpu = cpu_reference_spmv(csr, is_double)
        # the same x: check the GPU result of the bench matrix against the CPU reference while we are here
        eng.set_x(x_cpu)
        eng.spmv_dev()
        y_gpu = eng.get_y()
        scale = np.abs(y_cpu).max() + 1e-300
        line["check_vs_cpu_reference_max_abs_err_over_max_abs_y"] = float(np.abs(y_gpu - y_cpu).max() / scale)
        t = float(np.mean(times))
        line["cpu_baseline"] = {"value": 2.0 * nnz_local / t / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                                "sample": "whole workload matrix, %d passes of spmv_gold (csr.cpp:184-194), single thread"
                                          % len(times), "ms_per_pass": t * 1e3}
        try:  # the same loop over all host cores (oracle port, OpenMP over rows)
            import oracle_api as oa
            O = oa.OracleLib()
            O.spmv_gold_omp(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x_cpu, is_double)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                _, threads = O.spmv_gold_omp(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x_cpu, is_double)
                ts.append(time.perf_counter() - t0)
            line["cpu_baseline_all_cores"] = {"value": 2.0 * nnz_local / min(ts) / 1e9, "unit": "GFLOP/s", "cores": threads,
                                              "kind": "port", "sample": "same matrix, best of 5 passes, OpenMP over rows"}
        except Exception as exc:  # the headline baseline above does not depend on it
            line["cpu_baseline_all_cores"] = {"error": str(exc)}
    emit(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
