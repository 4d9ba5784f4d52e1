#!/usr/bin/env python
"""Benchmark of the hw_matrix SpMV hot path on B200 (metric of BASELINE.json: SpMV GFLOP/s + effective GB/s against the
HBM roofline at 1/2/4/8 GPUs, reference CPU path timed beside it).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload uniform|rmat|laplacian|band|poweriter]
                    [--impl reference]

Default workload = the matrix BASELINE.json's Target is stated on (configs[3]): uniform random, 2^26 rows x 2^26
columns, 16 non-zeros per row = 1 073 741 824 non-zeros, fp64.  STRONG scaling: the same matrix at every N, its rows
cut into N equal contiguous ranges (one per GPU = the reference's compute-unit dimension), x replicated; a single SpMV
has no data-path collective.  A "step" is one SpMV y = A x over the rank's rows: clear the rows that need it + one
kernel launch per row tile group.  Rank 0 prints ONE JSON line.  At N = 1 the line also carries `also`: the same
measurement of BASELINE configs[1] (Laplacian) and configs[2] (R-MAT) - skip with --also "".
"""
import argparse
import json
import os
import sys
import threading
import time

# torchrun exports OMP_NUM_THREADS=1 when it is unset.  The host side of the path is OpenMP (layout build, y_host += y
# in spmv_hw, the checker): give every rank its share of the cores instead - before anything loads an OpenMP runtime.
# The reference arm works on rank 0 alone and may use them all.
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS") == "1":
    _share = 1 if "reference" in sys.argv else int(os.environ.get("WORLD_SIZE", "1"))
    os.environ["OMP_NUM_THREADS"] = str(max(1, (os.cpu_count() or 1) // _share))

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

TOL = {True: 1e-12, False: 1e-5}  # north_star: |y - y_ref| <= tol * row-wise |A||x|
METRIC = "SpMV GFLOP/s (2*nnz/t)"


def parse_args(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="uniform", choices=["uniform", "rmat", "laplacian", "band", "poweriter"])
    ap.add_argument("--dtype", default=None, choices=["f64", "f32"])
    ap.add_argument("--scale", type=int, default=0, help="log2(rows): uniform 26, rmat / poweriter 24")
    ap.add_argument("--variant", type=int, default=0, help="kernel variant (0 = the engine's choice)")
    ap.add_argument("--cu", type=int, default=1, help="compute units of the API-visible hw_matrix layout (the reference's -DCU)")
    ap.add_argument("--cols-div-blocks", type=int, default=0, help="API column block width (0 = reference default 32768)")
    ap.add_argument("--opt", action="append", default=[], help="name=value, spmvb_set_option (A/B experiments)")
    ap.add_argument("--flush-l2", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the TIMING of the CPU baseline (the result check always runs)")
    ap.add_argument("--gpu-build", action="store_true", help="also time the GPU layout builder (setup_s.gpu_layout_build)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="end-to-end steps (default min(steps, 5))")
    ap.add_argument("--also", default=None, help="comma list of further workloads measured at N = 1 into `also` "
                    "(default: laplacian,rmat for the default workload)")
    ap.add_argument("--exchange", type=int, default=2, choices=[0, 1, 2], help="power iteration: 0 NCCL grouped broadcasts, "
                    "1 peer stores to every GPU, 2 peer stores to the forwarding GPU + all-gather (default)")
    ap.add_argument("--sample-log2", type=int, default=0, help="log2(rows) of the CPU baseline's sample (default: ~2^24 non-zeros)")
    return ap.parse_args(argv)


# ------------------------------------------------------------------------------------------------------ workloads

def workload_spec(name, args, world):
    """Global shape of the synthetic matrix + the config dict both arms print (identical in both)."""
    if name == "uniform":
        s = args.scale or 26
        spec = dict(kind="uniform", rows=1 << s, cols=1 << s, k=16, scaling="strong",
                    name="uniform random %d rows x 16 nnz/row, fp64 target matrix of BASELINE configs[3] (1B nnz at 2^26 rows)" % (1 << s))
    elif name in ("rmat", "poweriter"):
        s = args.scale or 24
        spec = dict(kind="rmat", scale=s, rows=1 << s, cols=1 << s, scaling="strong",
                    name=("R-MAT scale %d ef16 (0.57,0.19,0.19,0.05), no vertex permutation: many empty rows (BASELINE configs[2])" % s)
                    if name == "rmat" else
                    ("power iteration x <- A x / ||A x|| on R-MAT scale %d ef16, y slices exchanged into x over NCCL (BASELINE configs[4])" % s))
    elif name == "laplacian":  # weak scaling like round 1: the config-2 grid per GPU
        nx, ny = 2048, 2048 * world
        spec = dict(kind="laplacian", nx=nx, ny=ny, rows=nx * ny, cols=nx * ny, scaling="weak",
                    name="2D 5-point Laplacian %dx%d grid (BASELINE configs[1]: 4M rows, ~21M nnz per GPU)" % (nx, ny))
    else:
        assert world == 1, "the band workload (BASELINE configs[0]) is a single-GPU case"
        spec = dict(kind="band", rows=10000, cols=10000, scaling="weak", name="band 10k rows, 109 970 nnz (BASELINE configs[0])")
    return spec


def rmat_row_bounds(scale, world, p_one=0.24, pair_weight=0.0, edge_factor=16, cdb=32768):
    """Row ranges of (almost) equal expected COST for an R-MAT matrix without vertex permutation.  Every row bit is 1
    with probability c + d independently, so a row's expected length depends on its number of 1-bits only and prefix
    sums over rows have a closed form (no need to generate the matrix on every rank just to balance it).
    cost(row) = non-zeros + pair_weight x (row, column block) pairs: the sparse end of the matrix costs more per non-zero
    than the dense end - one y update and one row-map entry per pair (measured on the 8 shards of scale 24, each timed by
    itself on one GPU, scripts/exp_partition.py: t = 2.75 ns/Mnnz + 5.1 ns/Mpair in fp32, 3.8 + 5.2 in fp64).  The
    expected number of distinct column blocks a row of expected length L touches is sum_b 1 - (1 - q_b)^L with q_b the
    probability of block b (column bits are 1 with probability b + d); it agrees with the built layouts within 3 %.
    pair_weight = 0 balances the non-zeros alone."""
    from math import comb
    n = 1 << scale
    nnz_total = float(edge_factor) * n
    q_one = 0.24  # b + d
    block_bits = max(0, scale - (cdb.bit_length() - 1))
    q = [(q_one ** j) * ((1.0 - q_one) ** (block_bits - j)) for j in range(block_bits + 1)]

    def row_cost(ones):
        length = nnz_total * (p_one ** ones) * ((1.0 - p_one) ** (scale - ones))
        if pair_weight == 0.0:
            return length
        pairs = sum(comb(block_bits, j) * (1.0 - (1.0 - q[j]) ** length) for j in range(block_bits + 1))
        return length + pair_weight * pairs

    cost_of = [row_cost(o) for o in range(scale + 1)]

    def prefix(r):  # cost of rows [0, r): r splits into aligned cubes whose rows share their leading bits
        if r >= n:
            return sum(comb(scale, o) * cost_of[o] for o in range(scale + 1))
        acc, ones = 0.0, 0
        for k in range(scale - 1, -1, -1):
            if (r >> k) & 1:
                acc += sum(comb(k, j) * cost_of[ones + j] for j in range(k + 1))
                ones += 1
        return acc

    total = prefix(n)
    bounds = [0]
    for j in range(1, world):
        lo, hi = 0, n
        while lo < hi:
            mid = (lo + hi) // 2
            if prefix(mid) < total * j / world:
                lo = mid + 1
            else:
                hi = mid
        bounds.append(max(bounds[-1], lo // 4 * 4))
    bounds.append(n)
    return bounds


# pairs cost 1.85 (fp32) / 1.37 (fp64) non-zeros each on B200 (see rmat_row_bounds)
RMAT_PAIR_WEIGHT = {True: 1.37, False: 1.85}


def shard_bounds(spec, world, is_double=True):
    if spec["kind"] == "rmat":
        return rmat_row_bounds(spec["scale"], world, pair_weight=RMAT_PAIR_WEIGHT[bool(is_double)])
    rows = spec["rows"]
    return [rows * r // world // 4 * 4 for r in range(world)] + [rows]


def make_matrix(spmvb, spec, is_double, rb, re, L=None):
    """Rows [rb, re) of the workload's global matrix (L = the ctypes library that generates it)."""
    if spec["kind"] == "laplacian":
        return spmvb.Csr.laplacian2d(spec["nx"], spec["ny"], rb, re, is_double, L=L)
    if spec["kind"] == "rmat":
        return spmvb.Csr.rmat(spec["scale"], 16, 0.57, 0.19, 0.19, 1, rb, re, is_double, L=L)
    if spec["kind"] == "uniform":
        return spmvb.Csr.uniform(spec["rows"], spec["cols"], spec["k"], 1, rb, re, is_double, L=L)
    return spmvb.Csr.band(spec["rows"], 5, 1, is_double, L=L)


def x_vector(cols, vt):
    """The replicated x of every arm and rank: U(0,1) like init_vector_rand(x, 1) (main.cpp:57), seeded."""
    return np.random.default_rng(1).random(cols).astype(vt)


def public_config(spec, nnz):
    return {"workload": spec["name"], "rows": int(spec["rows"]), "cols": int(spec["cols"]), "nnz": int(nnz)}


def total_nnz(spec, csr_nnz_sum):
    return int(csr_nnz_sum)


def sample_rows(spec, args, target_nnz=1 << 24):
    """Row prefix of the global matrix that the CPU arms time: about target_nnz non-zeros (all rows for small cases)."""
    rows = spec["rows"]
    if args.sample_log2:
        return min(rows, 1 << args.sample_log2)
    per_row = {"uniform": 16.0, "rmat": 16.0, "laplacian": 5.0, "band": 11.0}[spec["kind"]]
    if spec["kind"] == "rmat":  # the first rows are the heavy ones: a prefix with ~1/16 of the rows holds far more than 1/16 of the entries
        return rows if rows * per_row <= 4 * target_nnz else max(4096, rows // 64)
    n = int(target_nnz / per_row)
    return rows if rows <= 2 * n else n


# ------------------------------------------------------------------------------------------------------ helpers

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


def ncu_traffic(workload, dtype, world, variant):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the main kernel, from the committed ncu --set full
    capture of this workload (profiles/traffic.json); None when no capture exists for it."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        return t.get("%s_%s_v%d" % (workload, dtype, variant)) if world == 1 else None
    except Exception:
        return None


def scatter_roofline(dev, nnz_local, k_ms):
    """The second bound of an irregular matrix: every (row, block) pair is one scattered 8-byte
    update of y that no blocking of this layout makes local, and a B200 sustains 193 G red.global.add.f64/s into an
    L2-resident vector (288 G/s scattered 8-byte loads; measured with tools/access_probe.cu, profiles/r2/
    access_probe_b200.txt).  achieved = pairs of the device layout / kernel time."""
    if dev.get("ell"):  # sliced ELLPACK of a regular matrix: no scattered access at all
        return None
    if dev.get("wide"):  # the wide image: the scattered access of every entry is the x gather out of an L2-resident range
        peak = 288.0
        achieved = nnz_local / (k_ms * 1e-3) / 1e9
        return {"bound": "scattered x gathers (ld.global.nc out of an L2-resident column block)", "achieved": achieved,
                "peak": peak, "unit": "G gathers/s", "frac": achieved / peak, "gathers_per_nnz": 1.0,
                "y_updates_per_nnz": dev["pairs"] / max(nnz_local, 1),
                "peak_source": "measured: tools/access_probe.cu on this pool's B200 (profiles/r2/access_probe_b200.txt)"}
    peak = 193.0
    achieved = dev["pairs"] / (k_ms * 1e-3) / 1e9
    return {"bound": "scattered y updates (red.global.add into L2)", "achieved": achieved, "peak": peak, "unit": "G updates/s",
            "frac": achieved / peak, "updates_per_nnz": dev["pairs"] / max(nnz_local, 1),
            "peak_source": "measured: tools/access_probe.cu on this pool's B200 (profiles/r2/access_probe_b200.txt)"}


KERNEL_NAMES = {7: "spmv_occ_kernel<3 CTAs/SM>", 6: "spmv_occ_kernel<4 CTAs/SM>", 8: "spmv_xs_kernel", 1: "spmv_direct_kernel",
                9: "spmv_wide_kernel", 10: "spmv_ell_kernel"}


def time_reference_gold(csr, x, is_double, steps, warmup, budget_s=None):
    """Times the reference's CPU SpMV (spmv_gold, csr.cpp:184-194) single-threaded, exactly as the reference runs it:
    oracle/_ref (the unmodified reference compiled here) when present, else the oracle port.  Returns (kind, times, y)."""
    import oracle_api as oa
    vt = np.float64 if is_double else np.float32
    rows, cols, nnz = csr.rows, csr.cols, csr.nnz
    y = np.zeros(rows, vt)
    ci, va = csr.col_ind, csr.values
    kind, fn = "port", None
    if nnz < 2 ** 32 and oa.have_ref(1, 1, is_double):
        try:
            R = oa.RefLib(1, 1, is_double)
            rp32 = csr.row_ptr.astype(np.uint32)
            a = (rows, cols, nnz, oa._ptr(rp32), oa._ptr(ci), oa._ptr(va), oa._ptr(x), oa._ptr(y))
            fn = lambda: R.L.ref_spmv_gold(*a)
            kind = "reference"
        except OSError:
            fn = None
    if fn is None:
        O = oa.OracleLib()
        rp = csr.row_ptr
        a = (rows, oa._ptr(rp), oa._ptr(ci), oa._ptr(va), oa._ptr(x), oa._ptr(y), int(is_double))
        fn = lambda: O.L.orc_spmv_gold(*a)
    for _ in range(max(warmup, 1)):
        fn()
    times = []
    t_end = time.perf_counter() + (budget_s or 1e9)
    while len(times) < steps and (len(times) < 3 or time.perf_counter() < t_end):
        t0 = time.perf_counter()
        fn()
        times.append(time.perf_counter() - t0)
    return kind, times, y


class RowCheck:
    """The checker: gold = the oracle's CSR SpMV on all host cores and the row-wise bound tol * |A||x|, computed once per
    (matrix, x), outside every timed region.  err(y) = max over the rows of |y - gold| / bound: <= 1 passes."""

    def __init__(self, O, csr, x, is_double):
        self.gold, _ = O.spmv_gold_omp(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x, is_double)
        self.gold = self.gold.astype(np.float64)
        self.bound = O.abs_ax(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x, is_double) * TOL[is_double] + 1e-300

    def err(self, y, scale=1.0):
        if not len(self.gold):
            return 0.0, 0
        e = np.abs(y.astype(np.float64) * scale - self.gold) / self.bound
        worst = int(np.argmax(e))
        return float(e[worst]), worst


# ------------------------------------------------------------------------------------------------------ GPU arm

class Ctx:
    """Process-wide state of a GPU-arm run."""

    def __init__(self, args):
        self.args = args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        import torch
        import spmvb
        self.torch, self.spmvb = torch, spmvb
        spmvb.lib()  # fails loudly when the CUDA library is missing: there is no fallback
        if not torch.cuda.is_available():
            raise RuntimeError("bench.py needs a B200; no CUDA device is visible (no CPU fallback exists)")
        torch.cuda.set_device(self.local_rank)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
            self.dist = dist
        for kv in args.opt:
            k, v = kv.split("=")
            spmvb.set_option(k, int(v))

    def barrier(self, eng=None):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()
        if eng is not None:
            eng.sync()

    def reduce(self, values, op="sum"):
        """all-reduce of a list of floats over the ranks (identity at N = 1)"""
        if self.dist is None:
            return [float(v) for v in values]
        t = self.torch.tensor([float(v) for v in values], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op={"sum": self.dist.ReduceOp.SUM, "max": self.dist.ReduceOp.MAX,
                                    "min": self.dist.ReduceOp.MIN}[op])
        return [float(v) for v in t.tolist()]


def measure_spmv(ctx, name, args, steps, warmup, e2e_steps, with_cpu_baseline):
    """One workload through the engine on this job's GPUs.  Returns the JSON line (rank 0) or None."""
    torch, spmvb, world, rank = ctx.torch, ctx.spmvb, ctx.world, ctx.rank
    dtype = args.dtype or "f64"
    is_double = dtype == "f64"
    vt = np.float64 if is_double else np.float32
    vb = 8 if is_double else 4
    spec = workload_spec(name, args, world)
    bounds = shard_bounds(spec, world, is_double)
    rb, re = bounds[rank], bounds[rank + 1]

    t0 = time.perf_counter()
    csr = make_matrix(spmvb, spec, is_double, rb, re)
    t_gen = time.perf_counter() - t0
    t0 = time.perf_counter()
    lay = spmvb.Layout.from_csr(csr, args.cu, 1, args.cols_div_blocks)
    t_layout = time.perf_counter() - t0
    api_info = dict(cu=int(lay.n_cu), blocks=int(lay.blocks), pairs=int(lay.pairs), stream_bytes=int(lay.stream_bytes))
    t0 = time.perf_counter()
    eng = spmvb.Engine(lay, ctx.local_rank, args.variant)
    t_upload = time.perf_counter() - t0
    lay.free()  # the engine keeps nothing of the host layout
    nnz_local = int(csr.nnz)
    alg_local = int(eng.algorithmic_bytes)
    x_upload_local = int(eng.x_upload_bytes)  # set_x copies only the column blocks this rank's rows touch

    # x replicated on every rank (pinned host copy for the e2e leg), y sharded by rows
    x_host = torch.empty(csr.cols, dtype=torch.float64 if is_double else torch.float32).pin_memory()
    x_np = x_host.numpy()
    x_np[:] = x_vector(csr.cols, vt)
    y_host = torch.zeros(csr.rows, dtype=x_host.dtype).pin_memory()
    eng.set_x(x_np)

    # ---- warm-up
    eng.enqueue_steps(max(warmup, 3), args.flush_l2)
    eng.collect_steps()

    # ---- timed region: K steps, device-timed (events at both ends of the region), clocks sampled while it runs
    sampler = ClockSampler(ctx.local_rank)
    ctx.barrier(eng)
    launches0 = eng.launches
    sampler.start()
    eng.enqueue_steps(steps, args.flush_l2, inner_events=args.flush_l2)
    total_ms, kernel_ms = eng.collect_steps()
    launches = eng.launches - launches0
    # ---- the same K steps again with events around every launch of the main kernel (roofline.achieved); the events
    #      sit between the row-clearing kernel and the SpMV kernel and cost ~1.5 us per step, hence a separate pass
    if not args.flush_l2:
        ctx.barrier(eng)
        eng.enqueue_steps(steps, False, inner_events=True)
        _, kernel_ms = eng.collect_steps()
    sampler.stop()
    ctx.barrier(eng)
    t_job_ms = float(np.sum(kernel_ms)) if args.flush_l2 else total_ms  # the flush kernels are not part of a step
    t_job_ms, k_ms_max = ctx.reduce([t_job_ms, float(np.mean(kernel_ms))], "max")
    nnz_total, alg_total, launches_total, x_upload_total = ctx.reduce([nnz_local, alg_local, launches, x_upload_local])
    ms_per_step = t_job_ms / steps
    gflops = 2.0 * nnz_total / (ms_per_step * 1e-3) / 1e9
    eff_gbs = alg_total / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the C-ABI call with host buffers (spmv_hw semantics: H2D x, SpMV, D2H y, y_host += y)
    #      On several GPUs the call is the multi-GPU one of the C ABI (spmvb_group_spmv_host_rows): the group decides how x
    #      reaches the GPUs - 1/N of it per PCIe link + one in-place NCCL all-gather over NVLink when every shard reads all
    #      of x, the band each shard reads otherwise.
    xp = (x_host.data_ptr(), csr.cols)
    yp = y_host.data_ptr()
    grp = None
    if world > 1:
        u = torch.from_numpy(spmvb.Group.unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        ctx.dist.broadcast(u, src=0)
        grp = spmvb.Group.adopt(eng, spec["rows"], csr.cols, bounds, ctx.local_rank, u.cpu().numpy(), rank, world)

    def e2e_call():
        if grp is not None:
            grp.spmv_host_rows(xp, yp, accumulate=True)
        else:
            eng.spmv_host(xp, yp, accumulate=True)

    for _ in range(2):
        e2e_call()
    ctx.barrier(eng)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_call()
    ctx.barrier(eng)
    e2e_s = ctx.reduce([time.perf_counter() - t0], "max")[0]
    x_links = grp is not None and grp.x_over_links == 1
    if x_links:  # x crosses the host links once in total, not once per GPU
        x_upload_local = int(-(-csr.cols // world) * vb)
        x_upload_total = ctx.reduce([x_upload_local])[0]
    if grp is not None:
        grp.free()
    e2e_gflops = 2.0 * nnz_total * e2e_steps / e2e_s / 1e9

    # ---- result check, every row of every rank, per-row tolerance: a wrong result is an error, not a number.  The
    #      e2e leg's accumulation is checked too: y_host now holds (2 + e2e_steps) x A x.
    import oracle_api as oa
    if rank == 0:
        oa.build_port()  # (re)compiles the checker if it is missing or stale: once, not by every rank at a time
    if ctx.dist is not None:
        ctx.dist.barrier()
    O = oa.OracleLib()
    eng.spmv_dev()
    y_gpu = eng.get_y()
    t0 = time.perf_counter()
    chk = RowCheck(O, csr, x_np, is_double)
    err, worst = chk.err(y_gpu)
    err_acc, _ = chk.err(y_host.numpy(), 1.0 / float(2 + e2e_steps))
    t_check = time.perf_counter() - t0
    err_all, err_acc_all = ctx.reduce([err, err_acc], "max")
    if not (err_all <= 1.0 and err_acc_all <= 4.0):
        raise RuntimeError("%s result check failed: max error = %g x the tolerance (row %d of rank %d's rows), accumulated e2e "
                           "result %g x" % ("multi-GPU" if world > 1 else "GPU", err_all, worst + rb, rank, err_acc_all))

    dev = eng.device_layout
    if rank != 0:
        return None

    peak, peak_src = measured_peak()
    k_ms = float(np.mean(kernel_ms))
    achieved = alg_local / (k_ms * 1e-3) / 1e9
    variant = int(eng.variant)
    line = {
        "metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": spec["scaling"],
        "vs_baseline": None, "dtype": dtype, "data": "synthetic",
        "config": public_config(spec, nnz_total),
        "engine": {"api_layout": "hw_matrix CU=%d VF=1 COLS_DIV_BLOCKS=%d per GPU: %d blocks, %d (row, block) pairs"
                                 % (api_info["cu"], args.cols_div_blocks or 32768, api_info["blocks"], api_info["pairs"]),
                   "device_layout": dev, "variant": variant, "variant_requested": int(args.variant),
                   "rows_per_gpu": [bounds[i + 1] - bounds[i] for i in range(world)],
                   "l2": "flushed between steps" if args.flush_l2 else "inputs larger than L2 (no flush)",
                   "step": "clear the rows that need it + one SpMV kernel launch",
                   "options": {kv.split("=")[0]: int(kv.split("=")[1]) for kv in args.opt}},
        "effective_gbs": eff_gbs,
        "roofline_nominal_frac": eff_gbs / (8000.0 * world),
        "clocks": sampler.summary(),
        "e2e": {"value": e2e_gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": int(x_upload_total),
                "d2h_bytes_per_step": int(spec["rows"] * vb), "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "what": ("spmvb_group_spmv_host_rows: every GPU uploads 1/N of the pinned x over its own PCIe link, one in-place "
                         "ncclAllGather over NVLink replicates it, kernels, y slices -> pinned host, y_host += y (spmv_hw semantics)")
                if x_links else
                ("spmvb_group_spmv_host_rows" if world > 1 else "spmvb_engine_spmv_host") +
                ": pinned x -> GPU, kernel, y -> pinned host, y_host += y (spmv_hw semantics)"},
        "gpu_launches": int(launches_total),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(name, dtype, world, variant), "peak_source": peak_src,
                     "kernel": KERNEL_NAMES.get(variant, "variant %d" % variant),
                     "kernel_ms_avg": k_ms, "kernel_ms_min": float(np.min(kernel_ms)), "kernel_ms_max_over_ranks": k_ms_max,
                     "algorithmic_bytes_per_launch": alg_local,
                     "aggregate_frac": alg_total / (k_ms_max * 1e-3) / 1e9 / (peak * world),
                     "what": "rank 0's kernel: nnz*(2+vb) + rows*vb + x_touched*vb bytes / CUDA-event time of the launch; "
                             "aggregate_frac = all ranks' bytes / slowest rank's kernel / (N x peak)"},
        "scatter_roofline": scatter_roofline(dev, nnz_local, k_ms),
        "l2_note": ("back-to-back steps run faster than the measured DRAM copy rate would allow for the algorithmic bytes: x and y "
                    "(%d MB of them) fit the 126 MB L2 next to the evict-first stream and are partly reused from one step to the "
                    "next, and the next launch's prologue overlaps the previous kernel's tail (programmatic dependent launch); "
                    "roofline.achieved uses the per-launch event time instead, roofline.traffic is what one launch moves in DRAM"
                    % (int((csr.rows + csr.cols) * vb) >> 20)) if eff_gbs > peak * world else None,
        "check": {"max_err_over_tolerance": err_all, "e2e_accumulated_max_err_over_tolerance": err_acc_all,
                  "what": "every row of every rank against the oracle's CSR SpMV, |y - gold| <= %g x row-wise |A||x|" % TOL[is_double],
                  "seconds": t_check},
        "setup_s": {"generate": t_gen, "layout_build": t_layout, "upload": t_upload},
    }
    if world == 1 and args.gpu_build and nnz_local < (1 << 29):
        # the same layout built by CUDA kernels straight into a second engine's image
        first = None
        lay1 = spmvb.Layout.from_csr(csr, args.cu, 1, args.cols_div_blocks)
        for rep in range(2):
            lay2, eng2 = spmvb.Engine.from_csr(csr.rows, csr.cols, csr.row_ptr, csr.col_ind, csr.values, args.cu, 1,
                                               is_double, args.cols_div_blocks, ctx.local_rank)
            ms = eng2.build_ms()
            if rep == 0:
                first = ms["total_ms"]
                eng2.fetch_layout()
                same = lay1.difference(lay2) == ""
            eng2.free(); lay2.free()
        lay1.free()
        line["setup_s"]["gpu_layout_build"] = {
            "csr_upload_ms": ms["h2d_ms"], "build_ms": ms["build_ms"], "call_ms": ms["total_ms"], "first_call_ms": first,
            "identical_to_host_build": bool(same)}
    if world == 1 and with_cpu_baseline:
        # the reference's own CPU loop on this box, single thread, on a bounded sample: a row prefix of the same matrix
        S = sample_rows(spec, args)
        sub = csr if S >= csr.rows else make_matrix(spmvb, spec, is_double, 0, S)
        x_cpu = np.array(x_np)  # a pageable copy first touched by this thread, like the reference arm's x (x_np is pinned)
        kind, times, _ = time_reference_gold(sub, x_cpu, is_double, steps=50, warmup=1, budget_s=12.0)
        t = float(np.mean(times))
        line["cpu_baseline"] = {"value": 2.0 * sub.nnz / t / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                                "sample": "rows [0, %d) of the workload matrix = %d nnz, x full length, %d passes of spmv_gold "
                                          "(csr.cpp:184-194), single thread as the reference runs it" % (sub.rows, sub.nnz, len(times)),
                                "ms_per_pass": t * 1e3}
        try:  # the same loop over all host cores (oracle port, OpenMP over rows)
            O.spmv_gold_omp(sub.rows, sub.row_ptr, sub.col_ind, sub.values, x_cpu, is_double)
            ts = []
            for _ in range(5):
                t0 = time.perf_counter()
                _, threads = O.spmv_gold_omp(sub.rows, sub.row_ptr, sub.col_ind, sub.values, x_cpu, is_double)
                ts.append(time.perf_counter() - t0)
            line["cpu_baseline_all_cores"] = {"value": 2.0 * sub.nnz / min(ts) / 1e9, "unit": "GFLOP/s", "cores": int(threads),
                                              "kind": "port", "sample": "same sample, best of 5 passes, OpenMP over rows"}
        except Exception as exc:  # the headline baseline above does not depend on it
            line["cpu_baseline_all_cores"] = {"error": str(exc)}
    eng.free()
    return line


def measure_poweriter(ctx, args, steps, warmup):
    """BASELINE configs[4]: fp32 power iteration x <- A x / ||A x|| on an R-MAT matrix, rows sharded over the GPUs,
    through the C ABI's multi-GPU group (spmvb_group_*): one step = one iteration = clear rows + SpMV kernel + sum of
    squares + all-reduce of one double + scale kernel + ONE grouped NCCL exchange of the y slices into every GPU's x."""
    torch, spmvb, world, rank = ctx.torch, ctx.spmvb, ctx.world, ctx.rank
    dtype = args.dtype or "f32"
    is_double = dtype == "f64"
    vt = np.float64 if is_double else np.float32
    vb = 8 if is_double else 4
    spec = workload_spec("poweriter", args, world)
    n = spec["rows"]
    bounds = shard_bounds(spec, world, is_double)   # equal expected cost (non-zeros + weighted pairs), not equal non-zeros
    rb, re = bounds[rank], bounds[rank + 1]
    t0 = time.perf_counter()
    csr = make_matrix(spmvb, spec, is_double, rb, re)
    t_gen = time.perf_counter() - t0
    uid = None
    if world > 1:  # the NCCL unique id of the group: made on rank 0, handed out through the launcher's process group
        u = torch.from_numpy(spmvb.Group.unique_id() if rank == 0 else np.zeros(128, np.uint8)).cuda()
        ctx.dist.broadcast(u, src=0)
        uid = u.cpu().numpy()
    t0 = time.perf_counter()
    grp = spmvb.Group.create_rank(n, n, bounds, csr.row_ptr, csr.col_ind, csr.values, is_double, ctx.local_rank, uid, rank,
                                  world, args.variant)
    t_build = time.perf_counter() - t0
    if world > 1 and args.exchange:
        # peer-memory exchange: every rank maps the other ranks' x (CUDA IPC handles travel through the launcher)
        h = torch.from_numpy(grp.ipc_handle()).cuda()
        allh = [torch.empty_like(h) for _ in range(world)]
        ctx.dist.all_gather(allh, h)
        grp.set_peer_handles(torch.stack(allh).cpu().numpy(), args.exchange)
    exchange_mode = int(grp.exchange)
    x0 = np.full(n, 1.0 / np.sqrt(n), vt)
    grp.set_x(x0)
    grp.power_iter(max(warmup, 3))
    # ---- check of one iteration on every rank, before the timed region: the rows of A x this rank owns against the
    #      oracle's CSR SpMV on the same x, the norm against the all-reduced CPU sum, x identical on all ranks afterwards
    import oracle_api as oa
    if rank == 0:
        oa.build_port()
    if ctx.dist is not None:
        ctx.dist.barrier()
    O = oa.OracleLib()
    x_before = grp.get_x()
    nrm_gpu = grp.power_iter(1)
    y_local = grp.get_y()[rb:re]
    chk = RowCheck(O, csr, x_before, is_double)
    err, worst = chk.err(y_local)
    gold = chk.gold
    ss = ctx.reduce([float(np.sum(gold ** 2))])[0]
    x_after = grp.get_x()
    mine = x_after[rb:re].astype(np.float64) - gold / np.sqrt(ss)
    x_err = float(np.max(np.abs(mine))) * np.sqrt(n) if len(mine) else 0.0
    digest = float(np.sum(x_after.astype(np.float64) * np.arange(1, n + 1, dtype=np.float64) % 7.0))
    d_lo, d_hi = ctx.reduce([digest], "min")[0], ctx.reduce([digest], "max")[0]
    err_all, x_err_all = ctx.reduce([err, x_err], "max")
    nrm_rel = abs(nrm_gpu - np.sqrt(ss)) / np.sqrt(ss)
    if not (err_all <= 1.0 and nrm_rel <= 1e-5 and d_lo == d_hi and x_err_all <= 1e-3):
        raise RuntimeError("power-iteration check failed: rows %g x tolerance, norm rel. error %g, x digests %r / %r, "
                           "x error %g" % (err_all, nrm_rel, d_lo, d_hi, x_err_all))
    # ---- timed region
    sampler = ClockSampler(ctx.local_rank)
    ctx.barrier()
    l0 = grp.launches()
    sampler.start()
    t0 = time.perf_counter()
    nrm = grp.power_iter(steps)
    ctx.barrier()
    wall_ms = (time.perf_counter() - t0) * 1e3
    sampler.stop()
    dev_ms = grp.last_iter_ms * steps
    phases = [ctx.reduce([v], "max")[0] for v in grp.phase_ms]
    ms = ctx.reduce([dev_ms], "max")[0]
    wall_ms = ctx.reduce([wall_ms], "max")[0]
    nnz_total, launches = ctx.reduce([float(csr.nnz), float(grp.launches() - l0)])
    nnz_total, launches = int(nnz_total), int(launches)
    # ---- end to end: host x in, `steps` iterations, host x out
    x_pin = torch.from_numpy(x0.copy()).pin_memory()
    ctx.barrier()
    t0 = time.perf_counter()
    grp.set_x(x_pin.numpy())
    grp.power_iter(steps)
    x_out = grp.get_x()
    ctx.barrier()
    e2e_s = ctx.reduce([time.perf_counter() - t0], "max")[0]
    if rank != 0:
        return None
    per = ms / steps
    alg = nnz_total * (2 + vb) + n * vb + world * n * vb
    alg_local = int(csr.nnz) * (2 + vb) + (re - rb) * vb + n * vb
    peak, peak_src = measured_peak()
    line = {
        "metric": METRIC, "value": 2.0 * nnz_total / (per * 1e-3) / 1e9, "unit": "GFLOP/s", "n_gpus": world, "steps": steps,
        "warmup": max(warmup, 3), "ms_per_step": per, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": dtype, "data": "synthetic", "config": public_config(spec, nnz_total),
        "engine": {"row_bounds": bounds, "exchange_mode": exchange_mode,
                   "exchange": ["one grouped NCCL call per iteration (a broadcast per row owner, in place in x)",
                                "the normalisation kernel stores its rows into every GPU's x over NVLink (peer memory) + 8-byte "
                                "all-reduce as barrier",
                                "the normalisation kernel stores its rows into x of the forwarding GPU over NVLink (peer memory), "
                                "barrier, all-gather of the equal chunks in place (NCCL)"][exchange_mode],
                   "step": "clear rows + SpMV kernel + sum of squares + all-reduce of one double + normalisation / exchange",
                   "wall_ms_per_step": wall_ms / steps,
                   "last_iteration_phase_ms_max_over_ranks": dict(zip(("spmv_and_sumsq", "norm_allreduce", "normalise_store_barrier",
                                                                       "gather"), phases))},
        "effective_gbs": alg / (per * 1e-3) / 1e9, "last_norm": nrm, "gpu_launches": launches, "clocks": sampler.summary(),
        "e2e": {"value": 2.0 * nnz_total * steps / e2e_s / 1e9, "unit": "GFLOP/s", "h2d_bytes_per_step": int(world * n * vb / steps),
                "d2h_bytes_per_step": int(n * vb / steps), "steps": steps,
                "what": "x0 from pinned host memory to every GPU, `steps` iterations on the devices, x back to the host"},
        "roofline": {"bound": "hbm", "achieved": alg_local / (per * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": alg_local / (per * 1e-3) / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                     "what": "rank 0's algorithmic bytes of one iteration's SpMV / time of the WHOLE iteration (exchange included)",
                     "aggregate_frac": alg / (per * 1e-3) / 1e9 / (peak * world)},
        "check": {"rows_max_err_over_tolerance": err_all, "norm_rel_err": nrm_rel, "x_identical_on_all_ranks": d_lo == d_hi,
                  "x_max_err_times_sqrt_n": x_err_all},
        "setup_s": {"generate": t_gen, "layout_build_and_upload": t_build},
    }
    if world == 1 and not args.no_cpu_baseline:
        S = sample_rows(spec, args)
        sub = csr if S >= csr.rows else make_matrix(spmvb, spec, is_double, 0, S)
        kind, times, _ = time_reference_gold(sub, x0, is_double, steps=50, warmup=1, budget_s=10.0)
        t = float(np.mean(times))
        line["cpu_baseline"] = {"value": 2.0 * sub.nnz / t / 1e9, "unit": "GFLOP/s", "cores": 1, "kind": kind,
                                "sample": "one iteration's SpMV on rows [0, %d) = %d nnz, %d passes of spmv_gold, single thread"
                                          % (sub.rows, sub.nnz, len(times))}
    grp.free()
    return line


# ------------------------------------------------------------------------------------------------------ reference arm

def run_reference(args, world, rank):
    """--impl reference: the reference's own CPU implementation of the path on the host cores, same config / metric /
    unit.  Every step is one pass of spmv_gold over a bounded sample of the workload (a row prefix of the same matrix,
    x full length).  The matrix comes from oracle/_ref/libmatgen.so: this arm never loads the engine."""
    if rank != 0:
        return
    import spmvb  # the ctypes module only: nothing of libspmvb.so is loaded on this arm
    name = args.workload
    dtype = args.dtype or ("f32" if name == "poweriter" else "f64")
    is_double = dtype == "f64"
    vt = np.float64 if is_double else np.float32
    spec = workload_spec(name, args, world)
    G = spmvb.gen_lib()
    S = sample_rows(spec, args)
    csr = make_matrix(spmvb, spec, is_double, 0, S, L=G)
    x = x_vector(spec["cols"], vt) if name != "poweriter" else np.full(spec["cols"], 1.0 / np.sqrt(spec["cols"]), vt)
    kind, times, _ = time_reference_gold(csr, x, is_double, steps=args.steps, warmup=max(args.warmup, 1))
    total = float(np.sum(times))
    gflops = 2.0 * csr.nnz * len(times) / total / 1e9
    vb = 8 if is_double else 4
    # whole-workload non-zero count for the config dict (the GPU arm prints what it generated; same generator, same seed)
    nnz_cfg = {"uniform": spec["rows"] * 16}.get(spec["kind"])
    if nnz_cfg is None:
        nnz_cfg = int(csr.nnz) if S >= spec["rows"] else None
    sample = "rows [0, %d) of %s = %d nnz, x full length (%d values): %d passes of spmv_gold (csr.cpp:184-194), single " \
             "thread as the reference runs it" % (csr.rows, spec["name"], csr.nnz, spec["cols"], len(times))
    cfg = public_config(spec, nnz_cfg if nnz_cfg is not None else 0)
    if nnz_cfg is None:
        cfg["nnz"] = None  # R-MAT: only known after de-duplication of the whole matrix, which this arm does not generate
    line = {
        "impl": "reference", "metric": METRIC, "value": gflops, "unit": "GFLOP/s", "n_gpus": world,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": total / len(times) * 1e3, "higher_is_better": True,
        "scaling": spec["scaling"], "vs_baseline": None, "dtype": dtype, "data": "synthetic", "config": cfg,
        "effective_gbs": (csr.nnz * (2 + vb) + csr.rows * vb + csr.cols * vb) * len(times) / total / 1e9,
        "cpu_baseline": {"value": gflops, "unit": "GFLOP/s", "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": gflops, "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    if spec["kind"] == "band" and world == 1:
        # BASELINE configs[0] IS the CPU emulation run (make TARGET=emu CU=1 VF=1 DOUBLE=1 ./run.elf): time the
        # unmodified reference's create_csr_hw_matrix + spmv_hw (HLS functions compiled for the CPU) as well
        try:
            import oracle_api as oa
            if oa.have_ref(1, 1, is_double):
                R = oa.RefLib(1, 1, is_double)
                t0 = time.perf_counter()
                h = R.build(csr.rows, csr.cols, csr.row_ptr, csr.col_ind, csr.values)
                t_build = time.perf_counter() - t0
                ts = []
                for _ in range(max(args.steps, 1)):
                    yy = np.zeros(csr.rows, x.dtype)
                    t0 = time.perf_counter()
                    rc = R.spmv_hw(h, x, yy)
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
        O.spmv_gold_omp(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x, is_double)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter()
            _, threads = O.spmv_gold_omp(csr.rows, csr.row_ptr, csr.col_ind, csr.values, x, is_double)
            ts.append(time.perf_counter() - t0)
        line["cpu_baseline_all_cores"] = {"value": 2.0 * csr.nnz / min(ts) / 1e9, "unit": "GFLOP/s", "cores": int(threads),
                                          "kind": "port", "sample": "same sample, best of 3 passes, OpenMP over rows"}
    except Exception as exc:
        line["cpu_baseline_all_cores"] = {"error": str(exc)}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------ main

def main():
    args = parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, world, rank)
        return

    # Rank 0 must print exactly ONE JSON line on stdout: libraries (NCCL's version banner, ...) write there too, so
    # stdout is pointed at stderr until the line is ready.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ctx = Ctx(args)
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    if args.workload == "poweriter":
        line = measure_poweriter(ctx, args, args.steps, args.warmup)
    else:
        line = measure_spmv(ctx, args.workload, args, args.steps, args.warmup, e2e_steps, not args.no_cpu_baseline)
        also = args.also if args.also is not None else ("laplacian,rmat" if args.workload == "uniform" and not args.scale else "")
        if line is not None and world == 1 and also:
            # BASELINE configs[1] and [2] through the same measurement, so that one default run records all three
            # single-SpMV configurations (each with its own result check)
            line["also"] = {}
            for extra in [w for w in also.split(",") if w]:
                sub_args = parse_args(["--workload", extra, "--variant", str(args.variant)] + sum((["--opt", o] for o in args.opt), []))
                try:
                    sub = measure_spmv(ctx, extra, sub_args, max(args.steps, 20), args.warmup, min(e2e_steps, 5), False)
                    line["also"][extra] = {k: sub[k] for k in ("value", "unit", "ms_per_step", "config", "engine", "effective_gbs",
                                                               "e2e", "roofline", "check", "setup_s", "dtype")}
                except Exception as exc:  # the headline does not depend on the extras, but their failure is recorded
                    line["also"][extra] = {"error": "%s: %s" % (type(exc).__name__, exc)}
    if line is not None:
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
