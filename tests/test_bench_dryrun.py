"""bench.py's plumbing without a GPU: the CUDA engine is replaced by a stand-in (tests/bench_dryrun_patch.py), so the
numbers are meaningless, but argument handling, the multi-rank reductions (gloo instead of NCCL) and the contract of the
ONE JSON line on stdout are exercised - a typo there would otherwise only show up on the GPU box."""
import json
import os
import socket
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = "import sys; sys.path.insert(0, %r); import bench_dryrun_patch; import bench; bench.main()" % os.path.join(ROOT, "tests")
REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
            "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches", "roofline"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _check_line(out, n_gpus, steps):
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, out                        # exactly one line on stdout
    d = json.loads(lines[0])
    for k in REQUIRED:
        assert k in d, k
    assert d["n_gpus"] == n_gpus and d["steps"] == steps and d["higher_is_better"] is True
    assert set(d["roofline"]) >= {"bound", "achieved", "peak", "unit", "frac", "traffic"}
    assert set(d["e2e"]) >= {"value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"}
    assert "workload" in d["config"] and d["e2e"]["h2d_bytes_per_step"] > 0
    return d


def test_single_rank_line():
    p = subprocess.run([sys.executable, "-c", CODE, "--workload", "band", "--steps", "4", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    d = _check_line(p.stdout, 1, 4)
    assert "cpu_baseline" in d and d["cpu_baseline"]["cores"] == 1 and d["cpu_baseline"]["kind"] in ("reference", "port")
    assert d["check"]["max_err_over_tolerance"] <= 1.0 and d["scaling"] == "weak"
    assert set(d["config"]) == {"workload", "rows", "cols", "nnz"}


def test_default_workload_is_the_target_matrix_strong_scaling_with_also_block():
    """The default run at reduced scale: uniform 16 nnz/row, strong scaling, and the `also` block with the Laplacian and
    the R-MAT workload (reduced through --also at a small scale is not possible: the block is checked for presence on the
    band-size run below instead)."""
    p = subprocess.run([sys.executable, "-c", CODE, "--scale", "14", "--steps", "3", "--warmup", "3", "--also", "band"], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert p.returncode == 0, p.stderr[-2000:]
    d = _check_line(p.stdout, 1, 3)
    assert d["scaling"] == "strong" and d["config"]["nnz"] == 16 << 14 and d["dtype"] == "f64"
    assert "band" in d["also"] and d["also"]["band"]["check"]["max_err_over_tolerance"] <= 1.0
    assert d["engine"]["device_layout"]["cdb"] in (16384, 32768)


def test_reference_arm_prints_the_same_config_without_loading_the_engine():
    env = dict(os.environ)
    code = ("import sys, json; sys.argv = ['bench.py', '--impl', 'reference', '--scale', '14', '--steps', '2', '--warmup', '1'];"
            "sys.path.insert(0, %r); import bench; bench.main();"
            "maps = open('/proc/self/maps').read(); assert 'libspmvb.so' not in maps, 'engine library loaded'; "
            "assert 'libmatgen.so' in maps") % ROOT
    p = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600, env=env)
    assert p.returncode == 0, p.stderr[-2000:]
    d = json.loads([l for l in p.stdout.splitlines() if l.strip()][0])
    assert d["impl"] == "reference" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["cpu_baseline"]["value"] == d["value"]
    g = subprocess.run([sys.executable, "-c", CODE, "--scale", "14", "--steps", "3", "--warmup", "3", "--also", ""], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert g.returncode == 0, g.stderr[-2000:]
    assert json.loads(g.stdout.strip())["config"] == d["config"]      # both arms: the identical config dict


def test_two_rank_line_over_gloo():
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", CODE,
                        "--gpus", "2", "--workload", "laplacian", "--steps", "4", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _check_line(p.stdout, 2, 4)
    assert d["config"]["rows"] == 2 * 4194304            # weak scaling: the config-2 Laplacian per rank
    # every rank uploads only the band of x its rows read, not the whole replicated vector
    assert d["e2e"]["h2d_bytes_per_step"] < 1.2 * d["config"]["cols"] * 8
    assert "cpu_baseline" not in d and "also" not in d   # N = 1 only


def test_two_rank_strong_scaling_of_the_default_workload_over_gloo():
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", CODE,
                        "--gpus", "2", "--scale", "14", "--steps", "3", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _check_line(p.stdout, 2, 3)
    assert d["scaling"] == "strong" and d["config"]["rows"] == 1 << 14 and d["config"]["nnz"] == 16 << 14
    assert d["engine"]["rows_per_gpu"] == [1 << 13, 1 << 13]
    # x is replicated over the GPU links: every rank uploads 1/N of it, so it crosses the host links once in total
    assert d["e2e"]["h2d_bytes_per_step"] == (1 << 14) * 8 and "ncclAllGather" in d["e2e"]["what"]


def test_power_iteration_line_over_gloo():
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", CODE,
                        "--gpus", "2", "--workload", "poweriter", "--scale", "14", "--steps", "3", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _check_line(p.stdout, 2, 3)
    assert d["dtype"] == "f32" and d["config"]["rows"] == 1 << 14 and d["scaling"] == "strong"
    assert d["check"]["x_identical_on_all_ranks"] is True and d["check"]["rows_max_err_over_tolerance"] <= 1.0
    assert d["last_norm"] > 0


def test_two_rank_run_refuses_a_wrong_result():
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", CODE,
                        "--gpus", "2", "--scale", "14", "--steps", "3", "--warmup", "3"], cwd=ROOT, capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, DRYRUN_BREAK_RANK="1"))
    assert p.returncode != 0
    assert "multi-GPU result check failed" in p.stderr
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")]   # no JSON line
