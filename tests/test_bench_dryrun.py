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
    assert "cpu_baseline" in d and d["cpu_baseline"]["cores"] == 1
    assert d["setup_s"]["gpu_layout_build"]["identical_to_host_build"] is True


def test_two_rank_line_over_gloo():
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", CODE,
                        "--gpus", "2", "--steps", "4", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stderr[-3000:]
    d = _check_line(p.stdout, 2, 4)
    assert d["config"]["rows"] == 2 * 4194304            # weak scaling: the config-2 Laplacian per rank
    # every rank uploads only the band of x its rows read, not the whole replicated vector
    assert d["e2e"]["h2d_bytes_per_step"] < 1.2 * d["config"]["cols"] * 8
    assert "cpu_baseline" not in d                       # N = 1 only


@pytest.mark.parametrize("mode", ["broadcast", "chunks"])
def test_power_iteration_line_over_gloo(mode):
    port = _free_port()
    env = dict(os.environ, SPMVB_EXCHANGE=mode)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", CODE,
                        "--gpus", "2", "--workload", "poweriter", "--scale", "14", "--steps", "3", "--warmup", "3"], cwd=ROOT,
                       capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0, p.stderr[-3000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, p.stdout
    d = json.loads(lines[0])
    assert d["n_gpus"] == 2 and d["steps"] == 3 and d["dtype"] == "f32" and d["config"]["rows"] == 1 << 14
    # the stand-in engine returns y = 1 on every rank: after normalisation ||x|| = 1 and the last norm is sqrt(rows)
    assert abs(d["last_norm"] - 128.0) < 1e-3
    assert ("broadcast" in d["config"]["step"]) == (mode == "broadcast")


def test_two_rank_run_refuses_a_wrong_result():
    port = _free_port()
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), "--no-python", sys.executable, "-c", CODE,
                        "--gpus", "2", "--steps", "3", "--warmup", "3"], cwd=ROOT, capture_output=True, text=True,
                       timeout=900, env=dict(os.environ, DRYRUN_BREAK_RANK="1"))
    assert p.returncode != 0
    assert "multi-GPU result check failed" in p.stderr
    assert not [l for l in p.stdout.splitlines() if l.strip().startswith("{")]   # no JSON line
