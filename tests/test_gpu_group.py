"""GPU tests of the multi-GPU group behind the C ABI (spmvb_group_*, the reference's CU dimension mapped to GPUs:
spmv.cpp:249-294).  World size 1 runs everywhere; the tests that need peers skip on a single-GPU box (they are run with
`gpurun --gpus 2` / 8, see profiles/r2/)."""
import numpy as np
import pytest

import matgen
import oracle_api as oa

pytestmark = pytest.mark.gpu

TOL = {True: 1e-12, False: 1e-5}


def _ngpus():
    import torch
    return torch.cuda.device_count()


def _parity(oracle, grp, M, isd):
    rows, cols, rp, ci, va = M
    vt = oa.vdtype(isd)
    va = va.astype(vt)
    x = np.random.default_rng(11).random(cols).astype(vt)
    y = np.zeros(rows, vt)
    grp.spmv_host(x, y, accumulate=True)
    gold = oracle.spmv_gold(rows, rp, ci, va, x, isd).astype(np.float64)
    bound = oracle.abs_ax(rows, rp, ci, va, x, isd) * TOL[isd] + np.finfo(vt).tiny
    assert np.all(np.abs(y.astype(np.float64) - gold) <= bound)
    grp.spmv_host(x, y, accumulate=True)  # spmv_hw accumulates into y_fpga (csr_hw.cpp:1557)
    assert np.all(np.abs(y.astype(np.float64) - 2 * gold) <= 4 * bound)


@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
@pytest.mark.parametrize("isd", [True, False], ids=["f64", "f32"])
def test_group_spmv_host_matches_the_oracle(spmvb, oracle, n_dev, isd):
    if n_dev > _ngpus():
        pytest.skip("needs %d GPUs" % n_dev)
    for M in (matgen.rmat(13, 8, seed=5), matgen.laplacian2d(300, 200), matgen.uniform(5000, 150000, 16, seed=3)):
        rows, cols, rp, ci, va = M
        grp = spmvb.Group.create(rows, cols, rp, ci, va.astype(oa.vdtype(isd)), isd, devices=range(n_dev))
        assert grp.world == n_dev and grp.local_count == n_dev and grp.bounds[0] == 0 and grp.bounds[-1] == rows
        assert np.all(np.diff(grp.bounds.astype(np.int64)) > 0)
        assert grp.x_over_links == -1     # decided (collectively) at the first spmv_host
        _parity(oracle, grp, M, isd)
        if cols == 150000:  # every shard of the uniform matrix reads all of x: 1/N per PCIe link + all-gather over NVLink
            assert grp.x_over_links == (1 if n_dev > 1 else 0)
            # a short x is zero padded (csr_hw.cpp:1478-1481) on that path as well
            vt = oa.vdtype(isd)
            xs = np.random.default_rng(5).random(cols - 40001).astype(vt)
            xfull = np.zeros(cols, vt); xfull[: len(xs)] = xs
            y = np.zeros(rows, vt)
            grp.spmv_host(xs, y, accumulate=False)
            gold = oracle.spmv_gold(rows, rp, ci, va.astype(vt), xfull, isd).astype(np.float64)
            bound = oracle.abs_ax(rows, rp, ci, va.astype(vt), np.abs(xfull) + 1e-3, isd) * TOL[isd] + np.finfo(vt).tiny
            assert np.all(np.abs(y.astype(np.float64) - gold) <= bound)
        grp.free()


def test_group_around_an_engine_of_the_caller(spmvb, oracle):
    """spmvb_group_adopt_engine: a group rank around an engine the caller made and keeps; y_rows = the local rows."""
    rows, cols, rp, ci, va = matgen.uniform(3000, 90000, 12, seed=2)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
    eng = spmvb.Engine(lay, 0)
    grp = spmvb.Group.adopt(eng, rows, cols, [0, rows], 0, None, 0, 1)
    x = np.random.default_rng(3).random(cols)
    y = np.zeros(rows)
    grp.spmv_host_rows(x, y, accumulate=False)
    gold = oracle.spmv_gold(rows, rp, ci, va, x, True)
    assert np.all(np.abs(y - gold) <= 1e-12 * oracle.abs_ax(rows, rp, ci, va, x, True) + 1e-300)
    assert grp.x_over_links == 0          # one GPU: nothing to replicate
    grp.free()
    eng.spmv_host(x, y, accumulate=False)  # the engine outlives the group
    assert np.all(np.abs(y - gold) <= 1e-12 * oracle.abs_ax(rows, rp, ci, va, x, True) + 1e-300)
    eng.free()


@pytest.mark.parametrize("exchange", [0, 1, 2], ids=["nccl_broadcasts", "peer_all", "peer_forward_allgather"])
@pytest.mark.parametrize("n_dev", [1, 2, 4, 8])
def test_group_power_iteration_matches_float64_numpy(spmvb, oracle, n_dev, exchange):
    if n_dev > _ngpus():
        pytest.skip("needs %d GPUs" % n_dev)
    if n_dev == 1 and exchange:
        pytest.skip("one GPU exchanges nothing")
    rows, cols, rp, ci, va = matgen.rmat(12, 8, seed=3)
    va = np.abs(va).astype(np.float32)
    grp = spmvb.Group.create(rows, cols, rp, ci, va, False, devices=range(n_dev))
    if n_dev > 1:
        assert grp.exchange == 2          # peers mapped at creation: the balanced peer-memory exchange is the default
        grp.set_exchange(exchange)
    x0 = np.full(cols, 1.0 / np.sqrt(cols), np.float32)
    grp.set_x(x0)
    nrm = grp.power_iter(20)
    x = x0.astype(np.float64)
    for _ in range(20):
        yv = oracle.spmv_gold(rows, rp, ci, va.astype(np.float64), x, True)
        n = np.linalg.norm(yv)
        x = yv / n
    assert abs(nrm - n) <= 1e-4 * n
    xg = grp.get_x().astype(np.float64)
    assert np.linalg.norm(xg - x) <= 1e-3
    assert grp.last_iter_ms > 0
    # a second call continues from the x the first one left on the devices
    nrm2 = grp.power_iter(1)
    y2 = oracle.spmv_gold(rows, rp, ci, va.astype(np.float64), x, True)
    assert abs(nrm2 - np.linalg.norm(y2)) <= 1e-4 * np.linalg.norm(y2)
    grp.free()


def test_group_rejects_bad_input(spmvb):
    rows, cols, rp, ci, va = matgen.uniform(50, 80, 3, seed=1)
    grp = spmvb.Group.create(rows, cols, rp, ci, va, True, devices=[0])
    with pytest.raises(spmvb.SpmvbError, match="square"):
        grp.power_iter(2)
    grp.free()
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Group.create(rows, cols, rp, ci, va, True, devices=[99])
