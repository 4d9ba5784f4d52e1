"""world_size-2 (and 3) gloo tests of the multi-GPU host logic on CPU: row partition, per-rank CU=1 layouts of row
slices, y-slice all-gather and the iterated caller.  The local SpMV is played by the oracle (there is no GPU here);
on the GPU box bench.py runs the same driver code over the CUDA engine and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import matgen
import oracle_api as oa


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, case, out_dir, mode):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [os.path.join(root, "spmv-fpga_b200"), os.path.join(root, "tests")]
    import host_driver
    import spmvb
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    orc = oa.OracleLib()
    rows, cols, rp, ci, va = matgen.rmat(11, 8, seed=3) if case == "rmat" else matgen.laplacian2d(64, 48)
    va = np.abs(va) + 0.1
    bounds = host_driver.row_bounds(rows, rp, world, 2, balanced=True)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    # this rank's row slice as its own CSR; its CU=1 layout must equal the oracle's layout of the same slice
    j0, j1 = int(rp[lo]), int(rp[hi])
    rp_l = (rp[lo:hi + 1] - rp[lo]).astype(np.uint64)
    ci_l, va_l = ci[j0:j1], va[j0:j1]
    lay = spmvb.Layout.build(hi - lo, cols, rp_l, ci_l, va_l, 1, 1, True)
    ho = orc.build(hi - lo, cols, rp_l, ci_l, va_l, 1, 1, True)
    snap = orc.snapshot(ho, hi - lo, 1, 1, True)
    for b in range(lay.blocks):
        assert lay.piece_info(0, b) == snap.info[(0, b)]
        assert np.array_equal(lay.piece_words(0, b), snap.masked_words(0, b))

    def spmv_local(x_full, y_local):  # the oracle's emulated spmv_hw on this rank's pieces
        y = np.zeros(hi - lo)
        assert orc.spmv_emu(ho, x_full.numpy()[:cols].copy(), y, True) == 0
        y_local[: hi - lo] = torch.from_numpy(y)

    plan = host_driver.GatherPlan(bounds, mode=mode)
    x = torch.full((cols,), 1.0 / np.sqrt(cols), dtype=torch.float64)
    y = torch.zeros(plan.max_len, dtype=torch.float64)
    nrm = host_driver.power_iteration(spmv_local, x, y, plan, 12, dist=dist)
    np.save(os.path.join(out_dir, "x_%d.npy" % rank), x.numpy())
    np.save(os.path.join(out_dir, "n_%d.npy" % rank), np.array([nrm, lo, hi]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,case,mode", [(2, "rmat", "chunks"), (2, "lap", "chunks"), (3, "rmat", "chunks"),
                                             (2, "rmat", "broadcast"), (3, "lap", "broadcast")])
def test_power_iteration_over_row_shards_matches_single_process(tmp_path, world, case, mode):
    port = _free_port()
    mp.spawn(_worker, args=(world, port, case, str(tmp_path), mode), nprocs=world, join=True)
    rows, cols, rp, ci, va = matgen.rmat(11, 8, seed=3) if case == "rmat" else matgen.laplacian2d(64, 48)
    va = np.abs(va) + 0.1
    orc = oa.OracleLib()
    x = np.full(cols, 1.0 / np.sqrt(cols))
    for _ in range(12):
        y = orc.spmv_gold(rows, rp, ci, va, x, True)
        n = np.linalg.norm(y)
        x = y / n
    xs = [np.load(str(tmp_path / ("x_%d.npy" % r))) for r in range(world)]
    ns = [np.load(str(tmp_path / ("n_%d.npy" % r))) for r in range(world)]
    for r in range(world):
        assert np.allclose(xs[r], x, rtol=1e-10, atol=1e-13)   # every rank ends with the same, full x
        assert abs(ns[r][0] - n) <= 1e-10 * n
    # the ranges tile the rows
    assert ns[0][1] == 0 and ns[-1][2] == rows and all(ns[r][2] == ns[r + 1][1] for r in range(world - 1))


def test_row_bounds_fall_back_to_equal_ranges():
    import sys
    import host_driver
    rp = np.arange(0, 11, dtype=np.uint64)  # 10 rows, 1 nnz each
    b = host_driver.row_bounds(10, rp, 4, 2, balanced=True, partition_fn=lambda rows, rp, w, rv: [0, 4, 8, 10, 10])
    assert list(b) == [0, 2, 4, 6, 10]  # a split did not fire -> equal ranges
    assert list(host_driver.row_bounds(10, None, 1)) == [0, 10]


def test_gather_plan_chunk_splits_tile_the_rows():
    import host_driver
    for bounds in ([0, 5, 6, 20], [0, 0, 7, 7, 9], [0, 1000, 1001, 1002, 1003, 2049]):
        plan = host_driver.GatherPlan(bounds, mode="chunks")
        W = plan.world
        send = [[plan._overlap(k, c) for c in range(W)] for k in range(W)]
        for k in range(W):
            assert sum(send[k]) == plan.lens[k]                      # every owned row is sent exactly once
        for c in range(W):
            want = max(0, min((c + 1) * plan.chunk, plan.rows) - c * plan.chunk)
            assert sum(send[k][c] for k in range(W)) == want         # every chunk is filled exactly
