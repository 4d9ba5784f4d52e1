"""GPU tests of the GPU layout builder (spmvb_engine_create_from_csr): the image the CUDA kernels build equals the host
builder's - every table and byte - and an engine created that way passes the same SpMV parity check as an uploaded one.
(The host builder is pinned to the reference by tests/test_layout_builder.py and the golden fixtures.)"""
import time

import numpy as np
import pytest

import matgen
import oracle_api as oa

pytestmark = pytest.mark.gpu

TOL = {True: 1e-12, False: 1e-5}

CASES = {
    "kat6x6": lambda: (6, 6, np.array([0, 3, 4, 4, 6, 7, 9]), np.array([0, 2, 5, 1, 0, 3, 4, 0, 5]),
                       np.arange(1, 10, dtype=float)),
    "band10k": lambda: matgen.band(10000, 5, seed=1),
    "lap_wide": lambda: matgen.laplacian2d(700, 150),
    "uniform": lambda: matgen.uniform(4000, 200000, 16, seed=3, empty_frac=0.2),
    "rmat13": lambda: matgen.rmat(13, 8, seed=5),
    "longrow": lambda: matgen.uniform(40, 30000, 9000, seed=9),
    "onerow": lambda: matgen.uniform(1, 5000, 3000, seed=11),
    "ragged_unsorted": lambda: matgen.ragged(5000, 100000, seed=7),   # rows visit their column blocks out of order
    "empty": lambda: (5, 9, np.zeros(6, np.uint64), np.zeros(0, np.uint32), np.zeros(0)),
}
CONFIGS = [(1, 1, True), (1, 2, False), (2, 2, True), (8, 4, True), (8, 8, False), (12, 8, False), (3, 1, True)]


def _parity(oracle, eng, M, isd):
    rows, cols, rp, ci, va = M
    vt = oa.vdtype(isd)
    va = va.astype(vt)
    x = np.random.default_rng(99).random(cols).astype(vt)
    y = np.zeros(rows, vt)
    eng.spmv_host(x, y, accumulate=True)
    gold = oracle.spmv_gold(rows, rp, ci, va, x, isd)
    scale = oracle.abs_ax(rows, rp, ci, va, x, isd)
    err = np.abs(y.astype(np.float64) - gold.astype(np.float64))
    assert np.all(err <= TOL[isd] * scale + np.finfo(vt).tiny)


@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "cu%d_vf%d_%s" % (c[0], c[1], "f64" if c[2] else "f32"))
@pytest.mark.parametrize("case", sorted(CASES))
def test_gpu_built_image_equals_host_built(spmvb, oracle, case, cfg):
    cu, vf, isd = cfg
    M = CASES[case]()
    rows, cols, rp, ci, va = M
    host = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
    lay, eng = spmvb.Engine.from_csr(rows, cols, rp, ci, va, cu, vf, isd)
    with pytest.raises(spmvb.SpmvbError):
        lay.piece_words(0, 0)                      # still on the device
    eng.fetch_layout()
    assert host.difference(lay) == ""
    for variant in (7, 8):
        eng.set_variant(variant)
        _parity(oracle, eng, M, isd)
    eng.free(); lay.free(); host.free()


@pytest.mark.parametrize("cdb", [16384, 256, 12])
def test_gpu_builder_custom_block_width_cu_major(spmvb, oracle, cdb):
    M = matgen.uniform(3000, 40000, 9, seed=4, empty_frac=0.3)
    rows, cols, rp, ci, va = M
    with spmvb.options(cu_major=1):
        host = spmvb.Layout.build(rows, cols, rp, ci, va, 4, 2, True, cdb)
        lay, eng = spmvb.Engine.from_csr(rows, cols, rp, ci, va, 4, 2, True, cdb)
    eng.fetch_layout()
    assert host.difference(lay) == ""
    _parity(oracle, eng, M, True)


def test_gpu_builder_takes_device_resident_csr(spmvb, oracle):
    import torch
    M = matgen.rmat(14, 8, seed=2)
    rows, cols, rp, ci, va = M
    d_rp = torch.from_numpy(np.ascontiguousarray(rp, np.uint64).view(np.int64)).cuda()
    d_ci = torch.from_numpy(np.ascontiguousarray(ci, np.uint32).view(np.int32)).cuda()
    d_va = torch.from_numpy(np.ascontiguousarray(va, np.float64)).cuda()
    torch.cuda.synchronize()
    lay, eng = spmvb.Engine.from_csr(rows, cols, d_rp.data_ptr(), d_ci.data_ptr(), d_va.data_ptr(), 2, 1, True,
                                     on_device=True)
    host = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 1, True)
    eng.fetch_layout()
    assert host.difference(lay) == ""
    _parity(oracle, eng, M, True)
    assert eng.build_ms()["h2d_ms"] < eng.build_ms()["total_ms"]


def test_gpu_builder_rejects_what_it_cannot_do(spmvb):
    rows, cols, rp, ci, va = matgen.band(100)
    bad = ci.copy(); bad[17] = cols + 3
    with pytest.raises(spmvb.SpmvbError, match="out of range"):
        spmvb.Engine.from_csr(rows, cols, rp, bad, va)
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Engine.from_csr(rows, cols, rp, ci, va, 1, 3, True)
    # a host-side Engine cannot be made from a layout whose image was never fetched
    lay, eng = spmvb.Engine.from_csr(rows, cols, rp, ci, va)
    with pytest.raises(spmvb.SpmvbError, match="fetch"):
        spmvb.Engine(lay, 0)


def test_gpu_builder_at_config2_scale(spmvb, oracle):
    """BASELINE config 2 (2048 x 2048 Laplacian, 4 M rows, 21 M nnz): identical image, SpMV parity, and the time."""
    A = spmvb.Csr.laplacian2d(2048, 2048)
    rows, cols = A.rows, A.cols
    rp, ci, va = A.row_ptr, A.col_ind, A.values
    t0 = time.perf_counter()
    host = spmvb.Layout.from_csr(A)
    t_host = time.perf_counter() - t0
    lay, eng = spmvb.Engine.from_csr(rows, cols, rp, ci, va)
    ms = eng.build_ms()
    eng.fetch_layout()
    assert host.difference(lay) == ""
    # a regular matrix: the engine also builds the sliced-ELLPACK image on the GPU and keeps whichever kernel it measured faster
    assert eng.variant in (7, 10) and eng.device_layout["tuned_us"]["ell_image"] > 0
    x = np.random.default_rng(5).random(cols)
    y = np.zeros(rows)
    eng.spmv_host(x, y, accumulate=True)
    gold = oracle.spmv_gold(rows, rp, ci, va, x, True)
    scale = oracle.abs_ax(rows, rp, ci, va, x, True)
    assert np.all(np.abs(y - gold) <= 1e-12 * scale + 1e-300)
    print("\nlayout build at config 2: host builder %.1f ms; GPU: upload %.1f ms + kernels %.2f ms (call %.1f ms)"
          % (t_host * 1e3, ms["h2d_ms"], ms["build_ms"], ms["total_ms"]))


@pytest.mark.parametrize("cu", [1, 8])
def test_gpu_builder_rmat_scale20(spmvb, cu):
    """Power-law matrix with many empty rows and rows that span hundreds of chunks, 1 M rows / 16 M nnz."""
    A = spmvb.Csr.rmat(20, 16, seed=3)
    host = spmvb.Layout.from_csr(A, cu, 1, 16384)
    lay, eng = spmvb.Engine.from_csr(A.rows, A.cols, A.row_ptr, A.col_ind, A.values, cu, 1, True, 16384)
    eng.fetch_layout()
    assert host.difference(lay) == ""
    print("\nR-MAT scale 20 CU=%d: GPU build kernels %.2f ms" % (cu, eng.build_ms()["build_ms"]))


@pytest.mark.parametrize("isd", [True, False], ids=["f64", "f32"])
@pytest.mark.parametrize("case", ["kat6x6", "band10k", "lap_wide"])
def test_gpu_built_ell_image_equals_host_built(spmvb, oracle, case, isd):
    """Regular matrices: the sliced-ELLPACK image built by CUDA kernels from the device-resident CSR (ell_gpu.cuh) is
    the host builder's, byte for byte; a GPU-built engine keeps it NEXT to the hw_matrix stream, so every kernel
    variant stays available and the API image can still be fetched."""
    M = CASES[case]()
    rows, cols, rp, ci, va = M
    va = va.astype(oa.vdtype(isd))
    with spmvb.options(ell=1):                      # use it whenever the format can hold the matrix
        host = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, isd)
        lay, eng = spmvb.Engine.from_csr(rows, cols, rp, ci, va, 1, 1, isd)
    assert eng.variant == 10 and eng.device_layout["ell"]
    assert np.array_equal(eng.ell_image(), host.ell_image())
    _parity(oracle, eng, M, isd)
    for variant in (7, 8, 10, 0):
        eng.set_variant(variant)
        _parity(oracle, eng, M, isd)
    eng.fetch_layout()
    assert host.difference(lay) == ""
    eng.free(); lay.free(); host.free()
    with spmvb.options(ell=0):
        lay, eng = spmvb.Engine.from_csr(rows, cols, rp, ci, va, 1, 1, isd)
    assert eng.ell_image() is None and eng.variant != 10
    with pytest.raises(spmvb.SpmvbError):
        eng.set_variant(10)
    eng.free(); lay.free()
