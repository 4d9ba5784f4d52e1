"""GPU parity tests: the CUDA engine (through the C ABI) against the oracle on the same seeded inputs.

Tolerances are the north-star's: |y - y_ref| <= 1e-12 (fp64) / 1e-5 (fp32) x row-wise |A||x|.
"""
import numpy as np
import pytest

import matgen
import oracle_api as oa

pytestmark = pytest.mark.gpu

TOL = {True: 1e-12, False: 1e-5}

CASES = {
    "kat6x6": lambda: (6, 6, np.array([0, 3, 4, 4, 6, 7, 9]), np.array([0, 2, 5, 1, 0, 3, 4, 0, 5]),
                       np.arange(1, 10, dtype=float)),
    "band10k": lambda: matgen.band(10000, 5, seed=1),             # BASELINE config 1
    "lap256": lambda: matgen.laplacian2d(256, 256),                # 2 column blocks
    "lap_wide": lambda: matgen.laplacian2d(700, 150),              # 4 column blocks, partial last block
    "ragged": lambda: matgen.ragged(5000, 100000, seed=7),         # empty rows, unsorted columns, 4 blocks
    "uniform": lambda: matgen.uniform(4000, 200000, 16, seed=3),   # ~1 entry per (row, block) pair
    "rmat13": lambda: matgen.rmat(13, 8, seed=5),                  # power law, long rows, empty rows
    "longrow": lambda: matgen.uniform(40, 30000, 9000, seed=9),    # rows spanning many chunks / warps
    "onerow": lambda: matgen.uniform(1, 5000, 3000, seed=11),
}

CONFIGS = [(1, 1, True), (1, 1, False), (2, 2, True), (8, 4, True), (8, 4, False), (12, 8, False), (4, 1, True)]


def _check(spmvb, oracle, M, cu, vf, isd, variant, cdb=0):
    rows, cols, rp, ci, va = M
    vt = oa.vdtype(isd)
    va = va.astype(vt)
    rng = np.random.default_rng(1234)
    x = rng.random(cols).astype(vt)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd, cdb)
    eng = spmvb.Engine(lay, 0, variant)
    y = np.zeros(rows, vt)
    eng.spmv_host(x, y, accumulate=True)
    gold = oracle.spmv_gold(rows, rp, ci, va, x, isd)
    scale = oracle.abs_ax(rows, rp, ci, va, x, isd)
    ho = oracle.build(rows, cols, rp, ci, va, cu, vf, isd, cdb)
    y_emu = np.zeros(rows, vt)
    assert oracle.spmv_emu(ho, x, y_emu, isd) == 0
    oracle.free(ho)
    tol = TOL[isd]
    tiny = np.finfo(vt).tiny
    err_gold = np.abs(y.astype(np.float64) - gold.astype(np.float64))
    err_emu = np.abs(y.astype(np.float64) - y_emu.astype(np.float64))
    assert np.all(err_gold <= tol * scale + tiny), "vs gold: max ratio %g" % np.max(err_gold / (scale + tiny))
    assert np.all(err_emu <= tol * scale + tiny), "vs emu: max ratio %g" % np.max(err_emu / (scale + tiny))
    # spmv_hw accumulates into y_fpga (csr_hw.cpp:1557): a second call doubles the result
    eng.spmv_host(x, y, accumulate=True)
    assert np.all(np.abs(y.astype(np.float64) - 2 * gold.astype(np.float64)) <= 4 * tol * scale + tiny)
    assert eng.launches >= 2
    eng.free()
    lay.free()


@pytest.mark.parametrize("variant", [0, 1, 7, 8])
@pytest.mark.parametrize("cfg", CONFIGS, ids=lambda c: "cu%d_vf%d_%s" % (c[0], c[1], "f64" if c[2] else "f32"))
@pytest.mark.parametrize("case", sorted(CASES))
def test_spmv_matches_oracle(spmvb, oracle, case, cfg, variant):
    cu, vf, isd = cfg
    _check(spmvb, oracle, CASES[case](), cu, vf, isd, variant)


@pytest.mark.parametrize("variant", [0, 1, 7, 8])
def test_small_column_blocks(spmvb, oracle, variant):
    """cols_div_blocks = 16384 (the reference's CU=10/12 setting) and a tiny block width: many blocks."""
    _check(spmvb, oracle, matgen.ragged(3000, 50000, seed=2), 1, 1, True, variant, cdb=16384)
    _check(spmvb, oracle, matgen.uniform(3000, 9000, 12, seed=4), 2, 2, True, variant, cdb=256)


@pytest.mark.parametrize("range_log2", [23, 16, 12])
@pytest.mark.parametrize("isd", [True, False], ids=["f64", "f32"])
@pytest.mark.parametrize("case", sorted(CASES))
def test_wide_image_kernel_matches_oracle(spmvb, oracle, case, isd, range_log2):
    """The wide image (column blocks of 2^range_log2 columns, rows ascending through a block) and its kernel (variant
    9): one block, a few blocks, many blocks; the API pieces next to it are the reference's as ever."""
    with spmvb.options(wide=1, wide_range_log2=range_log2):
        _check(spmvb, oracle, CASES[case](), 1, 1, isd, 9)


@pytest.mark.parametrize("opts", [dict(occ_run_log2=1), dict(occ_run_log2=5), dict(wide_hints=3), dict(wide_hints=1),
                                  dict(run_log2=3, zero_all=1)], ids=lambda o: ",".join("%s=%d" % kv for kv in o.items()))
def test_wide_image_kernel_options(spmvb, oracle, opts):
    with spmvb.options(wide=1, wide_range_log2=14, **opts):
        _check(spmvb, oracle, matgen.uniform(6000, 100000, 12, seed=8), 8, 4, True, 9)
        _check(spmvb, oracle, matgen.rmat(13, 8, seed=5), 2, 2, False, 9)
        _check(spmvb, oracle, matgen.uniform(40, 30000, 9000, seed=9), 1, 1, True, 9)


def test_wide_image_device_accumulate_and_variant_rules(spmvb, oracle):
    rows, cols, rp, ci, va = matgen.ragged(5000, 100000, seed=7)
    x = np.random.default_rng(0).random(cols)
    gold = oracle.spmv_gold(rows, rp, ci, va, x, True)
    scale = oracle.abs_ax(rows, rp, ci, va, x, True) + 1e-300
    with spmvb.options(wide=1, wide_range_log2=15):
        lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
    eng = spmvb.Engine(lay, 0, 9)
    assert eng.variant == 9 and eng.device_layout["wide"] and eng.device_layout["cdb"] == 32768
    eng.set_x(x)
    eng.spmv_dev()
    y1 = eng.get_y()
    eng.spmv_dev(accumulate=True)   # every update an atomic on top of the first result
    y2 = eng.get_y()
    assert np.all(np.abs(y1 - gold) <= 1e-12 * scale)
    assert np.all(np.abs(y2 - 2 * gold) <= 4e-12 * scale)
    with pytest.raises(spmvb.SpmvbError):
        eng.set_variant(7)            # a wide image has one kernel
    eng.free()
    plain = spmvb.Engine(lay, 0, 7)
    with pytest.raises(spmvb.SpmvbError):
        plain.set_variant(9)
    plain.free()
    with spmvb.options(wide=0):
        nowide = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Engine(nowide, 0, 9)
    # the engine's own choice (variant 0): whichever candidate it measured fastest, the result is the same
    auto = spmvb.Engine(lay, 0)
    y = np.zeros(rows)
    auto.spmv_host(x, y, accumulate=False)
    assert np.all(np.abs(y - gold) <= 1e-12 * scale)
    assert auto.device_layout["tuned_us"]["wide_image"] > 0
    auto.free()


ELL_CASES = ["kat6x6", "band10k", "lap256", "lap_wide"]   # rows of <= 16 entries whose columns stay close together


@pytest.mark.parametrize("tiles", [-1, 1, 3, 64])
@pytest.mark.parametrize("cfg", [(1, 1, True), (1, 1, False), (8, 4, True)], ids=lambda c: "cu%d_vf%d_%s" % (c[0], c[1], "f64" if c[2] else "f32"))
@pytest.mark.parametrize("case", ELL_CASES)
def test_ell_image_kernel_matches_oracle(spmvb, oracle, case, cfg, tiles):
    """The sliced-ELLPACK image of a regular matrix and its kernel (variant 10), through the end-to-end pipeline of
    spmv_host (x pieces up / kernel per row tile / y tiles down, overlapped) with several tile counts."""
    cu, vf, isd = cfg
    with spmvb.options(ell=1, ell_tiles=tiles):
        _check(spmvb, oracle, CASES[case](), cu, vf, isd, 10)


def test_ell_image_device_calls_and_variant_rules(spmvb, oracle):
    rows, cols, rp, ci, va = matgen.laplacian2d(300, 200)
    x = np.random.default_rng(0).random(cols)
    gold = oracle.spmv_gold(rows, rp, ci, va, x, True)
    scale = oracle.abs_ax(rows, rp, ci, va, x, True) + 1e-300
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
    eng = spmvb.Engine(lay, 0, 10)
    info = eng.device_layout
    assert eng.variant == 10 and info["ell"] and info["ell_width"] == 5 and info["zero_rows"] == 0
    eng.set_x(x)
    eng.spmv_dev()
    y1 = eng.get_y()
    eng.spmv_dev(accumulate=True)
    y2 = eng.get_y()
    assert np.all(np.abs(y1 - gold) <= 1e-12 * scale)
    assert np.all(np.abs(y2 - 2 * gold) <= 4e-12 * scale)
    with pytest.raises(spmvb.SpmvbError):
        eng.set_variant(7)
    # a short x is zero padded (csr_hw.cpp:1478-1481), also through the pipeline
    xs = x[: cols - 1000]
    xfull = np.zeros(cols); xfull[: len(xs)] = xs
    y = np.zeros(rows)
    eng.spmv_host(xs, y, accumulate=False)
    g2 = oracle.spmv_gold(rows, rp, ci, va, xfull, True)
    assert np.all(np.abs(y - g2) <= 1e-12 * scale)
    eng.spmv_host(x, y, accumulate=False)       # and a full one after it replaces every column
    assert np.all(np.abs(y - gold) <= 1e-12 * scale)
    nrm = None
    eng.set_x(np.full(cols, 1.0 / np.sqrt(cols)))
    nrm = eng.power_iter(3)
    assert nrm > 0
    eng.free()
    # the engine's own choice on a regular matrix: whichever it measured fastest gives the same result
    auto = spmvb.Engine(lay, 0)
    assert auto.device_layout["tuned_us"]["ell_image"] > 0
    y = np.zeros(rows)
    auto.spmv_host(x, y, accumulate=False)
    assert np.all(np.abs(y - gold) <= 1e-12 * scale)
    auto.free()
    rows, cols, rp, ci, va = matgen.uniform(2000, 200000, 8, seed=1)
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Engine(spmvb.Layout.build(rows, cols, rp, ci, va), 0, 10)   # not a regular matrix: no ELL image


def test_device_api_and_determinism_of_inputs(spmvb, oracle):
    rows, cols, rp, ci, va = matgen.laplacian2d(512, 512)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
    eng = spmvb.Engine(lay, 0)
    x = np.random.default_rng(0).random(cols)
    eng.set_x(x)
    eng.spmv_dev()
    y1 = eng.get_y()
    eng.spmv_dev(accumulate=True)
    y2 = eng.get_y()
    gold = oracle.spmv_gold(rows, rp, ci, va, x, True)
    scale = oracle.abs_ax(rows, rp, ci, va, x, True)
    assert np.all(np.abs(y1 - gold) <= 1e-12 * scale)
    assert np.all(np.abs(y2 - 2 * gold) <= 4e-12 * scale)
    ms = eng.time_spmv(3, flush_l2=True)
    assert np.all(ms > 0)


def test_linearity_at_scale(spmvb):
    """Size-independent property on a matrix too big for the O(blocks x rows) oracle: A(ax+bz) = aAx + bAz."""
    A = spmvb.Csr.laplacian2d(2048, 2048)  # BASELINE config 2: 4 194 304 rows, 20 963 328 nnz
    assert A.nnz == 20963328
    lay = spmvb.Layout.from_csr(A)
    eng = spmvb.Engine(lay, 0)
    rng = np.random.default_rng(5)
    x, z = rng.random(A.cols), rng.random(A.cols)
    y = {}
    for name, v in (("x", x), ("z", z), ("c", 2.0 * x - 3.0 * z)):
        out = np.zeros(A.rows)
        eng.spmv_host(v, out, accumulate=False)
        y[name] = out
    # |A||v| <= 8 * max|v| for this operator
    assert np.max(np.abs(y["c"] - (2.0 * y["x"] - 3.0 * y["z"]))) <= 1e-12 * 8 * 5
    # interior rows of the Laplacian applied to a constant vector vanish; row sums are known exactly
    ones = np.ones(A.cols)
    out = np.zeros(A.rows)
    eng.spmv_host(ones, out, accumulate=False)
    deg = np.diff(A.row_ptr.astype(np.int64)) - 1
    assert np.array_equal(out, 4.0 - deg)


def test_power_iteration_matches_numpy(spmvb, oracle):
    rows, cols, rp, ci, va = matgen.rmat(11, 8, seed=3)
    va = np.abs(va).astype(np.float32)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, False)
    eng = spmvb.Engine(lay, 0)
    x0 = np.full(cols, 1.0 / np.sqrt(cols), np.float32)
    eng.set_x(x0)
    nrm = eng.power_iter(20)
    x = x0.astype(np.float64)
    for _ in range(20):
        yv = oracle.spmv_gold(rows, rp, ci, va.astype(np.float64), x, True)
        n = np.linalg.norm(yv)
        x = yv / n
    assert abs(nrm - n) <= 1e-3 * n


@pytest.mark.parametrize("variant", [0, 1, 7, 8, 10])
def test_config2_scale_every_variant_repeated(spmvb, oracle, variant):
    """BASELINE config 2 at full size against the gold CSR SpMV, several launches per variant: every warp walks
    many chunks here, which is what exposes pipeline races that the small cases cannot."""
    A = spmvb.Csr.laplacian2d(2048, 2048)
    lay = spmvb.Layout.from_csr(A)
    eng = spmvb.Engine(lay, 0, variant)
    x = np.random.default_rng(5).random(A.cols)
    gold = oracle.spmv_gold(A.rows, A.row_ptr, A.col_ind, A.values, x, True)
    eng.set_x(x)
    for rep in range(6):
        eng.spmv_dev()
        y = eng.get_y()
        assert np.max(np.abs(y - gold)) <= 1e-12 * 8.0, "variant %d rep %d" % (variant, rep)


@pytest.mark.parametrize("variant", [7, 8])
def test_cu_major_device_order(spmvb, oracle, variant):
    """CU-major order of the pieces on the device (used when y does not fit the L2 cache): same results."""
    # cu_major + the explicit L2 eviction policies of the tall-matrix path
    with spmvb.options(cu_major=1, tall=1):
        _check(spmvb, oracle, matgen.uniform(6000, 100000, 12, seed=8), 4, 1, True, variant)
        _check(spmvb, oracle, matgen.ragged(5000, 100000, seed=7), 8, 4, False, variant)
        _check(spmvb, oracle, matgen.laplacian2d(256, 256), 2, 2, True, variant, cdb=16384)


@pytest.mark.parametrize("variant", [0, 7, 8])
def test_engine_private_device_layout(spmvb, oracle, variant):
    """What the GPU streams is the engine's choice (row tiles, narrower column blocks, no VF padding) while the API
    pieces stay what CU / VF / COLS_DIV_BLOCKS say: same y as the emulated spmv_hw of the API layout, within tolerance
    (only the association of the per-block partial sums differs)."""
    with spmvb.options(dev_tiles=3, dev_cdb=8192):
        _check(spmvb, oracle, matgen.uniform(6000, 100000, 12, seed=8), 1, 1, True, variant)
        _check(spmvb, oracle, matgen.rmat(13, 8, seed=5), 8, 4, False, variant)
        _check(spmvb, oracle, matgen.ragged(5000, 100000, seed=7), 2, 2, True, variant)
    with spmvb.options(dev_tiles=5, tall=1):       # CU-major tiles + the explicit L2 policies
        _check(spmvb, oracle, matgen.uniform(40, 30000, 9000, seed=9), 1, 1, True, variant)   # rows spanning many chunks
        _check(spmvb, oracle, matgen.laplacian2d(256, 256), 4, 1, True, variant)
    with spmvb.options(e2e_tiles=0, dev_tiles=4):  # spmv_host without the tile pipeline
        _check(spmvb, oracle, matgen.uniform(4000, 200000, 16, seed=3), 1, 1, True, variant)
    lay = spmvb.Layout.build(*matgen.uniform(4000, 200000, 16, seed=3), 1, 1, True)
    dp = lay.device_params
    assert dp["private"] and dp["cdb"] == 16384 and lay.n_cu == 1 and lay.blocks == 7   # API: 32 768-column blocks


@pytest.mark.parametrize("variant", [7, 8])
def test_x_upload_skips_untouched_column_blocks(spmvb, oracle, variant):
    """set_x / spmv_host copy only the column ranges the matrix can read (two ranges here, block 1 and 4 untouched):
    same result, fewer bytes; a second x replaces the first everywhere it matters."""
    rng = np.random.default_rng(3)
    cdb, rows, cols = 4096, 3000, 5 * 4096
    rp, ci = [0], []
    for _ in range(rows):
        c = np.concatenate([rng.integers(0, cdb, 2), rng.integers(2 * cdb, 4 * cdb, 3)])
        ci.extend(sorted(int(v) for v in set(c.tolist())))
        rp.append(len(ci))
    rp = np.array(rp, np.uint64); ci = np.array(ci, np.uint32)
    va = rng.uniform(-1, 1, len(ci))
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True, cdb)
    eng = spmvb.Engine(lay, 0, variant)
    assert eng.x_upload_bytes == 3 * cdb * 8
    for seed in (1, 2):
        x = np.random.default_rng(seed).random(cols)
        y = np.zeros(rows)
        eng.spmv_host(x, y, accumulate=False)
        gold = oracle.spmv_gold(rows, rp, ci, va, x, True)
        scale = oracle.abs_ax(rows, rp, ci, va, x, True)
        assert np.all(np.abs(y - gold) <= 1e-12 * scale + 1e-300)
    # a short x (fewer values than columns) is zero padded: entries beyond it contribute nothing
    xs = np.random.default_rng(5).random(3 * cdb - 100)
    y = np.zeros(rows)
    eng.spmv_host(xs, y, accumulate=False)
    xfull = np.zeros(cols); xfull[: len(xs)] = xs
    gold = oracle.spmv_gold(rows, rp, ci, va, xfull, True)
    scale = oracle.abs_ax(rows, rp, ci, va, np.abs(xfull) + 1e-3, True)
    assert np.all(np.abs(y - gold) <= 1e-12 * scale + 1e-300)
