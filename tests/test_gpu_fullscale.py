"""GPU parity at BASELINE.json's full sizes, against the oracle's CSR SpMV (gold) with the north-star tolerance
|y - gold| <= 1e-12 (fp64) / 1e-5 (fp32) x row-wise |A||x|: the R-MAT matrix of configs[2] / [4] (scale 24, ~263 M
non-zeros, 38 % empty rows) and the uniform target matrix of configs[3] (2^26 rows x 16 = 1 073 741 824 non-zeros in fp64;
fp32 at 2^25 rows of the same 2^26 columns), both production kernels on the same engine.  These sizes are beyond the
O(blocks x rows) emulation of the reference, so gold + size-independent properties are the checks: the
engine-private device layout (16 384-column blocks, row tiles) is what runs here, the API pieces are untouched."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = {True: 1e-12, False: 1e-5}


def _engine_vs_gold(spmvb, oracle, A, variants=(8, 7)):
    isd = A.is_double
    vt = np.float64 if isd else np.float32
    lay = spmvb.Layout.from_csr(A)
    dp = lay.device_params
    eng = spmvb.Engine(lay, 0)
    lay.free()
    x = np.random.default_rng(7).random(A.cols).astype(vt)
    gold, _ = oracle.spmv_gold_omp(A.rows, A.row_ptr, A.col_ind, A.values, x, isd)
    bound = oracle.abs_ax(A.rows, A.row_ptr, A.col_ind, A.values, x, isd) * TOL[isd] + np.finfo(vt).tiny
    gold = gold.astype(np.float64)
    eng.set_x(x)
    worst = {}
    for v in variants:
        eng.set_variant(v)
        for rep in range(2):  # twice: the second launch starts from a y that already holds a result
            eng.spmv_dev()
            y = eng.get_y().astype(np.float64)
            e = np.abs(y - gold) / bound
            assert np.all(e <= 1.0), "variant %d rep %d: row %d off by %g x the tolerance" % (v, rep, int(np.argmax(e)), e.max())
        worst[v] = float(e.max())
    # linearity, a property that needs no reference: A (2x) = 2 A x.  Scaling by 2 is exact, but rows that live in
    # several column blocks are summed with red.global.add in whatever order the warps arrive, so two launches may
    # round differently: the comparison carries the tolerance (twice: both sides are computed results)
    eng.set_variant(0)
    eng.set_x((2 * x).astype(vt))
    eng.spmv_dev()
    eng.set_x(x)
    y2 = eng.get_y().astype(np.float64)
    eng.spmv_dev()
    assert np.all(np.abs(y2 - 2 * eng.get_y().astype(np.float64)) <= 4 * bound)
    # rows without entries come out as exact zeros (empty_rows_bitmap: nothing is accumulated into them)
    empty = np.diff(A.row_ptr.astype(np.int64)) == 0
    assert not eng.get_y()[empty].any()
    info = eng.device_layout
    eng.free()
    return dp, info, worst


@pytest.mark.parametrize("isd", [True, False], ids=["f64", "f32"])
def test_rmat_scale24_both_kernels_match_gold(spmvb, oracle, isd):
    A = spmvb.Csr.rmat(24, 16, 0.57, 0.19, 0.19, 1, 0, 0, isd)
    assert A.rows == 1 << 24 and 250e6 < A.nnz < 270e6
    dp, info, worst = _engine_vs_gold(spmvb, oracle, A)
    print("\nR-MAT scale 24 %s: device layout %r, worst error / tolerance per kernel %r" % ("fp64" if isd else "fp32", info, worst))
    if isd:
        assert dp["private"] and dp["cdb"] == 16384   # fp64 x slice of a 32 768-column block does not fit the window


@pytest.mark.parametrize("isd,log2_rows", [(True, 26), (False, 25)], ids=["f64_2^26rows", "f32_2^25rows"])
def test_uniform_target_matrix_both_kernels_match_gold(spmvb, oracle, isd, log2_rows):
    A = spmvb.Csr.uniform(1 << 26, 1 << 26, 16, 1, 0, 1 << log2_rows, isd)
    assert A.nnz == 16 << log2_rows
    dp, info, worst = _engine_vs_gold(spmvb, oracle, A)
    print("\nuniform 2^%d rows x 16 %s: device layout %r, worst error / tolerance per kernel %r"
          % (log2_rows, "fp64" if isd else "fp32", info, worst))
    assert dp["private"] and dp["cu"] > 1 and dp["cu_major"]   # y does not fit the L2 cache: row tiles
    assert dp["cdb"] == (16384 if isd else 32768)
