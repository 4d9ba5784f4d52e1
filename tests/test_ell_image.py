"""The sliced-ELLPACK image (engine-private, for regular matrices): 32 consecutive rows per slice, one row
per lane, 16-bit slice-relative columns.  On the CPU: every slot against the CSR (real entries in CSR order, padding =
the row's first column with value 0), the eligibility rules, and that the API pieces do not change.  Its kernel and the
end-to-end pipeline over it are tested on the GPU (test_gpu_parity.py)."""
import numpy as np
import pytest

import matgen


def check_against_csr(lay, rows, rp, ci, va, is_double):
    e = lay.ell_params
    assert e["present"] and e["nnz"] == len(ci)
    lens = np.diff(rp.astype(np.int64))
    w = e["width"]
    assert w == lens.max() and e["slices"] == (rows + 31) // 32
    assert e["slice_bytes"] == 16 + w * 64 + w * 32 * (8 if is_double else 4) and e["bytes"] == e["slices"] * e["slice_bytes"]
    cols, vals = lay.ell_decode()
    assert cols.shape == (e["slices"] * 32, w)
    va = va.astype(np.float64 if is_double else np.float32)
    for r in range(rows):
        n, j = int(lens[r]), int(rp[r])
        assert np.array_equal(cols[r, :n], ci[j:j + n]) and np.array_equal(vals[r, :n], va[j:j + n]), r
        assert not vals[r, n:].any()
        if n:
            assert np.all(cols[r, n:] == ci[j]), r       # padding reads a column the row reads anyway
    assert not vals[rows:].any()


@pytest.mark.parametrize("is_double", [True, False])
@pytest.mark.parametrize("name,gen", [("lap", lambda: matgen.laplacian2d(70, 45)), ("band", lambda: matgen.band(1000, 3, seed=2)),
                                      ("lap_wide", lambda: matgen.laplacian2d(700, 30)), ("tiny", lambda: matgen.band(5, 1, seed=3))])
def test_ell_image_holds_the_matrix(spmvb, name, gen, is_double):
    rows, cols, rp, ci, va = gen()
    with spmvb.options(ell=1):
        lay = spmvb.Layout.build(rows, cols, rp, ci, va.astype(np.float64 if is_double else np.float32), 1, 1, is_double)
    check_against_csr(lay, rows, rp, ci, va, is_double)


def test_ell_image_eligibility(spmvb):
    # regular: built without being asked for
    rows, cols, rp, ci, va = matgen.laplacian2d(128, 128)
    assert spmvb.Layout.build(rows, cols, rp, ci, va).ell_params["present"]
    with spmvb.options(ell=0):
        assert not spmvb.Layout.build(rows, cols, rp, ci, va).ell_params["present"]
    # columns of 32 consecutive rows more than 65 535 apart: the 16-bit offsets cannot hold them
    rows, cols, rp, ci, va = matgen.uniform(2000, 200000, 8, seed=1)
    with spmvb.options(ell=1):
        assert not spmvb.Layout.build(rows, cols, rp, ci, va).ell_params["present"]
    # a row longer than 16 entries
    rows, cols, rp, ci, va = matgen.band(500, 9, seed=1)
    with spmvb.options(ell=1):
        assert not spmvb.Layout.build(rows, cols, rp, ci, va).ell_params["present"]
    # ragged rows inside a narrow band: the padding would cost more than 4 % - only on request
    rows, cols, rp, ci, va = matgen.ragged(3000, 4000, seed=5, max_len=12, empty_frac=0.3)
    assert not spmvb.Layout.build(rows, cols, rp, ci, va).ell_params["present"]
    with spmvb.options(ell=1):
        lay = spmvb.Layout.build(rows, cols, rp, ci, va)
    check_against_csr(lay, rows, rp, ci, va, True)   # unsorted columns, empty rows


def test_ell_image_leaves_the_api_pieces_alone(spmvb):
    rows, cols, rp, ci, va = matgen.laplacian2d(100, 60)
    with spmvb.options(ell=0):
        plain = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 2, True)
    withe = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 2, True)
    assert withe.ell_params["present"] and not plain.ell_params["present"]
    assert plain.difference(withe) == ""
