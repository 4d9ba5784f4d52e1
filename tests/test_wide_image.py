"""The wide image (engine-private, next to the API pieces): column blocks of up to 2^23 columns whose
x range the L2 cache holds, rows ascending through a block, chunks stored plane by plane.  On the CPU: the image holds
exactly the CSR's entries ordered by (column block, row, CSR position) - the library's own walk (planes, end-of-row
bits, row map, chunk metadata) is checked on the way - it leaves the API pieces untouched, and it is only built when
it is wanted.  The kernel that streams it is tested on the GPU (test_gpu_parity.py, test_gpu_fullscale.py)."""
import numpy as np
import pytest

import matgen


def expected_order(rows, row_ptr, col_ind, values, cdb):
    row_of = np.repeat(np.arange(rows, dtype=np.uint32), np.diff(row_ptr.astype(np.int64)))
    order = np.argsort(col_ind // cdb, kind="stable")  # CSR is row-major: a stable sort by block gives (block, row, CSR position)
    return row_of[order], col_ind[order], values[order]


CASES = {
    "uniform_tall": lambda: matgen.uniform(3000, 200000, 9, seed=3, empty_frac=0.1),
    "uniform_unsorted": lambda: matgen.uniform(500, 70000, 12, seed=4, sort_cols=False),
    "ragged": lambda: matgen.ragged(800, 150000, seed=5, max_len=700),
    "rmat": lambda: matgen.rmat(11, ef=8, seed=6),
    "band": lambda: matgen.band(5000, 3, seed=7),
    "one_column": lambda: matgen.uniform(300, 1, 1, seed=8),
}


@pytest.mark.parametrize("name", sorted(CASES))
@pytest.mark.parametrize("is_double", [True, False])
@pytest.mark.parametrize("range_log2", [23, 17, 15, 9, 2])
def test_wide_image_holds_the_matrix_in_block_row_order(spmvb, name, is_double, range_log2):
    rows, cols, rp, ci, va = CASES[name]()
    va = va.astype(np.float64 if is_double else np.float32)
    with spmvb.options(wide=1, wide_range_log2=range_log2):
        lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, is_double)
    w = lay.wide_params
    cdb = 1 << range_log2
    assert w["present"] and w["cdb"] == cdb and w["nnz"] == len(ci)
    assert w["blocks"] == (cols + cdb - 1) // cdb
    r, c, v = lay.wide_decode()
    er, ec, ev = expected_order(rows, rp, ci, va, cdb)
    assert np.array_equal(r, er)
    assert np.array_equal(c, ec)
    assert np.array_equal(v.view(np.uint64 if is_double else np.uint32), ev.view(np.uint64 if is_double else np.uint32))
    # pairs = non-empty (row, block) combinations
    assert w["pairs"] == len(np.unique(er.astype(np.int64) * (w["blocks"] + 1) + ec // cdb))
    assert w["bytes"] == w["chunks"] * (768 + 256 * (8 if is_double else 4))


def test_wide_image_leaves_the_api_pieces_alone(spmvb):
    rows, cols, rp, ci, va = matgen.uniform(2000, 100000, 8, seed=11)
    plain = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 2, True)
    with spmvb.options(wide=1):
        withw = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 2, True)
    assert withw.wide_params["present"]
    assert plain.difference(withw) == ""  # every API table and byte


def test_wide_image_only_on_request(spmvb):
    """It lost every measurement on B200: built only with option wide = 1."""
    rows, cols, rp, ci, va = matgen.uniform(4000, 60000, 16, seed=2)
    assert not spmvb.Layout.build(rows, cols, rp, ci, va).wide_params["present"]
    with spmvb.options(wide=0):
        assert not spmvb.Layout.build(rows, cols, rp, ci, va).wide_params["present"]
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Layout.build(rows, cols, rp, ci, va).wide_decode()
    with spmvb.options(wide=1):
        lay = spmvb.Layout.build(rows, cols, rp, ci, va)
    assert lay.wide_params["present"] and lay.wide_params["zero_rows"] >= -1
