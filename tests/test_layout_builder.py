"""CPU tests of the product's O(nnz) layout builder (through the C ABI) against the oracle and the golden fixtures:
bit-exact pieces, metadata and empty_rows_bitmap for every CU / VF / precision."""
import os

import numpy as np
import pytest

import matgen
import oracle_api as oa
from test_oracle import GOLDEN, load_fixture


def product_snapshot(lay, cu, vf, isd):
    s = oa.Layout(cu, vf, isd, lay.blocks)
    for b in range(lay.blocks):
        s.bitmap.append(lay.bitmap_row(b))
        for k in range(cu):
            s.info[(k, b)] = lay.piece_info(k, b)
            s.words[(k, b)] = lay.piece_words(k, b)
    return s


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_builder_matches_reference_golden(spmvb, path):
    fx = load_fixture(path)
    for build in (spmvb.Layout.build, spmvb.Layout.build_u32):
        lay = build(fx["rows"], fx["cols"], fx["rp"], fx["ci"], fx["va"], fx["cu"], fx["vf"], fx["isd"])
        assert oa.layouts_equal(fx["layout"], product_snapshot(lay, fx["cu"], fx["vf"], fx["isd"])) == []
        assert lay.expanded_cols == fx["hw_x_len"]
        px = lay.pack_x(fx["x"])
        assert np.array_equal(px[: fx["cols"]], fx["x"]) and not px[fx["cols"]:].any()
        lay.free()


MATS = {
    "band": lambda: matgen.band(5000, 5, seed=1),
    "lap": lambda: matgen.laplacian2d(220, 190),
    "ragged": lambda: matgen.ragged(6000, 120000, seed=21),
    "uniform": lambda: matgen.uniform(2500, 140000, 16, seed=22),
    "rmat": lambda: matgen.rmat(12, 8, seed=23),
    "longrows": lambda: matgen.uniform(30, 70000, 9000, seed=24),
    "unsorted_cols": lambda: matgen.uniform(3000, 90000, 20, seed=25, sort_cols=False),
    "tiny": lambda: matgen.uniform(3, 5, 2, seed=26),
}
CFGS = [(1, 1, True), (1, 2, False), (2, 1, True), (2, 4, False), (4, 2, True), (8, 4, True), (8, 8, False), (10, 2, True),
        (12, 8, False), (3, 1, True)]


@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: "cu%d_vf%d_%s" % (c[0], c[1], "f64" if c[2] else "f32"))
@pytest.mark.parametrize("mat", sorted(MATS))
def test_builder_matches_oracle(spmvb, oracle, mat, cfg):
    cu, vf, isd = cfg
    rows, cols, rp, ci, va = MATS[mat]()
    va = va.astype(oa.vdtype(isd))
    ho = oracle.build(rows, cols, rp, ci, va, cu, vf, isd)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
    assert oa.layouts_equal(oracle.snapshot(ho, rows, cu, vf, isd), product_snapshot(lay, cu, vf, isd)) == []
    assert lay.real_nnz == len(ci)
    assert lay.padded_nnz == sum(lay.piece_info(k, b)[2] for k in range(cu) for b in range(lay.blocks))
    assert lay.pairs == sum(int((lay.bitmap_row(b) == 0).sum()) for b in range(lay.blocks))
    oracle.free(ho); lay.free()


@pytest.mark.parametrize("cdb", [16384, 4096, 256, 64])
def test_builder_custom_column_block_width(spmvb, oracle, cdb):
    rows, cols, rp, ci, va = matgen.ragged(2000, 30000, seed=31)
    for cu, vf, isd in ((1, 1, True), (4, 2, False)):
        v = va.astype(oa.vdtype(isd))
        ho = oracle.build(rows, cols, rp, ci, v, cu, vf, isd, cdb)
        lay = spmvb.Layout.build(rows, cols, rp, ci, v, cu, vf, isd, cdb)
        assert lay.blocks == (cols + cdb - 1) // cdb
        assert oa.layouts_equal(oracle.snapshot(ho, rows, cu, vf, isd), product_snapshot(lay, cu, vf, isd)) == []
        oracle.free(ho); lay.free()


def test_builder_rejects_bad_arguments(spmvb):
    rows, cols, rp, ci, va = matgen.band(100)
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Layout.build(rows, cols, rp, ci, va, 1, 3, True)          # VF not in {1,2,4,8}
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Layout.build(rows, cols, rp, ci, va, 0, 1, True)          # CU < 1
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True, 65536)   # block wider than a 15-bit index
    bad = ci.copy(); bad[5] = cols + 7
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Layout.build(rows, cols, rp, bad, va, 1, 1, True)         # column out of range


def test_storage_overhead_matches_reference_formula(spmvb):
    """storage_overhead (csr_hw.cpp:1401-1409): 5 x 32 bit of metadata per block + (nr_ci + nr_val) bus words."""
    rows, cols, rp, ci, va = matgen.laplacian2d(150, 150)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 2, True)
    for k in range(2):
        bits = lay.blocks * 5 * 32 + sum((lay.piece_info(k, b)[3] + lay.piece_info(k, b)[4]) * 128 for b in range(lay.blocks))
        assert abs(lay.storage_mb(k) - bits / (8.0 * 1024 * 1024)) < 1e-12


def test_zero_row_list_covers_every_row_not_plainly_stored(spmvb):
    """Device aux data: rows outside the zero list must be single-block, non-empty rows."""
    rows, cols, rp, ci, va = matgen.laplacian2d(300, 300)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
    assert -1 <= lay.zero_rows <= rows
    rows2, cols2, rp2, ci2, va2 = matgen.uniform(3000, 200000, 16, seed=2)
    lay2 = spmvb.Layout.build(rows2, cols2, rp2, ci2, va2, 1, 1, True)
    assert lay2.zero_rows == -1  # every row spans several column blocks: the whole y is cleared


@pytest.mark.parametrize("case", ["lap", "rmat", "uniform16k", "ragged"])
def test_xs_plan_tiles_the_stream_and_windows_cover_the_columns(spmvb, case):
    """Work plan of the shared-memory-x kernel (host side of spmv_xs_kernel): items tile the chunk range exactly, each
    lies in one column block, is cut at block-relative multiples of run x 18 warps (fp64), and its x window (<= 128 KB,
    16-byte aligned) covers every column its chunks touch.  The plan is made for what the GPU streams: an irregular fp64
    matrix with 32 768-column API blocks gets an engine-private device layout with 16 384-column blocks (x slice of a
    block = 128 KB = the kernel's window, the split on index bit 14), so every window fits."""
    if case == "lap":
        rows, cols, rp, ci, va = matgen.laplacian2d(400, 300)
        cdb = 0
    elif case == "rmat":
        rows, cols, rp, ci, va = matgen.rmat(16, 4, seed=2)  # 65 536 columns: two 32 768-column blocks
        cdb = 0
    elif case == "uniform16k":
        rows, cols, rp, ci, va = matgen.uniform(20000, 60000, 16, seed=3)
        cdb = 16384
    else:
        rows, cols, rp, ci, va = matgen.ragged(30000, 150000, seed=4)
        cdb = 0
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True, cdb)
    n_cta, run_log2 = 12, 1
    items, first = lay.xs_plan(n_cta, run_log2)
    unit = (1 << run_log2) * 18
    dp = lay.device_params
    assert first[0] == 0 and first[-1] == len(items) and np.all(np.diff(first.astype(np.int64)) >= 0)
    pos = 0
    block_start = {}
    for c in range(lay.n_chunks):
        b = lay.chunk_cols(c)[2]
        block_start.setdefault(b, c)
    for it in items:
        begin, count, x_off, x_bytes, col_base, block = (int(v) for v in it[:6])
        assert begin == pos and count > 0
        pos += count
        assert (begin - block_start[block]) % unit == 0
        width = dp["cdb"]
        for c in range(begin, begin + count):
            lo, hi, b = lay.chunk_cols(c)
            assert b == block
            if x_bytes and lo <= hi:
                assert col_base <= lo and (hi - col_base + 1) * 8 <= x_bytes
        if x_bytes:
            assert x_bytes <= 128 * 1024 and x_bytes % 16 == 0 and (x_off * 8) % 16 == 0
            assert x_off == block * width + col_base
    assert pos == lay.n_chunks
    per_cta = [int(items[first[j]:first[j + 1], 1].sum()) for j in range(n_cta)]
    assert max(per_cta) - min(per_cta) <= 2 * unit + max(1, lay.n_chunks // n_cta // 4)
    assert np.all(items[:, 3] > 0)          # every window fits shared memory
    if case in ("rmat", "ragged"):          # fp64 x slice of a 32768-column block of an irregular matrix: 256 KB
        assert dp["private"] and dp["cdb"] == 16384 and lay.blocks == -(-cols // 32768)
    else:
        assert not dp["private"] and dp["cdb"] == (cdb or 32768)


def test_cu_major_device_order_keeps_pieces_bit_exact(spmvb, oracle):
    """The device order of the pieces (block-major or CU-major) is an engine decision: the pieces themselves do not
    change, and the work plan still tiles the stream."""
    rows, cols, rp, ci, va = matgen.uniform(6000, 100000, 12, seed=8)
    ho = oracle.build(rows, cols, rp, ci, va, 4, 2, True)
    ref = oracle.snapshot(ho, rows, 4, 2, True)
    for flag in (0, 1):
        with spmvb.options(cu_major=flag, dev_tiles=4, dev_cdb=32768):  # the device layout = the API layout
            lay = spmvb.Layout.build(rows, cols, rp, ci, va, 4, 2, True)
        assert oa.layouts_equal(ref, product_snapshot(lay, 4, 2, True)) == []
        items, first = lay.xs_plan(7, 1)
        assert int(items[:, 1].sum()) == lay.n_chunks
        blocks_seen = [lay.chunk_cols(int(c))[2] for c in items[:, 0]]
        assert all(int(b) == int(it[5]) for b, it in zip(blocks_seen, items))
        lay.free()
    oracle.free(ho)


def test_x_ranges_cover_exactly_the_touched_column_blocks(spmvb):
    """What set_x uploads: maximal runs of column blocks with at least one entry."""
    # whole matrix: one range over all blocks
    rows, cols, rp, ci, va = matgen.laplacian2d(300, 300)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 1, True)
    assert lay.x_ranges().tolist() == [[0, lay.blocks * 32768]]
    # a row shard of a banded matrix reads a band of x
    A = spmvb.Csr.laplacian2d(512, 2048, 300000, 500000)
    shard = spmvb.Layout.from_csr(A, 1, 1, 16384)
    r = shard.x_ranges()
    assert len(r) == 1
    lo, hi = (300000 - 512) // 16384 * 16384, -(-(500000 + 512) // 16384) * 16384
    assert r.tolist() == [[lo, hi]]
    # columns only in blocks 0, 2 and 3 of 5; every column index lies inside a range
    rng = np.random.default_rng(3)
    cdb, rows, cols = 4096, 200, 5 * 4096
    rp, ci = [0], []
    for _ in range(rows):
        c = np.concatenate([rng.integers(0, cdb, 2), rng.integers(2 * cdb, 4 * cdb, 3)])
        ci.extend(sorted(int(v) for v in set(c.tolist())))
        rp.append(len(ci))
    ci = np.array(ci, np.uint32)
    lay2 = spmvb.Layout.build(rows, cols, np.array(rp, np.uint64), ci, rng.random(len(ci)), 1, 1, True, cdb)
    r2 = lay2.x_ranges()
    assert r2.tolist() == [[0, cdb], [2 * cdb, 4 * cdb]]
    assert all(any(a <= c < b for a, b in r2.tolist()) for c in ci.tolist())
    # a matrix without entries reads nothing
    lay3 = spmvb.Layout.build(3, 10, np.zeros(4, np.uint64), np.zeros(0, np.uint32), np.zeros(0), 1, 1, True)
    assert lay3.x_ranges().shape == (0, 2)
