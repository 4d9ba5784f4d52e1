"""CPU tests of the oracle itself: the C restatement (oracle/spmv_oracle.c) against the golden fixtures generated from
the unmodified reference, and - where oracle/_ref exists - directly against the compiled reference."""
import glob
import os

import numpy as np
import pytest

import matgen
import oracle_api as oa

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def load_fixture(path):
    z = np.load(path)
    isd = bool(z["is_double"])
    vt = oa.vdtype(isd)
    fx = dict(rows=int(z["rows"]), cols=int(z["cols"]), rp=z["row_ptr"], ci=z["col_ind"], va=z["values"],
              cu=int(z["cu"]), vf=int(z["vf"]), isd=isd, blocks=int(z["blocks"]), info=z["info"], bitmap=z["bitmap"],
              y_hw=z["y_hw"], y_gold=z["y_gold"], hw_x_len=int(z["hw_x_len"]))
    fx["x"] = np.random.default_rng(int(z["x_seed"])).random(fx["cols"]).astype(vt)
    lay = oa.Layout(fx["cu"], fx["vf"], isd, fx["blocks"])
    for b in range(fx["blocks"]):
        lay.bitmap.append(z["bitmap"][b])
        for k in range(fx["cu"]):
            lay.info[(k, b)] = tuple(int(v) for v in z["info"][k, b])
            lay.words[(k, b)] = z["words_%d_%d" % (k, b)]
    fx["layout"] = lay
    return fx


def test_golden_fixtures_exist():
    assert len(GOLDEN) >= 10


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_layout_matches_reference_golden(oracle, path):
    fx = load_fixture(path)
    h = oracle.build(fx["rows"], fx["cols"], fx["rp"], fx["ci"], fx["va"], fx["cu"], fx["vf"], fx["isd"])
    snap = oracle.snapshot(h, fx["rows"], fx["cu"], fx["vf"], fx["isd"])
    assert oa.layouts_equal(fx["layout"], snap) == []
    assert oracle.expanded_cols(h) == fx["hw_x_len"]
    hwx = oracle.hw_x(h, fx["x"], fx["isd"])
    assert np.array_equal(hwx[: fx["cols"]], fx["x"]) and not hwx[fx["cols"]:].any()
    oracle.free(h)


@pytest.mark.parametrize("path", GOLDEN, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_spmv_matches_reference_golden_bit_exact(oracle, path):
    """The emulated kernel arithmetic (compute_results chunking, block order of the accumulation) and the gold loop
    are restated operation by operation, so y must be bit-identical to what the reference produced."""
    fx = load_fixture(path)
    vt = oa.vdtype(fx["isd"])
    h = oracle.build(fx["rows"], fx["cols"], fx["rp"], fx["ci"], fx["va"], fx["cu"], fx["vf"], fx["isd"])
    y = np.zeros(fx["rows"], vt)
    assert oracle.spmv_emu(h, fx["x"], y, fx["isd"]) == 0
    assert np.array_equal(y.view(np.uint8), fx["y_hw"].view(np.uint8))
    gold = oracle.spmv_gold(fx["rows"], fx["rp"], fx["ci"], fx["va"], fx["x"], fx["isd"])
    assert np.array_equal(gold.view(np.uint8), fx["y_gold"].view(np.uint8))
    # spmv_hw accumulates (csr_hw.cpp:1557): a second call on the same y doubles it
    assert oracle.spmv_emu(h, fx["x"], y, fx["isd"]) == 0
    scale = oracle.abs_ax(fx["rows"], fx["rp"], fx["ci"], fx["va"], fx["x"], fx["isd"])
    tol = 1e-12 if fx["isd"] else 1e-5
    assert np.all(np.abs(y.astype(np.float64) - 2.0 * fx["y_hw"].astype(np.float64)) <= 4 * tol * scale + 1e-300)
    assert oracle.L.orc_verification(fx["rows"], oa._ptr(gold), oa._ptr(fx["y_hw"].copy()), int(fx["isd"])) == 0
    oracle.free(h)


def test_kat_hand_example_words(oracle):
    """the 6x6 example, word for word."""
    rp = np.array([0, 3, 4, 4, 6, 7, 9]); ci = np.array([0, 2, 5, 1, 0, 3, 4, 0, 5]); va = np.arange(1, 10, dtype=float)
    h = oracle.build(6, 6, rp, ci, va, 1, 1, True)
    s = oracle.snapshot(h, 6, 1, 1, True)
    assert s.info[(0, 0)] == (6, 6, 10, 2, 5)
    assert list(s.bitmap[0]) == [0, 0, 1, 0, 0, 0]
    w = s.masked_words(0, 0).view(np.uint64)
    expect = [0x8001800500020000, 0x0000800480030000, 0x3ff0000000000000, 0x4000000000000000, 0x4008000000000000,
              0x4010000000000000, 0x4014000000000000, 0x4018000000000000, 0x401c000000000000, 0x4020000000000000,
              0x0000000080008005, 0, 0x4022000000000000, 0]
    assert [int(v) for v in w] == expect
    y = np.zeros(6)
    assert oracle.spmv_emu(h, np.arange(1.0, 7.0), y, True) == 0
    assert list(y) == [25.0, 8.0, 0.0, 29.0, 35.0, 62.0]
    oracle.free(h)


REF_CASES = [
    ("band", lambda: matgen.band(3000, 5, seed=2), (1, 1, True)),
    ("lap2blk", lambda: matgen.laplacian2d(200, 200), (2, 2, True)),
    ("lap2blk", lambda: matgen.laplacian2d(200, 200), (8, 4, False)),
    ("ragged", lambda: matgen.ragged(4000, 90000, seed=11), (4, 2, True)),
    ("ragged", lambda: matgen.ragged(4000, 90000, seed=11), (1, 4, False)),
    ("uniform", lambda: matgen.uniform(3000, 50000, 16, seed=12), (12, 4, True)),
    ("rmat", lambda: matgen.rmat(12, 8, seed=13), (2, 4, False)),
]


@pytest.mark.parametrize("case", REF_CASES, ids=lambda c: "%s_cu%d_vf%d_%s" % (c[0], c[2][0], c[2][1], "f64" if c[2][2] else "f32"))
def test_oracle_matches_compiled_reference(oracle, case):
    """Differential test against oracle/_ref (only in the build container, where /root/reference was compiled)."""
    name, build, (cu, vf, isd) = case
    if not oa.have_ref(cu, vf, isd):
        pytest.skip("oracle/_ref not built (needs /root/reference)")
    rows, cols, rp, ci, va = build()
    vt = oa.vdtype(isd)
    va = va.astype(vt)
    x = np.random.default_rng(3).random(cols).astype(vt)
    R = oa.RefLib(cu, vf, isd)
    hr = R.build(rows, cols, rp, ci, va)
    ho = oracle.build(rows, cols, rp, ci, va, cu, vf, isd)
    assert oa.layouts_equal(R.snapshot(hr, rows), oracle.snapshot(ho, rows, cu, vf, isd)) == []
    y_ref = np.zeros(rows, vt); y_orc = np.zeros(rows, vt)
    assert R.spmv_hw(hr, x, y_ref) == 0
    assert oracle.spmv_emu(ho, x, y_orc, isd) == 0
    assert np.array_equal(y_ref.view(np.uint8), y_orc.view(np.uint8))
    R.free(hr); oracle.free(ho)


def test_openmp_gold_is_bit_identical_to_the_single_thread_loop(oracle):
    rows, cols, rp, ci, va = matgen.ragged(5000, 60000, seed=17)
    x = np.random.default_rng(1).random(cols)
    y1 = oracle.spmv_gold(rows, rp, ci, va, x, True)
    y2, threads = oracle.spmv_gold_omp(rows, rp, ci, va, x, True)
    assert threads >= 1 and np.array_equal(y1.view(np.uint8), y2.view(np.uint8))
