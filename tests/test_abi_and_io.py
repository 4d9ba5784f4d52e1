"""CPU tests: the C-ABI library loads and exports every symbol include/spmvb.h declares; matrix-file reader/writer;
generators; row partition; engine calls fail loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import matgen
import oracle_api as oa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "spmvb.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(spmvb_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(spmvb):
    L = ctypes.CDLL(spmvb.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    missing = [n for n in names if not hasattr(L, n)]
    assert missing == []
    # and the python binding covers them all
    assert sorted(spmvb.SYMBOLS) == names
    assert spmvb.lib().spmvb_version() >= 100


def test_engine_fails_loudly_without_gpu(spmvb):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rows, cols, rp, ci, va = matgen.band(200)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va)
    with pytest.raises(spmvb.SpmvbError) as ei:
        spmvb.Engine(lay)
    assert ei.value.code == -3  # SPMVB_E_CUDA: there is no CPU fallback
    with pytest.raises(spmvb.SpmvbError) as ei:  # the GPU layout builder does not quietly build on the host either
        spmvb.Engine.from_csr(rows, cols, rp, ci, va)
    assert ei.value.code == -3


def test_host_calls_reject_null_handles(spmvb):
    """The end-to-end entry points return SPMVB_E_ARG (and say where) instead of touching a null engine / group / buffer."""
    import ctypes
    L = spmvb.lib()
    y = np.zeros(4)
    yp = y.ctypes.data_as(ctypes.c_void_p)
    for rc in (L.spmvb_engine_spmv_host(None, yp, 4, yp, 1), L.spmvb_engine_spmv_host_x_resident(None, yp, 1),
               L.spmvb_group_spmv_host(None, yp, 4, yp, 1), L.spmvb_group_spmv_host_rows(None, yp, 4, yp, 1)):
        assert rc == -1, rc  # SPMVB_E_ARG
        assert "spmv_host" in L.spmvb_last_error().decode()
    assert L.spmvb_group_x_over_links(None) == -1
    assert y.sum() == 0.0


def test_matrix_file_roundtrip_and_reference_reader(spmvb, tmp_path):
    rows, cols, rp, ci, va = matgen.ragged(300, 500, seed=4)
    path = str(tmp_path / "m.txt")
    matgen.write_matrix_file(path, rows, cols, rp, ci, va)
    A = spmvb.Csr.read(path, True)
    assert (A.rows, A.cols, A.nnz) == (rows, cols, len(ci))
    assert np.array_equal(A.row_ptr, rp) and np.array_equal(A.col_ind, ci) and np.array_equal(A.values, va)
    out = str(tmp_path / "m2.txt")
    A.write(out)
    B = spmvb.Csr.read(out, True)
    assert np.array_equal(B.row_ptr, rp) and np.array_equal(B.col_ind, ci) and np.array_equal(B.values, va)
    F = spmvb.Csr.read(out, False)
    assert np.array_equal(F.values, va.astype(np.float32))
    if oa.have_ref(1, 1, True):  # the reference's own reader parses what we write (csr.cpp:87-136)
        rc, got = oa.RefLib(1, 1, True).read_matrix_file(out)
        assert rc == 0
        r2, c2, n2, blocks, rp2, ci2, va2 = got
        assert (r2, c2, n2) == (rows, cols, len(ci)) and blocks == 1
        assert np.array_equal(rp2, rp) and np.array_equal(ci2, ci) and np.array_equal(va2, va)


def test_matrix_file_errors(spmvb, tmp_path):
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Csr.read(str(tmp_path / "missing.txt"))
    p = tmp_path / "bad.txt"
    p.write_text("3 3 2\n1 1 1.0\nnot a line\n")
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Csr.read(str(p))
    p.write_text("3 3 2\n2 1 1.0\n1 1 1.0\n")  # rows not sorted
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Csr.read(str(p))
    p.write_text("3 3 1\n1 1 1.0\n")  # trailing empty rows are fine (the reference leaves them undefined, Q3)
    A = spmvb.Csr.read(str(p))
    assert list(A.row_ptr) == [0, 1, 1, 1]


def test_reader_accepts_what_sscanf_lf_accepts(spmvb, tmp_path):
    """Value tokens are parsed in place (from_chars) with strtod behind it: everything the reference's "%lf" takes
    (csr.cpp:111) gives the same double - signs, exponents, hex floats, infinities, overflow and underflow."""
    toks = ["1", "-2.5", "+3.25", "1e-3", "-4.E+2", ".5", "5.", "0x1.8p1", "inf", "-INF", "1e400", "1e-400",
            "0.1", "-0.30000000000000004", "2.2250738585072011e-308", "1.7976931348623157e308", "123456789012345678901234567890"]
    p = tmp_path / "tok.txt"
    p.write_text("%d 1 %d\n" % (len(toks), len(toks)) + "".join("%d 1 %s\n" % (i + 1, t) for i, t in enumerate(toks)))
    A = spmvb.Csr.read(str(p), True)
    want = np.array([float.fromhex(t) if "x" in t else float(t) for t in toks])
    assert np.array_equal(A.values, want), (A.values, want)
    for bad in ("1.5x", "--1", "e5", "1 .5e", ""):
        p.write_text("1 1 1\n1 1 %s\n" % bad)
        if bad == "1 .5e":  # the value token is "1": what follows it on the line is ignored, like sscanf does
            B = spmvb.Csr.read(str(p), True)  # (views are valid while the object lives: keep it)
            assert B.values[0] == 1.0
            continue
        with pytest.raises(spmvb.SpmvbError):
            spmvb.Csr.read(str(p), True)


def test_generators_match_numpy_twins(spmvb):
    A = spmvb.Csr.laplacian2d(37, 23)
    rows, cols, rp, ci, va = matgen.laplacian2d(37, 23)
    assert np.array_equal(A.row_ptr, rp) and np.array_equal(A.col_ind, ci) and np.array_equal(A.values, va)
    S = spmvb.Csr.laplacian2d(37, 23, 100, 300)  # a row slice of the same operator (multi-GPU shard)
    lo, hi = int(rp[100]), int(rp[300])
    assert np.array_equal(S.col_ind, ci[lo:hi]) and np.array_equal(S.row_ptr, rp[100:301] - rp[100]) and S.cols == cols
    B = spmvb.Csr.band(1000, 5, 1)
    r2, c2, rp2, ci2, _ = matgen.band(1000, 5)
    assert np.array_equal(B.row_ptr, rp2) and np.array_equal(B.col_ind, ci2)
    assert spmvb.Csr.band(10000, 5, 1).nnz == 109970  # BASELINE config 1 (even: avoids the reference's Q1 defect)
    assert np.all(np.abs(B.values) <= 1.0) and np.all(B.values != 0.0)


def test_uniform_and_rmat_generators(spmvb):
    U = spmvb.Csr.uniform(5000, 100000, 16, seed=3)
    assert U.nnz == 5000 * 16
    c = U.col_ind.reshape(5000, 16).astype(np.int64)
    assert np.all(np.diff(c, axis=1) > 0) and c.max() < 100000          # distinct, sorted
    assert abs(c.mean() / 100000 - 0.5) < 0.01
    U2 = spmvb.Csr.uniform(5000, 100000, 16, seed=3, row_begin=1000, row_end=1500)
    assert np.array_equal(U2.col_ind, U.col_ind[1000 * 16:1500 * 16]) and np.array_equal(U2.values, U.values[16000:24000])
    R = spmvb.Csr.rmat(14, 16, seed=1)
    rp = R.row_ptr.astype(np.int64)
    assert R.rows == 1 << 14 and rp[-1] == R.nnz and R.nnz <= 16 << 14
    assert rp[-1] - rp[-2] >= 1                                          # last row non-empty (Q3)
    empty = float((np.diff(rp) == 0).mean())
    assert 0.2 < empty < 0.8                                             # the power-law tail leaves many empty rows
    keys = np.repeat(np.arange(R.rows, dtype=np.int64), np.diff(rp)) * R.cols + R.col_ind
    assert np.all(np.diff(keys) > 0)                                     # sorted by (row, col), no duplicates
    R2 = spmvb.Csr.rmat(14, 16, seed=1, row_begin=4096, row_end=8192)
    assert np.array_equal(R2.col_ind, R.col_ind[rp[4096]:rp[8192]])


def test_partition_rows_balances_nonzeros(spmvb):
    R = spmvb.Csr.rmat(15, 16, seed=2)
    for parts in (2, 4, 8):
        b = spmvb.partition_rows(R.rows, R.row_ptr, parts, 2)
        assert b[0] == 0 and b[-1] == R.rows and np.all(np.diff(b.astype(np.int64)) >= 0)
        rp = R.row_ptr.astype(np.int64)
        nnz = np.diff(rp[b])
        assert nnz.sum() == R.nnz
        assert nnz[:-1].min() > R.nnz // parts                            # S1: every fired part exceeds the mean
        assert np.all((b[1:-1] - b[:-2]) % 2 == 0)                        # S3: row counts multiple of RATIO_v


def test_binary_sidecar_roundtrip_and_cache(spmvb, tmp_path):
    """the parsed matrix is kept in binary next to the text file and reused while it is current."""
    import os
    import time
    A = spmvb.Csr.rmat(10, 8, seed=4)
    txt = str(tmp_path / "m.txt")
    A.write(txt)
    for isd in (True, False):
        B = spmvb.Csr.read_cached(txt, isd)                      # parses, writes the sidecar
        side = txt + (".f64.spmvb" if isd else ".f32.spmvb")
        assert os.path.exists(side)
        C = spmvb.Csr.read_cached(txt, isd)                      # loads the sidecar
        D = spmvb.Csr.load(side)
        for M in (B, C, D):
            assert (M.rows, M.cols, M.nnz, M.is_double) == (A.rows, A.cols, A.nnz, isd)
            assert np.array_equal(M.row_ptr, A.row_ptr) and np.array_equal(M.col_ind, A.col_ind)
            assert np.array_equal(M.values, A.values.astype(M.values.dtype))
    # a sidecar older than the text file is ignored and rewritten
    side = txt + ".f64.spmvb"
    A2 = spmvb.Csr.band(300, 2, seed=9)
    time.sleep(0.02)
    A2.write(txt)
    os.utime(side, (os.path.getmtime(txt) - 10, os.path.getmtime(txt) - 10))
    E = spmvb.Csr.read_cached(txt, True)
    assert E.rows == 300 and np.array_equal(E.col_ind, A2.col_ind)
    assert spmvb.Csr.load(side).rows == 300
    # corrupt / truncated / foreign files are errors, not garbage matrices
    raw = open(side, "rb").read()
    bad = str(tmp_path / "bad.spmvb")
    for blob in (raw[:-5], b"NOTSPMVB" + raw[8:], raw[:40] + b"\xff" * 8 + raw[48:]):
        open(bad, "wb").write(blob)
        with pytest.raises(spmvb.SpmvbError):
            spmvb.Csr.load(bad)
    with pytest.raises(spmvb.SpmvbError):
        spmvb.Csr.load(str(tmp_path / "missing.spmvb"))


def test_print_wide_of_the_drop_in_header(tmp_path):
    """print_wide (csr_hw.cpp:1493-1521): bus words as values / as 8 x (15-bit index <end-of-row>), the reference's text."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "pw.cpp"
    src.write_text('''#include "spmv_fpga_compat.h"
int main() {
  BusDataType w[2];
  ValueType v[RATIO_v];
  for (int k = 0; k < RATIO_v; k++) v[k] = (ValueType)(1.5 - k);
  memcpy(&w[0], v, 16);
  uint16_t ci[8] = {5, 0x8000 | 7, 0, 32767, 0x8000, 1, 2, 3};
  memcpy(&w[1], ci, 16);
  print_wide(&w[0], 1, 0);
  print_wide(&w[1], 1, 1);
  return 0;
}
''')
    lib = os.path.join(root, "spmv-fpga_b200", "lib")
    for isd, want0 in ((1, "| (63 : 0) = 1.5 | (127 : 64) = 0.5 |"),
                       (0, "| (31 : 0) = 1.5 | (63 : 32) = 0.5 | (95 : 64) = -0.5 | (127 : 96) = -1.5 |")):
        exe = str(tmp_path / ("pw%d" % isd))
        subprocess.check_call(["g++", "-O1", "-std=c++17", "-DCU=1", "-DVF=1", "-DDOUBLE=%d" % isd,
                               "-I" + os.path.join(root, "include"), "-o", exe, str(src), "-L" + lib, "-lspmvb",
                               "-Wl,-rpath," + lib])
        out = subprocess.run([exe], capture_output=True, text=True, check=True).stdout.splitlines()
        assert out[0] == want0
        assert out[1] == ("| (15 : 0) = 5 <0>\t| (31 : 16) = 7 <1>\t| (47 : 32) = 0 <0>\t| (63 : 48) = 32767 <0>\t"
                          "| (79 : 64) = 0 <1>\t| (95 : 80) = 1 <0>\t| (111 : 96) = 2 <0>\t| (127 : 112) = 3 <0>\t|")


def test_reference_main_on_the_engine_fails_loudly_without_gpu(spmvb, tmp_path):
    """oracle/_ref/refmain_on_b200_*.elf = the reference's unmodified main.cpp on the drop-in API: no GPU, no result."""
    import os
    import subprocess
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = os.path.join(root, "oracle", "_ref", "refmain_on_b200_cu1_vf1_d1.elf")
    if not os.path.exists(exe):
        pytest.skip("not built (needs /root/reference)")
    path = str(tmp_path / "b.txt")
    spmvb.Csr.band(500, 2, 1).write(path)
    p = subprocess.run([exe, path], capture_output=True, text=True)
    assert p.returncode != 0 and "no CUDA device" in p.stderr and "Verification" not in p.stdout


def test_options_are_explicit_and_scoped(spmvb):
    """Tuning goes through spmvb_set_option, never through the environment of a library call; the context manager
    restores what it changed; unknown names are errors."""
    assert spmvb.get_option("cu_major") == -1 and spmvb.get_option("tile_mb") == -1
    with spmvb.options(cu_major=1, tile_mb=8):
        assert spmvb.get_option("cu_major") == 1 and spmvb.get_option("tile_mb") == 8
    assert spmvb.get_option("cu_major") == -1 and spmvb.get_option("tile_mb") == -1
    with pytest.raises(spmvb.SpmvbError, match="unknown option"):
        spmvb.set_option("no_such_option", 1)
    # an environment variable alone changes nothing: two layouts of the same matrix are identical
    rows, cols, rp, ci, va = matgen.uniform(3000, 70000, 8, seed=2)
    a = spmvb.Layout.build(rows, cols, rp, ci, va, 4, 1, True)
    os.environ["SPMVB_CU_MAJOR"] = "1"
    try:
        b = spmvb.Layout.build(rows, cols, rp, ci, va, 4, 1, True)
    finally:
        del os.environ["SPMVB_CU_MAJOR"]
    assert a.difference(b) == ""
    # ... unless the executable asks for it (the C++ shim does, for drop-in programs)
    os.environ["SPMVB_TILE_MB"] = "8"
    try:
        assert spmvb.lib().spmvb_options_from_env() >= 1 and spmvb.get_option("tile_mb") == 8
    finally:
        del os.environ["SPMVB_TILE_MB"]
        spmvb.set_option("tile_mb", -1)


def test_device_layout_is_private_to_the_engine(spmvb, oracle):
    """The engine-private device layout (narrower column blocks, row tiles, VF 1) never shows through the API: piece
    tables, words and bitmap stay those of the caller's CU / VF / COLS_DIV_BLOCKS, bit-exact with the oracle."""
    rows, cols, rp, ci, va = matgen.uniform(5000, 100000, 16, seed=4)
    with spmvb.options(dev_tiles=3):
        lay = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 2, True)
    dp = lay.device_params
    assert dp["private"] and dp["cu"] == 3 and dp["vf"] == 1 and dp["cdb"] == 16384 and dp["cu_major"]
    assert lay.n_cu == 2 and lay.blocks == 4 and 150 < lay.x_lines_per_chunk <= 256
    ho = oracle.build(rows, cols, rp, ci, va, 2, 2, True)
    ref = oracle.snapshot(ho, rows, 2, 2, True)
    for b in range(lay.blocks):
        assert np.array_equal(lay.bitmap_row(b), ref.bitmap[b])
        for k in range(2):
            assert lay.piece_info(k, b) == ref.info[(k, b)]
    oracle.free(ho)
    # a banded matrix keeps its API layout as the device layout
    lap = spmvb.Layout.build(*matgen.laplacian2d(300, 300), 1, 1, True)
    assert not lap.device_params["private"] and lap.x_lines_per_chunk < 30


def test_builders_reject_a_bad_row_ptr_and_the_reader_a_short_line(spmvb, tmp_path):
    rows, cols, rp, ci, va = matgen.uniform(50, 80, 3, seed=1)
    bad = rp.copy().astype(np.uint64)
    bad[10] = bad[11] + 5            # decreasing
    with pytest.raises(spmvb.SpmvbError, match="row_ptr"):
        spmvb.Layout.build(rows, cols, bad, ci, va, 1, 1, True)
    bad = rp.copy().astype(np.uint64)
    bad[0] = 1                       # does not start at 0
    with pytest.raises(spmvb.SpmvbError, match="row_ptr"):
        spmvb.Layout.build(rows, cols, bad, ci, va, 1, 1, True)
    # a line with a missing field is a parse error, as with the reference's per-line sscanf (csr.cpp:111-113); the
    # value must not be taken from the next line, and a file that ends in such a line must not be read past its end
    p = tmp_path / "short.txt"
    p.write_text("3 3 3\n1 1 1.5\n2 2\n3 3 2.5\n")
    with pytest.raises(spmvb.SpmvbError, match="parse error"):
        spmvb.Csr.read(str(p))
    p.write_text("2 2 2\n1 1 1.5\n2 2")
    with pytest.raises(spmvb.SpmvbError, match="parse error"):
        spmvb.Csr.read(str(p))
