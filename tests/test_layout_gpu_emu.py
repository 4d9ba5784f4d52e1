"""The GPU layout builder's algorithm, checked without a GPU: tests/emu/layout_emu.cpp runs the very step bodies the
CUDA kernels run (spmv-fpga_b200/csrc/layout_gpu_steps.h) as serial loops in reverse index order and this file compares
the result - every table and byte - with the host builder, which the other tests pin to the reference.  The product
never runs this emulation; tests/test_gpu_layout_build.py checks the real kernels on the GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import matgen
import oracle_api as oa
from test_layout_builder import CFGS, MATS

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "layout_emu.cpp")
OUT = os.path.join(HERE, "_build", "liblayout_emu.so")
_vp, _u32, _int = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int


@pytest.fixture(scope="module")
def emu(spmvb):
    libdir = os.path.join(ROOT, "spmv-fpga_b200", "lib")
    deps = [SRC, os.path.join(ROOT, "spmv-fpga_b200", "csrc", "layout_gpu_steps.h"),
            os.path.join(ROOT, "spmv-fpga_b200", "csrc", "layout.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-o", OUT, SRC, "-L" + libdir, "-lspmvb",
                               "-Wl,-rpath," + libdir])
    L = ctypes.CDLL(OUT)
    L.emu_layout_build.restype = _int
    L.emu_layout_build.argtypes = [_u32, _u32, _vp, _vp, _vp, _int, _int, _int, _u32, _vp]

    def build(rows, cols, rp, ci, va, cu, vf, isd, cdb=0):
        rp = np.ascontiguousarray(rp, np.uint64); ci = np.ascontiguousarray(ci, np.uint32)
        va = np.ascontiguousarray(va, oa.vdtype(isd))
        out = _vp()
        rc = L.emu_layout_build(rows, cols, rp.ctypes.data_as(_vp), ci.ctypes.data_as(_vp), va.ctypes.data_as(_vp), cu, vf,
                                int(isd), cdb, ctypes.byref(out))
        if rc != 0:
            raise spmvb.SpmvbError(rc, spmvb.lib().spmvb_last_error().decode(errors="replace"))
        return spmvb.Layout(out.value, isd)
    return build


SORTED_MATS = sorted(MATS)  # includes "ragged" and "unsorted_cols": rows that visit their column blocks out of order


@pytest.mark.parametrize("cfg", CFGS, ids=lambda c: "cu%d_vf%d_%s" % (c[0], c[1], "f64" if c[2] else "f32"))
@pytest.mark.parametrize("mat", SORTED_MATS)
def test_gpu_builder_steps_match_host_builder(spmvb, emu, mat, cfg):
    cu, vf, isd = cfg
    rows, cols, rp, ci, va = MATS[mat]()
    host = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
    dev = emu(rows, cols, rp, ci, va, cu, vf, isd)
    assert host.difference(dev) == ""
    host.free(); dev.free()


@pytest.mark.parametrize("cdb", [16384, 4096, 256, 64, 12])
def test_gpu_builder_steps_custom_block_width_and_empty_rows(spmvb, emu, cdb):
    rows, cols, rp, ci, va = matgen.uniform(2000, 30000, 7, seed=31, empty_frac=0.4)
    for cu, vf, isd in ((1, 1, True), (4, 2, False), (2, 8, True)):
        host = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd, cdb)
        dev = emu(rows, cols, rp, ci, va, cu, vf, isd, cdb)
        assert host.difference(dev) == ""
        host.free(); dev.free()


def test_gpu_builder_steps_cu_major_and_run_granularity(spmvb, emu):
    rows, cols, rp, ci, va = matgen.uniform(3000, 100000, 12, seed=5)
    for opt in ({"cu_major": 1}, {"run_log2": 3}, {"zero_all": 1}, {"dev_tiles": 3, "dev_cdb": 8192}):
        with spmvb.options(**opt):
            host = spmvb.Layout.build(rows, cols, rp, ci, va, 4, 1, True, 16384)
            dev = emu(rows, cols, rp, ci, va, 4, 1, True, 16384)
            assert host.difference(dev) == ""
            if "dev_tiles" in opt:  # an engine-private device layout next to untouched API pieces
                assert host.device_params["private"] and host.device_params["cu"] == 3 and host.device_params["cdb"] == 8192
                assert host.n_cu == 4
            host.free(); dev.free()


def test_host_builder_large_allocation_path_against_the_gpu_steps(spmvb, emu):
    """Above 16 MB the host builder puts the stream and the row map into 2 MB-aligned, huge-page-advised memory touched
    by all cores (layout_big_alloc) - a path the small cases never take.  2 M non-zeros over 128 narrow blocks, CU = 1
    and CU = 4 (pass 2 + split), every table and byte against the independent builder."""
    rows, cols, rp, ci, va = matgen.uniform(1 << 17, 1 << 17, 16, seed=11)
    for cu, vf, isd in ((1, 1, True), (4, 2, True)):
        host = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd, 1024)
        assert host.stream_bytes > (16 << 20)
        dev = emu(rows, cols, rp, ci, va, cu, vf, isd, 1024)
        assert host.difference(dev) == ""
        host.free(); dev.free()


def test_gpu_builder_steps_degenerate_inputs(spmvb, emu):
    # one entry; one row; empty matrix (no non-zeros at all); duplicates inside a row
    cases = [
        (1, 1, [0, 1], [0], [2.5]),
        (1, 70000, [0, 3], [5, 40000, 69999], [1.0, 2.0, 3.0]),
        (4, 4, [0, 0, 0, 0, 0], [], []),
        (3, 50, [0, 4, 4, 6], [7, 7, 7, 9, 0, 49], [1, 2, 3, 4, 5, 6]),
    ]
    for rows, cols, rp, ci, va in cases:
        for cu, vf in ((1, 1), (2, 2), (8, 4)):
            args = (rows, cols, np.array(rp, np.uint64), np.array(ci, np.uint32), np.array(va, np.float64), cu, vf, True)
            host = spmvb.Layout.build(*args)
            dev = emu(*args)
            assert host.difference(dev) == ""
            host.free(); dev.free()


def test_gpu_builder_steps_reject_what_they_cannot_do(spmvb, emu):
    rows, cols, rp, ci, va = matgen.band(100)
    bad = ci.copy(); bad[17] = cols + 3
    with pytest.raises(spmvb.SpmvbError, match="out of range"):
        emu(rows, cols, rp, bad, va, 1, 1, True)
    with pytest.raises(spmvb.SpmvbError):
        emu(rows, cols, rp, ci, va, 1, 3, True)
    # columns unsorted INSIDE one column block are fine (order inside a (row, block) pair is kept as given)
    rows, cols, rp, ci, va = matgen.uniform(400, 3000, 9, seed=3, sort_cols=False)
    host = spmvb.Layout.build(rows, cols, rp, ci, va, 2, 2, True)
    dev = emu(rows, cols, rp, ci, va, 2, 2, True)
    assert host.difference(dev) == ""
