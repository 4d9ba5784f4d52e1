"""GPU tests of the drop-in C++ host API (include/spmv_fpga_compat.h): the run.elf analogue built with the reference's
compile-time macros, run on matrix files in the reference's format, next to the reference's own executable."""
import os
import re
import subprocess

import pytest

import matgen

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OURS = os.path.join(ROOT, "spmv-fpga_b200", "lib", "run_cu%d_vf%d_d%d.elf")
REF = os.path.join(ROOT, "oracle", "_ref", "ref_run_cu%d_vf%d_d%d.elf")
REFMAIN = os.path.join(ROOT, "oracle", "_ref", "refmain_on_b200_cu%d_vf%d_d%d.elf")


def run(exe, path, env=None):
    p = subprocess.run([exe, path], capture_output=True, text=True, timeout=300, env=dict(os.environ, **(env or {})))
    return p.returncode, p.stdout + p.stderr


def last_line(out):
    return [l for l in out.splitlines() if l.startswith("CSR representation")][-1]


@pytest.mark.parametrize("cfg", [(1, 1, 1), (8, 4, 1), (8, 4, 0)], ids=lambda c: "cu%d_vf%d_d%d" % c)
def test_run_elf_verifies_and_reports_like_the_reference(spmvb, tmp_path, cfg):
    exe = OURS % cfg
    if not os.path.exists(exe):
        pytest.fail("driver %s not built (run __graft_entry__.build())" % exe)
    if cfg == (1, 1, 1):
        A = spmvb.Csr.band(10000, 5, 1)  # BASELINE config 0: 10k-row band, 109 970 nnz
        path = str(tmp_path / "band10k.txt")
        A.write(path)
    else:
        rows, cols, rp, ci, va = matgen.laplacian2d(250, 250)  # 2 column blocks; every CU split fires
        path = str(tmp_path / "lap.txt")
        matgen.write_matrix_file(path, rows, cols, rp, ci, va, fmt="%.17g" if cfg[2] else "%.9g")
    rc, out = run(exe, path)
    assert rc == 0, out
    assert "Verification PASSED!" in out
    # the same run with create_csr_hw_matrix building the layout on the GPU: identical report
    grc, gout = run(exe, path, {"SPMVB_GPU_BUILD": "1"})
    assert grc == 0, gout
    assert "Verification PASSED!" in gout
    assert last_line(gout) == last_line(out)
    assert re.search(r"Welcome to SpMV \(Compute Units : %d, Vectorization Factor : %d" % cfg[:2], out)
    ref = REF % cfg
    if os.path.exists(ref):  # prebuilt from the unmodified reference (oracle/Makefile ref_elf); travels with the repo
        rrc, rout = run(ref, path)
        assert rrc == 0 and "Verification PASSED!" in rout
        assert last_line(out) == last_line(rout)  # same storage-overhead report, digit for digit
        assert [l for l in out.splitlines() if l.startswith("Total non-zeros")][0].split(".")[0] == \
               [l for l in rout.splitlines() if l.startswith("Total non-zeros")][0].split(".")[0]


def test_missing_file_is_reported_like_the_reference(tmp_path):
    exe = OURS % (1, 1, 1)
    rc, out = run(exe, str(tmp_path / "nope.txt"))
    assert rc == 1 and "Could not open file" in out and "Error reading matrix header" in out


@pytest.mark.parametrize("cfg", [(1, 1, 1), (8, 4, 1), (8, 4, 0)], ids=lambda c: "cu%d_vf%d_d%d" % c)
def test_unmodified_reference_main_runs_on_the_b200_engine(spmvb, tmp_path, cfg):
    """The reference's own src/main.cpp, unmodified, compiled against include/refnames/ + libspmvb.so (oracle/Makefile
    dropin_elf, built where /root/reference exists): it must verify its result and print the reference's report."""
    exe = REFMAIN % cfg
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/refmain_on_b200_*.elf not built (needs /root/reference at build time)")
    if cfg == (1, 1, 1):
        # BASELINE config 0; an even non-zero count per block: the reference's own run.elf under-runs its value FIFO
        # otherwise and could not be run next to it
        path = str(tmp_path / "band10k.txt")
        spmvb.Csr.band(10000, 5, 1).write(path)
    else:
        rows, cols, rp, ci, va = matgen.laplacian2d(250, 250)  # 2 column blocks; every CU split fires
        path = str(tmp_path / "lap.txt")
        matgen.write_matrix_file(path, rows, cols, rp, ci, va, fmt="%.17g" if cfg[2] else "%.9g")
    rc, out = run(exe, path)
    assert rc == 0, out
    assert "Verification PASSED!" in out
    assert re.search(r"Welcome to SpMV \(Compute Units : %d, Vectorization Factor : %d" % cfg[:2], out)
    ref = REF % cfg
    if os.path.exists(ref):
        rrc, rout = run(ref, path)
        if rrc != 0:  # the reference's own emulation build aborts on some inputs (FIFO under-run)
            pytest.skip("the reference's run.elf does not survive this input")
        assert "Verification PASSED!" in rout
        assert last_line(out) == last_line(rout)
