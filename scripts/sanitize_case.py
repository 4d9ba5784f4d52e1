"""Small SpMV cases through every kernel variant, meant to be run under compute-sanitizer (scripts/gpu_sanitize.sh):
memcheck for out-of-bounds accesses (the staged row-map slices, the x windows, the TMA ring), racecheck for shared-memory
hazards between the bulk copies and the lanes' loads.  Checks the results against the oracle as it goes."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import matgen  # noqa: E402
import oracle_api as oa  # noqa: E402
import spmvb  # noqa: E402

for kv in os.environ.get("SANITIZE_OPTS", "").split(","):  # e.g. SANITIZE_OPTS=xs_config=1
    if kv:
        spmvb.set_option(kv.split("=")[0], int(kv.split("=")[1]))
O = oa.OracleLib()
cases = [("ragged", matgen.ragged(1500, 70000, seed=7), 2, 2), ("uniform", matgen.uniform(1200, 70000, 16, seed=3), 1, 1),
         ("lap", matgen.laplacian2d(96, 96), 1, 1), ("rmat", matgen.rmat(11, 8, seed=5), 4, 1),
         ("longrow", matgen.uniform(12, 30000, 4000, seed=9), 1, 1)]
for isd in (True, False):
    vt = np.float64 if isd else np.float32
    for name, (rows, cols, rp, ci, va), cu, vf in cases:
        va = va.astype(vt)
        x = np.random.default_rng(1).random(cols).astype(vt)
        gold = O.spmv_gold(rows, rp, ci, va, x, isd).astype(np.float64)
        bound = O.abs_ax(rows, rp, ci, va, x, isd) * (1e-12 if isd else 1e-5) + np.finfo(vt).tiny
        for opts in ({}, {"dev_tiles": 3, "dev_cdb": 8192, "tall": 1}):
            with spmvb.options(**opts):
                lay = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
                for variant in (1, 7, 8):
                    eng = spmvb.Engine(lay, 0, variant)
                    y = np.zeros(rows, vt)
                    eng.spmv_host(x, y, accumulate=False)
                    eng.spmv_host(x, y, accumulate=False)
                    assert np.all(np.abs(y.astype(np.float64) - gold) <= bound), (name, isd, variant, opts)
                    eng.free()
                for rl in (23, 16, 13):  # the wide image and its kernel: one column block, a few, many
                    with spmvb.options(wide=1, wide_range_log2=rl, wide_hints=3 if rl == 16 else -1):
                        layw = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
                        eng = spmvb.Engine(layw, 0, spmvb.VARIANT_WIDE)
                        y = np.zeros(rows, vt)
                        eng.spmv_host(x, y, accumulate=False)
                        eng.spmv_host(x, y, accumulate=False)
                        assert np.all(np.abs(y.astype(np.float64) - gold) <= bound), (name, isd, "wide", rl, opts)
                        eng.free(); layw.free()
                if lay.ell_params["present"]:  # regular matrix: the sliced-ELLPACK image, one launch and the pipeline
                    for tiles in (1, 5):
                        with spmvb.options(ell_tiles=tiles):
                            eng = spmvb.Engine(lay, 0, spmvb.VARIANT_ELL)
                            y = np.zeros(rows, vt)
                            eng.spmv_host(x, y, accumulate=False)
                            eng.spmv_host(x, y, accumulate=False)
                            assert np.all(np.abs(y.astype(np.float64) - gold) <= bound), (name, isd, "ell", tiles, opts)
                            eng.free()
                lay2, eng2 = spmvb.Engine.from_csr(rows, cols, rp, ci, va, cu, vf, isd)  # the GPU layout builder
                y = np.zeros(rows, vt)
                eng2.spmv_host(x, y, accumulate=False)
                assert np.all(np.abs(y.astype(np.float64) - gold) <= bound), (name, isd, "gpu build", opts)
                eng2.free(); lay2.free(); lay.free()
        print("ok", name, "fp64" if isd else "fp32", flush=True)
# iterated callers: power iteration and CG
rows, cols, rp, ci, va = matgen.laplacian2d(64, 64)
lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
eng = spmvb.Engine(lay, 0)
xs, it, rel = eng.cg(np.ones(rows), max_iters=200, rel_tol=1e-8)
assert rel <= 1e-8
eng.set_x(np.full(cols, 1.0 / np.sqrt(cols)))
assert eng.power_iter(5) > 0
grp = spmvb.Group.create(rows, cols, rp, ci, va, True, devices=[0])
grp.set_x(np.full(cols, 1.0 / np.sqrt(cols)))
assert grp.power_iter(3) > 0
be = spmvb.bounds_errors()
if be is not None:  # bounds-checked build: every index of every launch above stayed inside its array
    print("bounds-checked kernels, violations caught:", be)
    assert not any(be.values()), be
print("sanitize_case: all results correct")
