"""Diagnostic: host-built vs GPU-built engine on one small case, before and after other engines have used the device."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import matgen  # noqa: E402
import oracle_api as oa  # noqa: E402
import spmvb  # noqa: E402

O = oa.OracleLib()
rows, cols, rp, ci, va = matgen.ragged(1500, 70000, seed=7)
cu, vf, isd = 2, 2, True
x = np.random.default_rng(1).random(cols)
gold = O.spmv_gold(rows, rp, ci, va, x, isd)
bound = O.abs_ax(rows, rp, ci, va, x, isd) * 1e-12 + 1e-300


def gpu_build(tag):
    lay2, eng2 = spmvb.Engine.from_csr(rows, cols, rp, ci, va, cu, vf, isd)
    y = np.zeros(rows)
    eng2.spmv_host(x, y, accumulate=False)
    bad = np.nonzero(np.abs(y - gold) > bound)[0]
    eng2.fetch_layout()
    host = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
    print(tag, "variant", eng2.variant, "bad rows", len(bad), bad[:8], "max err/tol", float(np.max(np.abs(y - gold) / bound)),
          "layout difference: %r" % host.difference(lay2), eng2.device_layout, flush=True)
    if len(bad):
        print("   y", y[bad[:4]], "gold", gold[bad[:4]], flush=True)
    eng2.free(); lay2.free(); host.free()


with spmvb.options(dev_tiles=3, dev_cdb=8192, tall=1):
    gpu_build("fresh process       ")
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
    for variant in (1, 7, 8):
        eng = spmvb.Engine(lay, 0, variant)
        y = np.zeros(rows)
        eng.spmv_host(x, y, accumulate=False)
        assert np.all(np.abs(y - gold) <= bound), variant
        eng.free()
    gpu_build("after variants 1,7,8")
    for rl in (23, 16, 13):
        with spmvb.options(wide=1, wide_range_log2=rl):
            layw = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd)
            eng = spmvb.Engine(layw, 0, 9)
            y = np.zeros(rows)
            eng.spmv_host(x, y, accumulate=False)
            assert np.all(np.abs(y - gold) <= bound), ("wide", rl)
            eng.free(); layw.free()
        gpu_build("after wide 2^%d     " % rl)
