"""End-to-end call (spmv_hw semantics: pinned x up, kernel, y down, y_host += y) on the Laplacian of BASELINE configs[1]
for several engine choices: python scripts/exp_e2e.py [nx ny].  One line per configuration: ms per call, GFLOP/s."""
import ctypes
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle_api as oa  # noqa: E402
import spmvb  # noqa: E402

nx, ny = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2048, 2048)
A = spmvb.Csr.laplacian2d(nx, ny)
spmvb.lib()
rt = ctypes.CDLL("libcudart.so.12")


def pinned(a):
    rc = rt.cudaHostRegister(ctypes.c_void_p(a.ctypes.data), ctypes.c_size_t(a.nbytes), 0)
    assert rc == 0, rc
    return a


x = pinned(np.random.default_rng(1).random(A.cols))
y = pinned(np.zeros(A.rows))
O = oa.OracleLib()
gold, _ = O.spmv_gold_omp(A.rows, A.row_ptr, A.col_ind, A.values, x, True)
bound = O.abs_ax(A.rows, A.row_ptr, A.col_ind, A.values, x, True) * 1e-12 + 1e-300
lay = spmvb.Layout.from_csr(A)
configs = [("variant 7 (global-gather kernel, x up / kernel / y down one after the other)", 7, {}, True)]
for t in (1, 2, 4, 8, 16, 32, 64):
    configs.append(("variant 10 (ELL), %d row tiles" % t, 10, {"ell_tiles": t}, True))
configs.append(("variant 10 (ELL), 8 row tiles, accumulate = 0", 10, {"ell_tiles": 8}, False))
configs.append(("variant 10 (ELL), 16 row tiles, accumulate = 0", 10, {"ell_tiles": 16}, False))
configs.append(("variant 10 (ELL), 1 row tile, accumulate = 0", 10, {"ell_tiles": 1}, False))
for name, variant, opts, acc in configs:
    with spmvb.options(**opts):
        eng = spmvb.Engine(lay, 0, variant)
        y[:] = 0
        eng.spmv_host(x, y, accumulate=acc)
        err = float(np.max(np.abs(y - gold) / bound))
        for _ in range(3):
            eng.spmv_host(x, y, accumulate=acc)
        best = 1e9
        t0 = time.perf_counter()
        reps = 20
        for _ in range(reps):
            t1 = time.perf_counter()
            eng.spmv_host(x, y, accumulate=acc)
            best = min(best, time.perf_counter() - t1)
        ms = (time.perf_counter() - t0) / reps * 1e3
        print(json.dumps(dict(config=name, ms_per_call=ms, ms_best=best * 1e3, gflops=2.0 * A.nnz / ms / 1e6, err_over_tol=err)), flush=True)
        assert err <= 1.0
        eng.free()
