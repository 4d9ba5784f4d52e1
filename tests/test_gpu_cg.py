"""GPU tests of the conjugate-gradient caller (spmvb_engine_cg) on the 2-D Laplacian: the solution
is checked through the oracle's CSR SpMV (true residual) and against a float64 CG written with the same oracle SpMV."""
import time

import numpy as np
import pytest

import matgen
import oracle_api as oa

pytestmark = pytest.mark.gpu


def oracle_cg(oracle, M, b, iters_max, tol):
    rows, cols, rp, ci, va = M
    x = np.zeros(rows); r = b.astype(np.float64).copy(); p = r.copy()
    rr = r @ r; b2 = rr
    it = 0
    while it < iters_max and rr > tol * tol * b2:
        q = oracle.spmv_gold(rows, rp, ci, va, p, True)
        alpha = rr / (p @ q)
        x += alpha * p; r -= alpha * q
        rr_new = r @ r
        p = r + (rr_new / rr) * p
        rr = rr_new; it += 1
    return x, it


@pytest.mark.parametrize("isd,tol,slack", [(True, 1e-10, 1e-9), (False, 1e-5, 2e-4)], ids=["f64", "f32"])
def test_cg_solves_the_laplacian(spmvb, oracle, isd, tol, slack):
    M = matgen.laplacian2d(200, 150)
    rows, cols, rp, ci, va = M
    vt = oa.vdtype(isd)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va.astype(vt), 1, 1, isd)
    eng = spmvb.Engine(lay, 0)
    b = np.random.default_rng(7).standard_normal(rows).astype(vt)
    x, it, rel = eng.cg(b, max_iters=5000, rel_tol=tol)
    assert 0 < it < 5000 and rel <= tol
    # true residual through the oracle's gold SpMV, in float64
    ax = oracle.spmv_gold(rows, rp, ci, va, x.astype(np.float64), True)
    true_rel = np.linalg.norm(b.astype(np.float64) - ax) / np.linalg.norm(b.astype(np.float64))
    assert true_rel <= slack, true_rel
    x_ref, it_ref = oracle_cg(oracle, M, b, 5000, tol)
    assert abs(it - it_ref) <= 8 + it_ref // 10          # the GPU looks at the residual every 8 iterations
    assert np.linalg.norm(x.astype(np.float64) - x_ref) <= (1e-5 if isd else 2e-2) * np.linalg.norm(x_ref)
    # b = 0 -> x = 0 without iterating; the engine is still usable for plain SpMV afterwards
    x0, it0, rel0 = eng.cg(np.zeros(rows, vt))
    assert it0 == 0 and not x0.any()
    xx = np.random.default_rng(1).random(cols).astype(vt)
    y = np.zeros(rows, vt)
    eng.spmv_host(xx, y, accumulate=True)
    gold = oracle.spmv_gold(rows, rp, ci, va.astype(vt), xx, isd)
    scale = oracle.abs_ax(rows, rp, ci, va.astype(vt), xx, isd)
    assert np.all(np.abs(y.astype(np.float64) - gold.astype(np.float64)) <= (1e-12 if isd else 1e-5) * scale + 1e-300)


def test_cg_rejects_bad_input(spmvb):
    rows, cols, rp, ci, va = matgen.uniform(50, 80, 3, seed=1)
    lay = spmvb.Layout.build(rows, cols, rp, ci, va, 1, 1, True)
    eng = spmvb.Engine(lay, 0)
    with pytest.raises(spmvb.SpmvbError, match="square"):
        eng.cg(np.ones(rows))


@pytest.mark.perf
def test_cg_at_config2_scale_runs_and_reports_its_iteration_time(spmvb):
    """Config-2 size: the iteration count is honoured and the residual stays finite.  Times are PRINTED, never
    asserted: a correctness run may sit under a tracer or share the box, and wall-clock says nothing about parity.
    (Round 1 asserted < 1 ms per iteration here; the driver's box measured 2.46 ms against 0.145 ms on the builder's,
    `-x` then hid every parity test.)  The device time comes from CUDA events around the iteration loop."""
    A = spmvb.Csr.laplacian2d(2048, 2048)
    lay = spmvb.Layout.from_csr(A)
    eng = spmvb.Engine(lay, 0)
    b = np.ones(A.rows)
    for iters in (8, 208):
        t0 = time.perf_counter()
        x, it, rel = eng.cg(b, max_iters=iters, rel_tol=0.0)
        wall = time.perf_counter() - t0
        assert it == iters and np.isfinite(rel) and np.all(np.isfinite(x))
        print("\nCG on the 2048 x 2048 Laplacian, %d iterations: device %.1f us / iteration (CUDA events), whole call %.1f ms "
              "wall clock (includes the 32 MB upload of b and download of x), residual %.3g"
              % (iters, eng.last_iter_ms * 1e3, wall * 1e3, rel))
