"""Randomised differential test of the oracle (C restatement) against the compiled, unmodified reference
(oracle/_ref, built where /root/reference exists): arbitrary small matrices with empty rows, unsorted and duplicate
columns, for every (CU, VF, DOUBLE) the reference was compiled for.  The reference is only called on inputs where it is
well defined: every CU split fires in every block (else it frees garbage) and every piece holds a whole
number of value words (else hw_matrix_alloc under-allocates, Q1) - the oracle's own result says which inputs those are."""
import numpy as np
import pytest
from hypothesis import HealthCheck, assume, given, settings, strategies as st

import oracle_api as oa

CONFIGS = [(1, 1, False), (1, 1, True), (1, 2, True), (1, 4, False), (1, 4, True), (2, 1, True), (2, 2, True),
           (2, 4, False), (4, 2, True), (4, 4, False), (8, 4, False), (8, 4, True), (8, 8, True), (10, 2, True),
           (12, 4, True), (12, 8, False)]
AVAILABLE = [c for c in CONFIGS if oa.have_ref(*c)]
_refs = {}


def ref_lib(cfg):
    if cfg not in _refs:
        _refs[cfg] = oa.RefLib(*cfg)
    return _refs[cfg]


@st.composite
def case(draw):
    cfg = draw(st.sampled_from(AVAILABLE))
    cu, vf, isd = cfg
    rows = cu * draw(st.integers(6, 40)) + draw(st.integers(0, 7))
    cols = draw(st.sampled_from([7, 300, 300, 16384, 32768, 32769, 50000, 70000]))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    empty = draw(st.floats(0.0, 0.5))
    maxlen = draw(st.sampled_from([4, 12, 300]))
    sort_cols = draw(st.booleans())
    dups = draw(st.booleans())
    rng = np.random.default_rng(seed)
    rp, ci = [0], []
    for r in range(rows):
        n = 0 if (r != rows - 1 and rng.random() < empty) else int(rng.integers(1, maxlen + 1))
        c = rng.integers(0, cols, size=n) if dups else rng.choice(cols, size=min(n, cols), replace=False)
        if sort_cols:
            c = np.sort(c)
        ci.extend(int(v) for v in c)
        rp.append(len(ci))
    va = rng.uniform(-1, 1, len(ci))
    return cfg, rows, cols, np.array(rp, np.uint64), np.array(ci, np.uint32), va, seed


@pytest.mark.skipif(not AVAILABLE, reason="oracle/_ref not built (needs /root/reference)")
@settings(max_examples=250, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.filter_too_much,
                                 HealthCheck.data_too_large])
@given(c=case())
def test_oracle_equals_compiled_reference_on_random_matrices(oracle, c):
    (cu, vf, isd), rows, cols, rp, ci, va, seed = c
    vt = oa.vdtype(isd)
    va = va.astype(vt)
    ratio_v = 2 if isd else 4
    ho = oracle.build(rows, cols, rp, ci, va, cu, vf, isd)
    so = oracle.snapshot(ho, rows, cu, vf, isd)
    safe = all(so.info[(k, b)][0] > 0 and so.info[(k, b)][2] % ratio_v == 0 for k in range(cu) for b in range(so.blocks))
    if not safe:
        oracle.free(ho)
        assume(False)
    R = ref_lib((cu, vf, isd))
    hr = R.build(rows, cols, rp, ci, va)
    assert oa.layouts_equal(R.snapshot(hr, rows), so) == []
    x = np.random.default_rng(seed ^ 5).random(cols).astype(vt)
    y_ref = np.zeros(rows, vt); y_orc = np.zeros(rows, vt)
    assert R.spmv_hw(hr, x, y_ref) == 0
    assert oracle.spmv_emu(ho, x, y_orc, isd) == 0
    assert np.array_equal(y_ref.view(np.uint8), y_orc.view(np.uint8))      # bit-identical y
    gold_ref = R.spmv_gold(rows, cols, rp, ci, va, x)
    gold_orc = oracle.spmv_gold(rows, rp, ci, va, x, isd)
    assert np.array_equal(gold_ref.view(np.uint8), gold_orc.view(np.uint8))
    R.free(hr); oracle.free(ho)
