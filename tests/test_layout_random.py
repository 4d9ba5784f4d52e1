"""Randomised three-way layout parity (hypothesis): for arbitrary small CSR matrices - empty rows, duplicate and unsorted
columns, one-column matrices, rows that span several chunks - and arbitrary CU / VF / precision / block width, the
oracle (pinned to the reference), the host builder and the GPU builder's steps (emulated on the CPU) give the same
pieces, metadata and bitmap.  Seeds are derandomised so that the suite is reproducible."""
import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings, strategies as st

import oracle_api as oa
from test_layout_builder import product_snapshot
from test_layout_gpu_emu import emu  # noqa: F401  (fixture)


@st.composite
def csr_case(draw):
    rows = draw(st.integers(1, 40))
    cdb = draw(st.sampled_from([0, 0, 16384, 256, 64, 12, 4]))
    width = cdb or 32768
    cols = draw(st.one_of(st.integers(1, 300), st.integers(width - 3, 3 * width + 5)))
    cols = max(1, min(cols, 120000))
    seed = draw(st.integers(0, 2 ** 31 - 1))
    kind = draw(st.sampled_from(["short", "short", "long", "dups", "dense_row"]))
    empty = draw(st.floats(0.0, 0.7))
    rng = np.random.default_rng(seed)
    rp = [0]
    ci = []
    for r in range(rows):
        if r != rows - 1 and rng.random() < empty:
            n = 0
        elif kind == "long":
            n = int(rng.integers(1, 700))
        elif kind == "dense_row" and r == rows // 2:
            n = int(rng.integers(200, 1500))
        else:
            n = int(rng.integers(1, 9))
        if kind == "dups":
            c = rng.integers(0, cols, size=n)                      # duplicates kept as separate entries
        else:
            c = rng.choice(cols, size=min(n, cols), replace=False)
        if draw(st.booleans()):
            c = np.sort(c)
        ci.extend(int(v) for v in c)
        rp.append(len(ci))
    va = rng.uniform(-1, 1, len(ci))
    cu = draw(st.sampled_from([1, 1, 2, 3, 4, 8, 10, 12]))
    vf = draw(st.sampled_from([1, 2, 4, 8]))
    isd = draw(st.booleans())
    return rows, cols, np.array(rp, np.uint64), np.array(ci, np.uint32), va, cu, vf, isd, cdb


@settings(max_examples=300, deadline=None, derandomize=True,
          suppress_health_check=[HealthCheck.function_scoped_fixture, HealthCheck.too_slow, HealthCheck.data_too_large])
@given(case=csr_case())
def test_oracle_host_builder_and_gpu_builder_steps_agree(spmvb, oracle, emu, case):
    rows, cols, rp, ci, va, cu, vf, isd, cdb = case
    va = va.astype(oa.vdtype(isd))
    ho = oracle.build(rows, cols, rp, ci, va, cu, vf, isd, cdb)
    host = spmvb.Layout.build(rows, cols, rp, ci, va, cu, vf, isd, cdb)
    assert oa.layouts_equal(oracle.snapshot(ho, rows, cu, vf, isd), product_snapshot(host, cu, vf, isd)) == []
    dev = emu(rows, cols, rp, ci, va, cu, vf, isd, cdb)
    assert host.difference(dev) == ""
    # what set_x uploads: block-aligned, ascending, disjoint ranges that contain every column index of the matrix
    width = cdb or (16384 if cu in (10, 12) else 32768)
    r = host.x_ranges().astype(np.int64)
    assert np.array_equal(r, dev.x_ranges().astype(np.int64))
    assert np.all(r % width == 0) and np.all(r[:, 0] < r[:, 1]) and np.all(r[1:, 0] > r[:-1, 1])
    if len(ci):
        idx = np.searchsorted(r[:, 0], ci.astype(np.int64), side="right") - 1
        assert np.all(idx >= 0) and np.all(ci.astype(np.int64) < r[idx, 1])
        blocks_touched = np.unique(ci.astype(np.int64) // width)
        assert int((r[:, 1] - r[:, 0]).sum()) == len(blocks_touched) * width
    else:
        assert len(r) == 0
    oracle.free(ho); host.free(); dev.free()
