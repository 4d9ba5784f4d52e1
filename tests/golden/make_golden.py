"""Generates tests/golden/*.npz from the UNMODIFIED reference compiled under oracle/_ref (run in the build container,
where /root/reference exists: `make -C oracle ref && python tests/golden/make_golden.py`).

Each fixture holds a small seeded CSR matrix, x, and what the reference produced for one (CU, VF, DOUBLE) build:
piece metadata (nr_rows, nr_cols, nr_nzeros, nr_ci, nr_val), the submatrix words (unused index slots zeroed, SURVEY
Q5), the empty_rows_bitmap, the length of hw_x (checked here to be x zero-padded), y of spmv_hw (emulated) and y of
spmv_gold.  x itself is not stored: it is np.random.default_rng(x_seed).random(cols) cast to the value type.  Only inputs on which the reference is
well defined are used (even padded nnz for the last CU, every CU split fires, last row non-empty).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import matgen  # noqa: E402
import oracle_api as oa  # noqa: E402

CASES = [
    # name, matrix builder, (cu, vf, is_double)
    ("kat6x6_cu1_vf1_f64", lambda: (6, 6, np.array([0, 3, 4, 4, 6, 7, 9]), np.array([0, 2, 5, 1, 0, 3, 4, 0, 5]),
                                     np.arange(1, 10, dtype=float)), (1, 1, True)),
    ("band1k_cu1_vf1_f64", lambda: matgen.band(1000, 5, seed=1), (1, 1, True)),          # BASELINE config 1 shape
    ("band1k_cu2_vf2_f64", lambda: matgen.band(1000, 5, seed=1), (2, 2, True)),
    ("lap60x60_cu1_vf2_f64", lambda: matgen.laplacian2d(60, 60), (1, 2, True)),
    ("lap60x60_cu8_vf4_f64", lambda: matgen.laplacian2d(60, 60), (8, 4, True)),
    ("lap60x60_cu8_vf4_f32", lambda: matgen.laplacian2d(60, 60), (8, 4, False)),
    ("ragged_cu4_vf2_f64", lambda: matgen.ragged(1500, 70000, seed=3), (4, 2, True)),     # empty rows, 3 blocks
    ("ragged_cu1_vf4_f32", lambda: matgen.ragged(1500, 70000, seed=3), (1, 4, False)),
    ("uniform_cu12_vf8_f32", lambda: matgen.uniform(1600, 40000, 12, seed=6), (12, 8, False)),  # 16384-column blocks
    ("rmat11_cu2_vf4_f32", lambda: matgen.rmat(11, 8, seed=5), (2, 4, False)),
    ("uniform_cu10_vf2_f64", lambda: matgen.uniform(1600, 40000, 12, seed=6), (10, 2, True)),
    ("uniform_cu8_vf8_f64", lambda: matgen.uniform(1200, 100000, 8, seed=2), (8, 8, True)),
]


def main():
    for name, build, (cu, vf, isd) in CASES:
        rows, cols, rp, ci, va = build()
        vt = oa.vdtype(isd)
        va = va.astype(vt)
        x = np.random.default_rng(42).random(cols).astype(vt)
        R = oa.RefLib(cu, vf, isd)
        h = R.build(rows, cols, rp, ci, va)
        snap = R.snapshot(h, rows)
        hwx = np.concatenate(R.hw_x(h, x))
        assert np.array_equal(hwx[:cols], x) and not hwx[cols:].any()  # hw_x = x zero-padded (csr_hw.cpp:1470-1488)
        y = np.zeros(rows, vt)
        rc = R.spmv_hw(h, x, y)
        assert rc == 0, name
        gold = R.spmv_gold(rows, cols, rp, ci, va, x)
        assert R.L.ref_verification(rows, oa._ptr(gold), oa._ptr(y)) == 0, name
        out = dict(rows=rows, cols=cols, row_ptr=np.asarray(rp, np.uint32), col_ind=np.asarray(ci, np.uint32), values=va,
                   x_seed=42, cu=cu, vf=vf, is_double=int(isd), blocks=snap.blocks, hw_x_len=len(hwx), y_hw=y, y_gold=gold,
                   bitmap=np.stack(snap.bitmap))
        info = np.zeros((cu, snap.blocks, 5), np.uint32)
        for k in range(cu):
            for b in range(snap.blocks):
                info[k, b] = snap.info[(k, b)]
                out["words_%d_%d" % (k, b)] = snap.masked_words(k, b)
        out["info"] = info
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        R.free(h)
        print(name, "blocks", snap.blocks, "nnz", len(ci))


if __name__ == "__main__":
    main()
