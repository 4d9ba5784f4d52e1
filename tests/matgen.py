"""Small seeded numpy matrix generators for the tests (independent of the product's C++ generators).

All return (rows, cols, row_ptr[uint64], col_ind[uint32], values[float64]) sorted by row, last row non-empty
(reference reader defect Q3), in the order the reference's file reader would produce.
"""
import numpy as np


def _finish(rows, cols, r, c, v, sort_cols=True):
    if sort_cols:
        order = np.lexsort((c, r))
    else:
        order = np.argsort(r, kind="stable")
    r, c, v = r[order], c[order], v[order]
    row_ptr = np.zeros(rows + 1, np.uint64)
    np.add.at(row_ptr, r + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.uint64)
    return rows, cols, row_ptr, c.astype(np.uint32), v.astype(np.float64)


def band(n, hb=5, seed=1):
    rng = np.random.default_rng(seed)
    r = np.repeat(np.arange(n), 2 * hb + 1)
    c = r + np.tile(np.arange(-hb, hb + 1), n)
    keep = (c >= 0) & (c < n)
    r, c = r[keep], c[keep]
    v = rng.uniform(-1, 1, len(r))
    return _finish(n, n, r, c, v)


def laplacian2d(nx, ny):
    n = nx * ny
    idx = np.arange(n)
    ix, iy = idx % nx, idx // nx
    rs, cs, vs = [idx], [idx], [np.full(n, 4.0)]
    for cond, off in ((iy > 0, -nx), (ix > 0, -1), (ix + 1 < nx, 1), (iy + 1 < ny, nx)):
        rs.append(idx[cond]); cs.append(idx[cond] + off); vs.append(np.full(cond.sum(), -1.0))
    return _finish(n, n, np.concatenate(rs), np.concatenate(cs), np.concatenate(vs))


def uniform(rows, cols, k, seed=1, empty_frac=0.0, sort_cols=True):
    """k random distinct columns per row; a fraction of rows left empty (never the last)."""
    rng = np.random.default_rng(seed)
    rs, cs = [], []
    for r in range(rows):
        if r != rows - 1 and rng.random() < empty_frac:
            continue
        kk = min(k, cols)
        c = rng.choice(cols, size=kk, replace=False)
        rs.append(np.full(kk, r)); cs.append(c)
    r = np.concatenate(rs); c = np.concatenate(cs)
    v = rng.uniform(-1, 1, len(r))
    return _finish(rows, cols, r, c, v, sort_cols=sort_cols)


def rmat(scale, ef=8, a=0.57, b=0.19, c=0.19, seed=1):
    rng = np.random.default_rng(seed)
    n = 1 << scale
    m = ef * n
    r = np.zeros(m, np.int64); cc = np.zeros(m, np.int64)
    for _ in range(scale):
        u = rng.random(m)
        rb = u >= a + b
        cb = ((u >= a) & (u < a + b)) | (u >= a + b + c)
        r = (r << 1) | rb; cc = (cc << 1) | cb
    key = np.unique(r * n + cc)
    r, cc = key // n, key % n
    if r[-1] != n - 1:
        r = np.append(r, n - 1); cc = np.append(cc, n - 1)
    v = rng.uniform(-1, 1, len(r))
    return _finish(n, n, r, cc, v)


def ragged(rows, cols, seed=1, max_len=40, empty_frac=0.3):
    """Power-law-ish row lengths with many empty rows and unsorted columns inside a row."""
    rng = np.random.default_rng(seed)
    rs, cs = [], []
    for r in range(rows):
        if r != rows - 1 and rng.random() < empty_frac:
            continue
        kk = int(min(cols, max(1, rng.pareto(1.2) * 3)))
        kk = min(kk, max_len)
        c = rng.choice(cols, size=kk, replace=False)
        rs.append(np.full(kk, r)); cs.append(c)
    r = np.concatenate(rs); c = np.concatenate(cs)
    v = rng.uniform(-1, 1, len(r))
    return _finish(rows, cols, r, c, v, sort_cols=False)


def write_matrix_file(path, rows, cols, row_ptr, col_ind, values, fmt="%.17g"):
    with open(path, "w") as f:
        f.write("%d %d %d\n" % (rows, cols, len(col_ind)))
        for r in range(rows):
            for j in range(int(row_ptr[r]), int(row_ptr[r + 1])):
                f.write(("%d %d " + fmt + "\n") % (r + 1, col_ind[j] + 1, values[j]))
