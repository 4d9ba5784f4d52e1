"""ctypes access to the test oracles (TEST INFRASTRUCTURE ONLY).

Two oracles, both under oracle/ (see oracle/Makefile):
  * ``RefLib``     - the UNMODIFIED reference compiled from /root/reference/src into
                     oracle/_ref/libref_cu<C>_vf<V>_d<D>.so (one per compile-time config).
  * ``OracleLib``  - the plain-C restatement oracle/spmv_oracle.c -> oracle/_ref/liboracle.so.
Nothing in the product path (spmv-fpga_b200/) imports this module.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_DIR = os.path.join(ORACLE_DIR, "_ref")

_vp = ctypes.c_void_p


def _ptr(a):
    return a.ctypes.data_as(_vp)


def vdtype(is_double):
    return np.float64 if is_double else np.float32


def build_port():
    """Compile the C restatement if it is missing or stale (gcc only; no reference needed)."""
    so = os.path.join(REF_DIR, "liboracle.so")
    src = os.path.join(ORACLE_DIR, "spmv_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
    return so


def ref_so_path(cu, vf, is_double):
    return os.path.join(REF_DIR, "libref_cu%d_vf%d_d%d.so" % (cu, vf, 1 if is_double else 0))


def have_ref(cu, vf, is_double):
    return os.path.exists(ref_so_path(cu, vf, is_double))


class Layout:
    """Plain-python snapshot of an hw_matrix layout (from either oracle or the product)."""

    def __init__(self, cu, vf, is_double, blocks):
        self.cu, self.vf, self.is_double, self.blocks = cu, vf, is_double, blocks
        self.info = {}    # (cu, block) -> (nr_rows, nr_cols, nr_nzeros, nr_ci, nr_val)
        self.words = {}   # (cu, block) -> np.uint8 array of the piece bytes
        self.bitmap = []  # per block np.uint8[rows]

    def masked_words(self, k, b):
        """Piece bytes with the never-written index slots (>= nnz) of the last group zeroed
        and trimmed to nr_ci + ceil(nnz/RATIO_v) words."""
        nr_rows, nr_cols, nnz, nr_ci, nr_val = self.info[(k, b)]
        ratio_v = 2 if self.is_double else 4
        rcv = 8 // ratio_v + 1
        nwords = nr_ci + (nnz + ratio_v - 1) // ratio_v
        w = np.array(self.words[(k, b)][: nwords * 16], dtype=np.uint8, copy=True)
        rem = nnz % 8
        if rem:
            g = nnz // 8
            base = g * rcv * 16
            w[base + 2 * rem: base + 16] = 0
            # value lanes of the last value word beyond nnz are never written either
            vb = 8 if self.is_double else 4
            end_vals = base + 16 + rem * vb
            w[end_vals:] = 0
        return w


def _snapshot(lib, h, prefix, cu, vf, is_double, rows, words_fn=None):
    blocks = getattr(lib, prefix + "blocks")(h)
    lay = Layout(cu, vf, is_double, blocks)
    ratio_v = 2 if is_double else 4
    info = (ctypes.c_uint32 * 5)()
    for b in range(blocks):
        bm = getattr(lib, prefix + "bitmap_row")(h, b)
        lay.bitmap.append(np.ctypeslib.as_array(ctypes.cast(bm, ctypes.POINTER(ctypes.c_uint8)), shape=(rows,)).copy())
        for k in range(cu):
            getattr(lib, prefix + "piece_info")(h, k, b, info)
            lay.info[(k, b)] = tuple(int(v) for v in info)
            nnz, nr_ci = info[2], info[3]
            nwords = nr_ci + (nnz + ratio_v - 1) // ratio_v
            # the reference allocates nr_ci + floor(nnz/ratio_v) words (Q1): never read past that
            avail = nwords if words_fn is None else words_fn(info)
            p = getattr(lib, prefix + "piece_words")(h, k, b)
            if avail:
                arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(avail * 16,)).copy()
            else:
                arr = np.zeros(0, np.uint8)
            if avail < nwords:
                arr = np.concatenate([arr, np.zeros((nwords - avail) * 16, np.uint8)])
            lay.words[(k, b)] = arr
    return lay


class RefLib:
    """The compiled, unmodified reference for one (CU, VF, DOUBLE)."""

    def __init__(self, cu, vf, is_double):
        self.cu, self.vf, self.is_double = cu, vf, bool(is_double)
        L = ctypes.CDLL(ref_so_path(cu, vf, is_double))
        L.ref_build.restype = _vp
        L.ref_build.argtypes = [ctypes.c_uint32] * 3 + [_vp] * 3
        L.ref_piece_words.restype = _vp
        L.ref_piece_words.argtypes = [_vp, ctypes.c_int, ctypes.c_int]
        L.ref_piece_info.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp]
        L.ref_bitmap_row.restype = _vp
        L.ref_bitmap_row.argtypes = [_vp, ctypes.c_int]
        L.ref_blocks.argtypes = [_vp]
        L.ref_make_hw_x.argtypes = [_vp, _vp, ctypes.c_uint32]
        L.ref_hw_x_words.restype = _vp
        L.ref_hw_x_words.argtypes = [_vp, ctypes.c_int]
        L.ref_hw_x_nr_values.restype = ctypes.c_uint32
        L.ref_hw_x_nr_values.argtypes = [_vp, ctypes.c_int]
        L.ref_spmv_hw.argtypes = [_vp, _vp, ctypes.c_uint32]
        L.ref_spmv_gold.argtypes = [ctypes.c_uint32] * 3 + [_vp] * 5
        L.ref_verification.argtypes = [ctypes.c_uint32, _vp, _vp]
        L.ref_storage_overhead_mb.restype = ctypes.c_double
        L.ref_storage_overhead_mb.argtypes = [_vp]
        L.ref_read_matrix_file.argtypes = [ctypes.c_char_p, _vp, _vp, _vp, _vp]
        L.ref_free.argtypes = [_vp]
        assert L.ref_cu() == cu and L.ref_vf() == vf and L.ref_double() == int(is_double)
        self.L = L

    def build(self, rows, cols, row_ptr, col_ind, values):
        rp = np.ascontiguousarray(row_ptr, np.uint32)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(self.is_double))
        h = self.L.ref_build(rows, cols, len(ci), _ptr(rp), _ptr(ci), _ptr(va))
        if not h:
            raise RuntimeError("reference create_csr_hw_matrix failed")
        return h

    def snapshot(self, h, rows):
        ratio_v = 2 if self.is_double else 4
        return _snapshot(self.L, h, "ref_", self.cu, self.vf, self.is_double, rows,
                         words_fn=lambda info: info[3] + info[2] // ratio_v)

    def hw_x(self, h, x):
        xx = np.ascontiguousarray(x, vdtype(self.is_double))
        self.L.ref_make_hw_x(h, _ptr(xx), len(xx))
        out = []
        vt = vdtype(self.is_double)
        for b in range(self.L.ref_blocks(h)):
            n = self.L.ref_hw_x_nr_values(h, b)
            p = self.L.ref_hw_x_words(h, b)
            out.append(np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)),
                                             shape=(n * np.dtype(vt).itemsize,)).copy().view(vt))
        return out

    def spmv_hw(self, h, x, y):
        """y += A x through the emulated FPGA path. Returns 0, or 1 on FIFO under-run (Q1)."""
        xx = np.ascontiguousarray(x, vdtype(self.is_double))
        self.L.ref_make_hw_x(h, _ptr(xx), len(xx))
        assert y.dtype == vdtype(self.is_double) and y.flags.c_contiguous
        return self.L.ref_spmv_hw(h, _ptr(y), len(y))

    def spmv_gold(self, rows, cols, row_ptr, col_ind, values, x):
        rp = np.ascontiguousarray(row_ptr, np.uint32)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(self.is_double))
        xx = np.ascontiguousarray(x, vdtype(self.is_double))
        y = np.zeros(rows, vdtype(self.is_double))
        self.L.ref_spmv_gold(rows, cols, len(ci), _ptr(rp), _ptr(ci), _ptr(va), _ptr(xx), _ptr(y))
        return y

    def read_matrix_file(self, path):
        hdr = np.zeros(4, np.uint32)
        rc = self.L.ref_read_matrix_file(path.encode(), _ptr(hdr), None, None, None)
        if rc:
            return rc, None
        rows, cols, nnz, blocks = (int(v) for v in hdr)
        rp = np.zeros(rows + 1, np.uint32)
        ci = np.zeros(nnz, np.uint32)
        va = np.zeros(nnz, vdtype(self.is_double))
        rc = self.L.ref_read_matrix_file(path.encode(), _ptr(hdr), _ptr(rp), _ptr(ci), _ptr(va))
        return rc, (rows, cols, nnz, blocks, rp, ci, va)

    def free(self, h):
        self.L.ref_free(h)


class OracleLib:
    """The C restatement (oracle/spmv_oracle.c)."""

    def __init__(self):
        L = ctypes.CDLL(build_port())
        L.orc_layout_build.restype = _vp
        L.orc_layout_build.argtypes = [ctypes.c_uint32, ctypes.c_uint32, _vp, _vp, _vp,
                                       ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_uint32]
        L.orc_layout_free.argtypes = [_vp]
        L.orc_blocks.argtypes = [_vp]
        L.orc_expanded_cols.restype = ctypes.c_uint32
        L.orc_expanded_cols.argtypes = [_vp]
        L.orc_piece_info.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _vp]
        L.orc_piece_words.restype = _vp
        L.orc_piece_words.argtypes = [_vp, ctypes.c_int, ctypes.c_int]
        L.orc_bitmap_row.restype = _vp
        L.orc_bitmap_row.argtypes = [_vp, ctypes.c_int]
        L.orc_hw_x.argtypes = [_vp, _vp, ctypes.c_uint32, _vp]
        L.orc_spmv_emu.argtypes = [_vp, _vp, ctypes.c_uint32, _vp]
        L.orc_spmv_gold.argtypes = [ctypes.c_uint32, _vp, _vp, _vp, _vp, _vp, ctypes.c_int]
        L.orc_spmv_gold_omp.argtypes = [ctypes.c_uint32, _vp, _vp, _vp, _vp, _vp, ctypes.c_int]
        L.orc_abs_ax.argtypes = [ctypes.c_uint32, _vp, _vp, _vp, _vp, _vp, ctypes.c_int]
        L.orc_verification.argtypes = [ctypes.c_uint32, _vp, _vp, ctypes.c_int]
        self.L = L

    def build(self, rows, cols, row_ptr, col_ind, values, cu, vf, is_double, cols_div_blocks=0):
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        h = self.L.orc_layout_build(rows, cols, _ptr(rp), _ptr(ci), _ptr(va), cu, vf, int(is_double), cols_div_blocks)
        return h

    def snapshot(self, h, rows, cu, vf, is_double):
        class _Shim:
            pass
        shim = _Shim()
        shim.orc_blocks = self.L.orc_blocks
        shim.orc_bitmap_row = self.L.orc_bitmap_row
        shim.orc_piece_info = self.L.orc_piece_info
        shim.orc_piece_words = self.L.orc_piece_words
        return _snapshot(shim, h, "orc_", cu, vf, is_double, rows)

    def expanded_cols(self, h):
        return self.L.orc_expanded_cols(h)

    def hw_x(self, h, x, is_double):
        xx = np.ascontiguousarray(x, vdtype(is_double))
        out = np.zeros(self.L.orc_expanded_cols(h), vdtype(is_double))
        self.L.orc_hw_x(h, _ptr(xx), len(xx), _ptr(out))
        return out

    def spmv_emu(self, h, x, y, is_double):
        xx = np.ascontiguousarray(x, vdtype(is_double))
        assert y.dtype == vdtype(is_double) and y.flags.c_contiguous
        return self.L.orc_spmv_emu(h, _ptr(xx), len(xx), _ptr(y))

    def spmv_gold(self, rows, row_ptr, col_ind, values, x, is_double):
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        xx = np.ascontiguousarray(x, vdtype(is_double))
        y = np.zeros(rows, vdtype(is_double))
        self.L.orc_spmv_gold(rows, _ptr(rp), _ptr(ci), _ptr(va), _ptr(xx), _ptr(y), int(is_double))
        return y

    def spmv_gold_omp(self, rows, row_ptr, col_ind, values, x, is_double):
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        xx = np.ascontiguousarray(x, vdtype(is_double))
        y = np.zeros(rows, vdtype(is_double))
        threads = self.L.orc_spmv_gold_omp(rows, _ptr(rp), _ptr(ci), _ptr(va), _ptr(xx), _ptr(y), int(is_double))
        return y, threads

    def abs_ax(self, rows, row_ptr, col_ind, values, x, is_double):
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        xx = np.ascontiguousarray(x, vdtype(is_double))
        out = np.zeros(rows, np.float64)
        self.L.orc_abs_ax(rows, _ptr(rp), _ptr(ci), _ptr(va), _ptr(xx), _ptr(out), int(is_double))
        return out

    def free(self, h):
        self.L.orc_layout_free(h)


def layouts_equal(a, b):
    """Field-by-field + masked byte compare of two Layout snapshots. Returns list of differences."""
    diffs = []
    if a.blocks != b.blocks:
        return ["blocks %d != %d" % (a.blocks, b.blocks)]
    for blk in range(a.blocks):
        if not np.array_equal(a.bitmap[blk], b.bitmap[blk]):
            diffs.append("bitmap block %d" % blk)
        for k in range(a.cu):
            if a.info[(k, blk)] != b.info[(k, blk)]:
                diffs.append("info cu %d block %d: %s != %s" % (k, blk, a.info[(k, blk)], b.info[(k, blk)]))
                continue
            wa, wb = a.masked_words(k, blk), b.masked_words(k, blk)
            if not np.array_equal(wa, wb):
                bad = np.nonzero(wa != wb)[0]
                diffs.append("words cu %d block %d: %d bytes differ, first at %d" % (k, blk, len(bad), bad[0]))
    return diffs
