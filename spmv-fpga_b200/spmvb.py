"""ctypes binding of the C ABI (include/spmvb.h) for tests, bench.py and the multi-GPU host driver.

This module is plumbing only: every compute call goes into lib/libspmvb.so (hand-written sm_100a CUDA).  It never
falls back to a CPU implementation - a missing library or a missing GPU raises.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SPMVB_LIB") or os.path.join(HERE, "lib", "libspmvb.so")  # SPMVB_LIB: A/B builds

_vp = ctypes.c_void_p
_u32 = ctypes.c_uint32
_u64 = ctypes.c_uint64
_int = ctypes.c_int

# every symbol include/spmvb.h declares: (restype, argtypes)
SYMBOLS = {
    "spmvb_last_error": (ctypes.c_char_p, []),
    "spmvb_version": (_int, []),
    "spmvb_set_option": (_int, [ctypes.c_char_p, ctypes.c_int64]),
    "spmvb_get_option": (ctypes.c_int64, [ctypes.c_char_p]),
    "spmvb_options_from_env": (_int, []),
    "spmvb_layout_device_params": (_int, [_vp, _vp]),
    "spmvb_layout_x_lines_per_chunk": (ctypes.c_double, [_vp]),
    "spmvb_layout_ell_params": (_int, [_vp, _vp]),
    "spmvb_layout_ell_image": (ctypes.c_int64, [_vp, _vp, ctypes.c_uint64]),
    "spmvb_layout_ell_decode": (ctypes.c_int64, [_vp, _vp, _vp, ctypes.c_uint64]),
    "spmvb_layout_wide_params": (_int, [_vp, _vp]),
    "spmvb_layout_wide_decode": (ctypes.c_int64, [_vp, _vp, _vp, _vp, ctypes.c_uint64]),
    "spmvb_debug_bounds_errors": (_int, [_vp]),
    "spmvb_engine_last_iter_ms": (ctypes.c_float, [_vp]),
    "spmvb_engine_device_layout": (_int, [_vp, _vp]),
    "spmvb_engine_ell_image": (ctypes.c_int64, [_vp, _vp, ctypes.c_uint64]),
    "spmvb_layout_build": (_int, [_u32, _u32, _vp, _vp, _vp, _int, _int, _int, _u32, _vp]),
    "spmvb_layout_build_u32": (_int, [_u32, _u32, _vp, _vp, _vp, _int, _int, _int, _u32, _vp]),
    "spmvb_layout_free": (None, [_vp]),
    "spmvb_layout_blocks": (_int, [_vp]),
    "spmvb_layout_n_cu": (_int, [_vp]),
    "spmvb_layout_rows": (_u32, [_vp]),
    "spmvb_layout_cols": (_u32, [_vp]),
    "spmvb_layout_expanded_cols": (_u32, [_vp]),
    "spmvb_layout_real_nnz": (_u64, [_vp]),
    "spmvb_layout_padded_nnz": (_u64, [_vp]),
    "spmvb_layout_pairs": (_u64, [_vp]),
    "spmvb_layout_stream_bytes": (_u64, [_vp]),
    "spmvb_layout_zero_rows": (ctypes.c_int64, [_vp]),
    "spmvb_layout_piece_info": (_int, [_vp, _int, _int, _vp]),
    "spmvb_layout_piece_words": (_vp, [_vp, _int, _int]),
    "spmvb_layout_bitmap_row": (_int, [_vp, _int, _vp]),
    "spmvb_layout_storage_mb": (ctypes.c_double, [_vp, _int]),
    "spmvb_layout_pack_x": (_int, [_vp, _vp, _u32, _vp]),
    "spmvb_layout_xs_plan": (ctypes.c_int64, [_vp, _int, _int, _vp, _u64, _vp]),
    "spmvb_layout_x_ranges": (ctypes.c_int64, [_vp, _vp, _u64]),
    "spmvb_layout_chunks": (_u64, [_vp]),
    "spmvb_layout_chunk_cols": (_int, [_vp, _u64, _vp, _vp, _vp]),
    "spmvb_layout_equal": (_int, [_vp, _vp, _vp, ctypes.c_size_t]),
    "spmvb_partition_rows": (_int, [_u32, _vp, _int, _int, _vp]),
    "spmvb_engine_create": (_int, [_vp, _int, _int, _vp]),
    "spmvb_engine_create_from_csr": (_int, [_u32, _u32, _vp, _vp, _vp, _int, _int, _int, _u32, _int, _int, _int, _vp, _vp]),
    "spmvb_engine_fetch_layout": (_int, [_vp, _vp]),
    "spmvb_engine_build_ms": (_int, [_vp, _vp]),
    "spmvb_engine_free": (None, [_vp]),
    "spmvb_engine_set_variant": (_int, [_vp, _int]),
    "spmvb_engine_variant": (_int, [_vp]),
    "spmvb_engine_launches": (_u64, [_vp]),
    "spmvb_engine_algorithmic_bytes": (_u64, [_vp]),
    "spmvb_engine_x_upload_bytes": (_u64, [_vp]),
    "spmvb_engine_x_dev": (_vp, [_vp]),
    "spmvb_engine_y_dev": (_vp, [_vp]),
    "spmvb_engine_stream": (_vp, [_vp]),
    "spmvb_engine_set_x": (_int, [_vp, _vp, _u32]),
    "spmvb_engine_spmv_dev": (_int, [_vp, _vp, _vp, _int, _vp]),
    "spmvb_engine_sync": (_int, [_vp]),
    "spmvb_engine_get_y": (_int, [_vp, _vp, _u32, _int]),
    "spmvb_engine_spmv_host": (_int, [_vp, _vp, _u32, _vp, _int]),
    "spmvb_engine_spmv_host_x_resident": (_int, [_vp, _vp, _int]),
    "spmvb_engine_time_spmv": (_int, [_vp, _int, _int, _vp]),
    "spmvb_engine_enqueue_steps": (_int, [_vp, _int, _int]),
    "spmvb_engine_steps_done": (_int, [_vp]),
    "spmvb_engine_collect_steps": (_int, [_vp, _vp, _vp]),
    "spmvb_engine_power_iter": (_int, [_vp, _int, _vp]),
    "spmvb_engine_cg": (_int, [_vp, _vp, _vp, _int, ctypes.c_double, _vp, _vp]),
    "spmvb_engine_scale_copy": (_int, [_vp, _vp, _vp, _u32, ctypes.c_double, _vp]),
    "spmvb_engine_scale_rsqrt": (_int, [_vp, _vp, _vp, _u32, _vp, _vp]),
    "spmvb_engine_sumsq": (_int, [_vp, _vp, _u32, _vp, _vp]),
    "spmvb_group_create": (_int, [_u32, _u32, _vp, _vp, _vp, _int, _int, _vp, _int, _vp]),
    "spmvb_group_unique_id": (_int, [_vp]),
    "spmvb_group_create_rank": (_int, [_u32, _u32, _vp, _vp, _vp, _vp, _int, _int, _int, _vp, _int, _int, _vp]),
    "spmvb_group_free": (None, [_vp]),
    "spmvb_group_world": (_int, [_vp]),
    "spmvb_group_local_count": (_int, [_vp]),
    "spmvb_group_bounds": (_int, [_vp, _vp]),
    "spmvb_group_engine": (_vp, [_vp, _int]),
    "spmvb_group_rank": (_int, [_vp, _int]),
    "spmvb_group_spmv_host": (_int, [_vp, _vp, _u32, _vp, _int]),
    "spmvb_group_spmv_host_rows": (_int, [_vp, _vp, _u32, _vp, _int]),
    "spmvb_group_x_over_links": (_int, [_vp]),
    "spmvb_group_adopt_engine": (_int, [_u32, _u32, _vp, _vp, _int, _int, _vp, _int, _int, _vp]),
    "spmvb_group_set_x": (_int, [_vp, _vp, _u32]),
    "spmvb_group_get_x": (_int, [_vp, _vp, _u32]),
    "spmvb_group_get_y": (_int, [_vp, _vp]),
    "spmvb_group_power_iter": (_int, [_vp, _int, _vp]),
    "spmvb_group_last_iter_ms": (ctypes.c_float, [_vp]),
    "spmvb_group_phase_ms": (_int, [_vp, _vp]),
    "spmvb_group_ipc_handle": (_int, [_vp, _vp]),
    "spmvb_group_set_peer_handles": (_int, [_vp, _vp, _int]),
    "spmvb_group_set_exchange": (_int, [_vp, _int]),
    "spmvb_group_exchange": (_int, [_vp]),
    "spmvb_engine_x_len": (_u64, [_vp]),
    "spmvb_csr_free": (None, [_vp]),
    "spmvb_csr_rows": (_u32, [_vp]),
    "spmvb_csr_cols": (_u32, [_vp]),
    "spmvb_csr_nnz": (_u64, [_vp]),
    "spmvb_csr_is_double": (_int, [_vp]),
    "spmvb_csr_row_ptr": (_vp, [_vp]),
    "spmvb_csr_col_ind": (_vp, [_vp]),
    "spmvb_csr_values": (_vp, [_vp]),
    "spmvb_layout_build_csr": (_int, [_vp, _int, _int, _u32, _vp]),
    "spmvb_csr_read": (_int, [ctypes.c_char_p, _int, _vp]),
    "spmvb_csr_write": (_int, [_vp, ctypes.c_char_p]),
    "spmvb_csr_save": (_int, [_vp, ctypes.c_char_p]),
    "spmvb_csr_load": (_int, [ctypes.c_char_p, _vp]),
    "spmvb_csr_read_cached": (_int, [ctypes.c_char_p, _int, _vp]),
    "spmvb_csr_gen_band": (_int, [_u32, _int, _u64, _int, _vp]),
    "spmvb_csr_gen_laplacian2d": (_int, [_u32, _u32, _u32, _u32, _int, _vp]),
    "spmvb_csr_gen_uniform": (_int, [_u32, _u32, _int, _u64, _u32, _u32, _int, _vp]),
    "spmvb_csr_gen_rmat": (_int, [_int, _int, ctypes.c_double, ctypes.c_double, ctypes.c_double, _u64, _u32, _u32,
                                  _int, _vp]),
}

_lib = None


class SpmvbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("spmvb error %d: %s" % (code, msg))
        self.code = code


def build_library(force=False):
    """Compile lib/libspmvb.so in-tree (nvcc, sm_100a)."""
    if force or not os.path.exists(LIB_PATH):
        subprocess.check_call(["make", "-C", HERE] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SpmvbError(-3, "libspmvb.so is not built (run `make -C spmv-fpga_b200` or __graft_entry__.build()); "
                                 "there is no fallback path")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


GEN_LIB_PATH = os.path.join(os.path.dirname(HERE), "oracle", "_ref", "libmatgen.so")
_GEN_SYMBOLS = [n for n in SYMBOLS if n.startswith("spmvb_csr_gen_") or n in (
    "spmvb_last_error", "spmvb_csr_free", "spmvb_csr_rows", "spmvb_csr_cols", "spmvb_csr_nnz", "spmvb_csr_is_double",
    "spmvb_csr_row_ptr", "spmvb_csr_col_ind", "spmvb_csr_values")]
_gen_lib = None


def gen_lib():
    """The synthetic-matrix generators alone (oracle/_ref/libmatgen.so: matrix_gen.cpp built without the engine).  The
    reference arm of bench.py generates its workload through this handle, so that it never loads libspmvb.so."""
    global _gen_lib
    if _gen_lib is None:
        if not os.path.exists(GEN_LIB_PATH):
            subprocess.check_call(["make", "-C", os.path.dirname(os.path.dirname(GEN_LIB_PATH)), "gen"],
                                  stdout=subprocess.DEVNULL)
        L = ctypes.CDLL(GEN_LIB_PATH)
        for name in _GEN_SYMBOLS:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = SYMBOLS[name]
        _gen_lib = L
    return _gen_lib


def _check(rc, L=None):
    if rc != 0:
        raise SpmvbError(rc, (L or lib()).spmvb_last_error().decode(errors="replace"))


def set_option(name, value):
    """Process-wide tuning option (include/spmvb.h: spmvb_set_option); -1 restores the library's own choice."""
    _check(lib().spmvb_set_option(name.encode(), int(value)))


def bounds_errors():
    """Bounds-checked build only (SPMVB_LIB=.../libspmvb_check.so): the kernels' out-of-range counters, else None."""
    out = (ctypes.c_uint64 * 5)()
    rc = lib().spmvb_debug_bounds_errors(out)
    return None if rc < 0 else dict(zip(("chunk", "rowmap", "y_row", "x_index", "x_window"), (int(v) for v in out)))


def get_option(name):
    return int(lib().spmvb_get_option(name.encode()))


class options:
    """with spmvb.options(cu_major=1, tall=1): ...  - sets the options and restores the previous values on exit."""

    def __init__(self, **kw):
        self.kw = kw

    def __enter__(self):
        self.old = {k: get_option(k) for k in self.kw}
        for k, v in self.kw.items():
            set_option(k, v)
        return self

    def __exit__(self, *exc):
        for k, v in self.old.items():
            set_option(k, v)


def _ptr(a):
    return a.ctypes.data_as(_vp) if a is not None else None


def vdtype(is_double):
    return np.float64 if is_double else np.float32


class Csr:
    """Library-owned CSR matrix (numpy views are zero-copy and valid while the object lives).  L = the ctypes library
    that owns it: libspmvb.so by default, gen_lib() for matrices generated without the engine."""

    def __init__(self, handle, L=None):
        self.h = _vp(handle)
        self.L = L = L or lib()
        self.rows = L.spmvb_csr_rows(self.h)
        self.cols = L.spmvb_csr_cols(self.h)
        self.nnz = L.spmvb_csr_nnz(self.h)
        self.is_double = bool(L.spmvb_csr_is_double(self.h))

    def _view(self, p, n, dt):
        if n == 0:
            return np.zeros(0, dt)
        buf = (ctypes.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(p)
        return np.frombuffer(buf, dtype=dt, count=n)

    @property
    def row_ptr(self):
        return self._view(self.L.spmvb_csr_row_ptr(self.h), self.rows + 1, np.uint64)

    @property
    def col_ind(self):
        return self._view(self.L.spmvb_csr_col_ind(self.h), self.nnz, np.uint32)

    @property
    def values(self):
        return self._view(self.L.spmvb_csr_values(self.h), self.nnz, vdtype(self.is_double))

    def write(self, path):
        _check(lib().spmvb_csr_write(self.h, path.encode()))

    def save(self, path):
        """Binary form (header + the three arrays as they sit in memory)."""
        _check(lib().spmvb_csr_save(self.h, path.encode()))

    def free(self):
        if self.h:
            self.L.spmvb_csr_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    # ---- constructors
    @staticmethod
    def _new(L, name, *args):
        L = L or lib()
        out = _vp()
        _check(getattr(L, name)(*args, ctypes.byref(out)), L)
        return Csr(out.value, L)

    @staticmethod
    def read(path, is_double=True):
        return Csr._new(None, "spmvb_csr_read", path.encode(), int(is_double))

    @staticmethod
    def load(path):
        return Csr._new(None, "spmvb_csr_load", path.encode())

    @staticmethod
    def read_cached(path, is_double=True):
        """Parses the text file once and keeps a binary sidecar next to it for the following calls."""
        return Csr._new(None, "spmvb_csr_read_cached", path.encode(), int(is_double))

    @staticmethod
    def band(n, half_bw=5, seed=1, is_double=True, L=None):
        return Csr._new(L, "spmvb_csr_gen_band", n, half_bw, seed, int(is_double))

    @staticmethod
    def laplacian2d(nx, ny, row_begin=0, row_end=0, is_double=True, L=None):
        return Csr._new(L, "spmvb_csr_gen_laplacian2d", nx, ny, row_begin, row_end, int(is_double))

    @staticmethod
    def uniform(rows, cols, nnz_per_row, seed=1, row_begin=0, row_end=0, is_double=True, L=None):
        return Csr._new(L, "spmvb_csr_gen_uniform", rows, cols, nnz_per_row, seed, row_begin, row_end, int(is_double))

    @staticmethod
    def rmat(scale, edge_factor=16, a=0.57, b=0.19, c=0.19, seed=1, row_begin=0, row_end=0, is_double=True, L=None):
        return Csr._new(L, "spmvb_csr_gen_rmat", scale, edge_factor, a, b, c, seed, row_begin, row_end, int(is_double))


class Layout:
    """Host hw_matrix layout (create_csr_hw_matrix)."""

    def __init__(self, handle, is_double):
        self.h = _vp(handle)
        self.is_double = bool(is_double)
        L = lib()
        self.blocks = L.spmvb_layout_blocks(self.h)
        self.n_cu = L.spmvb_layout_n_cu(self.h)
        self.rows = L.spmvb_layout_rows(self.h)
        self.cols = L.spmvb_layout_cols(self.h)
        self.expanded_cols = L.spmvb_layout_expanded_cols(self.h)
        self.real_nnz = L.spmvb_layout_real_nnz(self.h)
        self.padded_nnz = L.spmvb_layout_padded_nnz(self.h)
        self.pairs = L.spmvb_layout_pairs(self.h)
        self.stream_bytes = L.spmvb_layout_stream_bytes(self.h)
        self.zero_rows = L.spmvb_layout_zero_rows(self.h)

    @staticmethod
    def build(rows, cols, row_ptr, col_ind, values, n_cu=1, vf=1, is_double=True, cols_div_blocks=0):
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        out = _vp()
        _check(lib().spmvb_layout_build(rows, cols, _ptr(rp), _ptr(ci), _ptr(va), n_cu, vf, int(is_double),
                                        cols_div_blocks, ctypes.byref(out)))
        return Layout(out.value, is_double)

    @staticmethod
    def build_u32(rows, cols, row_ptr, col_ind, values, n_cu=1, vf=1, is_double=True, cols_div_blocks=0):
        rp = np.ascontiguousarray(row_ptr, np.uint32)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        out = _vp()
        _check(lib().spmvb_layout_build_u32(rows, cols, _ptr(rp), _ptr(ci), _ptr(va), n_cu, vf, int(is_double),
                                            cols_div_blocks, ctypes.byref(out)))
        return Layout(out.value, is_double)

    @staticmethod
    def from_csr(csr, n_cu=1, vf=1, cols_div_blocks=0):
        out = _vp()
        _check(lib().spmvb_layout_build_csr(csr.h, n_cu, vf, cols_div_blocks, ctypes.byref(out)))
        return Layout(out.value, csr.is_double)

    @property
    def device_params(self):
        """What the GPU streams: dict(cu, vf, cdb, cu_major, private, pairs, chunks, zero_rows (-1 = all), bytes)."""
        out = (ctypes.c_uint64 * 9)()
        _check(lib().spmvb_layout_device_params(self.h, out))
        v = [int(x) for x in out]
        return dict(cu=v[0], vf=v[1], cdb=v[2], cu_major=bool(v[3]), private=bool(v[4]), pairs=v[5], chunks=v[6],
                    zero_rows=-1 if v[7] == 2 ** 64 - 1 else v[7], bytes=v[8])

    @property
    def ell_params(self):
        """The sliced-ELLPACK image (regular matrices): dict(present, width, slices, slice_bytes, slots, bytes, nnz)."""
        out = (ctypes.c_uint64 * 8)()
        _check(lib().spmvb_layout_ell_params(self.h, out))
        v = [int(x) for x in out]
        return dict(present=bool(v[0]), width=v[1], slices=v[2], slice_bytes=v[3], slots=v[4], bytes=v[5], nnz=v[6])

    def ell_image(self):
        """The bytes of the host-built ELL image (None without one)."""
        n = lib().spmvb_layout_ell_image(self.h, None, 0)
        if n <= 0:
            return None
        out = np.zeros(n, np.uint8)
        assert lib().spmvb_layout_ell_image(self.h, _ptr(out), n) == n
        return out

    def ell_decode(self):
        """(cols, values) of every slot of the ELL image, row-major: arrays of shape (slices * 32, width)."""
        e = self.ell_params
        if not e["present"]:
            raise SpmvbError(-1, "no ELL image")
        n = e["slots"]
        cols = np.zeros(n, np.uint32)
        vals = np.zeros(n, np.float64 if self.is_double else np.float32)
        got = lib().spmvb_layout_ell_decode(self.h, _ptr(cols), _ptr(vals), n)
        if got != n:
            raise SpmvbError(int(got), lib().spmvb_last_error().decode(errors="replace"))
        return cols.reshape(-1, e["width"]), vals.reshape(-1, e["width"])

    @property
    def wide_params(self):
        """The wide image: dict(present, cdb, blocks, pairs, chunks, zero_rows (-1 = all), bytes, nnz)."""
        out = (ctypes.c_uint64 * 8)()
        _check(lib().spmvb_layout_wide_params(self.h, out))
        v = [int(x) for x in out]
        return dict(present=bool(v[0]), cdb=v[1], blocks=v[2], pairs=v[3], chunks=v[4],
                    zero_rows=-1 if v[5] == 2 ** 64 - 1 else v[5], bytes=v[6], nnz=v[7])

    def wide_decode(self):
        """(rows, cols, values) of the wide image's entries in image order, after the library's own consistency walk."""
        w = self.wide_params
        if not w["present"]:
            raise SpmvbError(-1, "no wide image")
        n = w["nnz"]
        rows = np.zeros(max(n, 1), np.uint32)
        cols = np.zeros(max(n, 1), np.uint32)
        vals = np.zeros(max(n, 1), np.float64 if self.is_double else np.float32)
        got = lib().spmvb_layout_wide_decode(self.h, _ptr(rows), _ptr(cols), _ptr(vals), n)
        if got < 0:
            raise SpmvbError(int(got), lib().spmvb_last_error().decode(errors="replace"))
        if got != n:
            raise SpmvbError(-1, "wide image holds %d entries, expected %d" % (got, n))
        return rows[:n], cols[:n], vals[:n]

    @property
    def x_lines_per_chunk(self):
        return float(lib().spmvb_layout_x_lines_per_chunk(self.h))

    def difference(self, other):
        """'' if both layouts are identical in every table and byte, else the first component that differs."""
        why = ctypes.create_string_buffer(256)
        rc = lib().spmvb_layout_equal(self.h, other.h, why, 256)
        if rc < 0:
            _check(rc)
        return "" if rc == 1 else (why.value.decode() or "different")

    def piece_info(self, cu, block):
        info = (ctypes.c_uint32 * 5)()
        _check(lib().spmvb_layout_piece_info(self.h, cu, block, info))
        return tuple(int(v) for v in info)

    def piece_words(self, cu, block):
        nr_rows, nr_cols, nnz, nr_ci, nr_val = self.piece_info(cu, block)
        ratio_v = 2 if self.is_double else 4
        nwords = nr_ci + (nnz + ratio_v - 1) // ratio_v
        p = lib().spmvb_layout_piece_words(self.h, cu, block)
        if not p:
            raise SpmvbError(-1, lib().spmvb_last_error().decode(errors="replace"))
        if nwords == 0:
            return np.zeros(0, np.uint8)
        return np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(nwords * 16,)).copy()

    def bitmap_row(self, block):
        out = np.zeros(self.rows, np.uint8)
        _check(lib().spmvb_layout_bitmap_row(self.h, block, _ptr(out)))
        return out

    def storage_mb(self, cu):
        return lib().spmvb_layout_storage_mb(self.h, cu)

    def pack_x(self, x):
        xx = np.ascontiguousarray(x, vdtype(self.is_double))
        out = np.empty(self.expanded_cols, vdtype(self.is_double))
        _check(lib().spmvb_layout_pack_x(self.h, _ptr(xx), len(xx), _ptr(out)))
        return out

    def xs_plan(self, n_cta=148, run_log2=1):
        """Work plan of the shared-memory-x kernel: (items[n, 8] uint32, cta_first[n_cta + 1])."""
        n = lib().spmvb_layout_xs_plan(self.h, n_cta, run_log2, None, 0, None)
        if n < 0:
            _check(int(n))
        items = np.zeros((max(n, 1), 8), np.uint32)
        first = np.zeros(n_cta + 1, np.uint32)
        lib().spmvb_layout_xs_plan(self.h, n_cta, run_log2, _ptr(items), n, _ptr(first))
        return items[:n], first

    def x_ranges(self):
        """[first, end) column ranges of x that a SpMV with this layout can read (whole column blocks)."""
        n = lib().spmvb_layout_x_ranges(self.h, None, 0)
        if n < 0:
            _check(int(n))
        out = np.zeros((int(n), 2), np.uint64)
        if n:
            lib().spmvb_layout_x_ranges(self.h, _ptr(out), int(n))
        return out

    @property
    def n_chunks(self):
        return lib().spmvb_layout_chunks(self.h)

    def chunk_cols(self, c):
        lo, hi, blk = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_uint32()
        _check(lib().spmvb_layout_chunk_cols(self.h, c, ctypes.byref(lo), ctypes.byref(hi), ctypes.byref(blk)))
        return lo.value, hi.value, blk.value

    def free(self):
        if self.h:
            lib().spmvb_layout_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def partition_rows(rows, row_ptr, parts, ratio_v=2):
    rp = np.ascontiguousarray(row_ptr, np.uint64)
    bounds = np.zeros(parts + 1, np.uint32)
    _check(lib().spmvb_partition_rows(rows, _ptr(rp), parts, ratio_v, _ptr(bounds)))
    return bounds


VARIANT_AUTO, VARIANT_DIRECT, VARIANT_OCC4, VARIANT_OCC3, VARIANT_XS, VARIANT_WIDE, VARIANT_ELL = 0, 1, 6, 7, 8, 9, 10


class Engine:
    """Device-resident hw_matrix + the SpMV kernels (spmv_hw)."""

    def __init__(self, layout, device=0, variant=VARIANT_AUTO):
        out = _vp()
        _check(lib().spmvb_engine_create(layout.h, device, variant, ctypes.byref(out)))
        self.h = _vp(out.value)
        self.is_double = layout.is_double
        self.rows, self.cols, self.expanded_cols = layout.rows, layout.cols, layout.expanded_cols
        self.real_nnz = layout.real_nnz

    @staticmethod
    def from_csr(rows, cols, row_ptr, col_ind, values, n_cu=1, vf=1, is_double=True, cols_div_blocks=0, device=0,
                 variant=VARIANT_AUTO, on_device=False):
        """create_csr_hw_matrix on the GPU: returns (layout, engine).  With on_device the three arrays are device
        pointers (ints); otherwise numpy arrays that are uploaded inside the call."""
        lay, eng = _vp(), _vp()
        if on_device:
            a = (_vp(row_ptr), _vp(col_ind), _vp(values))
        else:
            rp = np.ascontiguousarray(row_ptr, np.uint64)
            ci = np.ascontiguousarray(col_ind, np.uint32)
            va = np.ascontiguousarray(values, vdtype(is_double))
            a = (_ptr(rp), _ptr(ci), _ptr(va))
        _check(lib().spmvb_engine_create_from_csr(rows, cols, a[0], a[1], a[2], n_cu, vf, int(is_double),
                                                  cols_div_blocks, device, variant, int(on_device),
                                                  ctypes.byref(lay), ctypes.byref(eng)))
        layout = Layout(lay.value, is_double)
        e = Engine.__new__(Engine)
        e.h = _vp(eng.value)
        e.layout = layout
        e.is_double = layout.is_double
        e.rows, e.cols, e.expanded_cols = layout.rows, layout.cols, layout.expanded_cols
        e.real_nnz = layout.real_nnz
        return layout, e

    def fetch_layout(self, layout=None):
        """Copies the pieces and the row map of a GPU-built layout to the host (piece_words / bitmap_row need them)."""
        _check(lib().spmvb_engine_fetch_layout(self.h, (layout or self.layout).h))

    def build_ms(self):
        out = (ctypes.c_float * 3)()
        _check(lib().spmvb_engine_build_ms(self.h, out))
        return {"h2d_ms": out[0], "build_ms": out[1], "total_ms": out[2]}

    @property
    def last_iter_ms(self):
        """Device time per iteration of the last power_iter / cg call (CUDA events)."""
        return float(lib().spmvb_engine_last_iter_ms(self.h))

    @property
    def device_layout(self):
        out = (ctypes.c_uint64 * 20)()
        _check(lib().spmvb_engine_device_layout(self.h, out))
        v = [int(x) for x in out]
        return dict(cu=v[0], vf=v[1], cdb=v[2], cu_major=bool(v[3]), pairs=v[4], chunks=v[5],
                    zero_rows=-1 if v[6] == 2 ** 64 - 1 else v[6], bytes=v[7], e2e_tiles=v[8], tall=bool(v[9]), xs_config=v[10],
                    wide=bool(v[13]), blocks=v[15], ell=bool(v[16]), ell_width=v[18], ell_e2e_tiles=v[19],
                    tuned_us=dict(api_image=v[11], device_layout=v[12], wide_image=v[14], ell_image=v[17]))

    def ell_image(self):
        """The bytes of the sliced-ELLPACK image the engine streams (None when it streams something else)."""
        n = lib().spmvb_engine_ell_image(self.h, None, 0)
        if n <= 0:
            return None
        out = np.zeros(n, np.uint8)
        assert lib().spmvb_engine_ell_image(self.h, _ptr(out), n) == n
        return out

    def set_variant(self, v):
        _check(lib().spmvb_engine_set_variant(self.h, v))

    @property
    def variant(self):
        """The kernel variant in use (variant 0 resolves to the autotuned choice)."""
        return lib().spmvb_engine_variant(self.h)

    @property
    def launches(self):
        return lib().spmvb_engine_launches(self.h)

    @property
    def algorithmic_bytes(self):
        return lib().spmvb_engine_algorithmic_bytes(self.h)

    @property
    def x_upload_bytes(self):
        """Bytes one set_x of a full-length x moves to the device (only the column blocks the matrix touches)."""
        return lib().spmvb_engine_x_upload_bytes(self.h)

    @property
    def x_dev(self):
        return lib().spmvb_engine_x_dev(self.h)

    @property
    def y_dev(self):
        return lib().spmvb_engine_y_dev(self.h)

    @property
    def stream(self):
        return lib().spmvb_engine_stream(self.h)

    def set_x(self, x):
        xx = np.ascontiguousarray(x, vdtype(self.is_double))
        _check(lib().spmvb_engine_set_x(self.h, _ptr(xx), len(xx)))
        _check(lib().spmvb_engine_sync(self.h))

    def spmv_dev(self, x_dev=None, y_dev=None, accumulate=False, stream=None):
        _check(lib().spmvb_engine_spmv_dev(self.h, x_dev, y_dev, int(accumulate), stream))

    def sync(self):
        _check(lib().spmvb_engine_sync(self.h))

    def get_y(self, out=None, accumulate=False):
        if out is None:
            out = np.zeros(self.rows, vdtype(self.is_double))
        _check(lib().spmvb_engine_get_y(self.h, _ptr(out), len(out), int(accumulate)))
        return out

    def spmv_host(self, x, y, accumulate=True):
        """spmv_hw: y (+)= A x with host buffers (x, y numpy arrays or raw host pointers)."""
        if isinstance(x, np.ndarray):
            assert x.dtype == vdtype(self.is_double) and x.flags.c_contiguous
            xp, n = _ptr(x), len(x)
        else:
            xp, n = x
        yp = _ptr(y) if isinstance(y, np.ndarray) else y
        _check(lib().spmvb_engine_spmv_host(self.h, xp, n, yp, int(accumulate)))
        return y

    def time_spmv(self, iters, flush_l2=False):
        ms = np.zeros(iters, np.float32)
        _check(lib().spmvb_engine_time_spmv(self.h, iters, int(flush_l2), _ptr(ms)))
        return ms

    def enqueue_steps(self, steps, flush_l2=False, inner_events=True):
        self._steps = steps if inner_events else 0
        _check(lib().spmvb_engine_enqueue_steps(self.h, steps, int(bool(flush_l2)) | (0 if inner_events else 2)))

    def steps_done(self):
        return bool(lib().spmvb_engine_steps_done(self.h))

    def collect_steps(self):
        total = ctypes.c_float()
        ker = np.zeros(self._steps, np.float32)
        _check(lib().spmvb_engine_collect_steps(self.h, ctypes.byref(total), _ptr(ker)))
        return total.value, ker

    def power_iter(self, iters):
        nrm = ctypes.c_double()
        _check(lib().spmvb_engine_power_iter(self.h, iters, ctypes.byref(nrm)))
        return nrm.value

    def cg(self, b, max_iters=1000, rel_tol=1e-10):
        """Conjugate gradients for A x = b (A symmetric positive definite): returns (x, iterations, ||r|| / ||b||)."""
        bb = np.ascontiguousarray(b, vdtype(self.is_double))
        x = np.zeros(self.rows, vdtype(self.is_double))
        it, rel = ctypes.c_int(), ctypes.c_double()
        _check(lib().spmvb_engine_cg(self.h, _ptr(bb), _ptr(x), max_iters, rel_tol, ctypes.byref(it), ctypes.byref(rel)))
        return x, it.value, rel.value

    def scale_copy(self, src_dev, dst_dev, n, scale, stream=None):
        _check(lib().spmvb_engine_scale_copy(self.h, src_dev, dst_dev, n, scale, stream))

    def scale_rsqrt(self, src_dev, dst_dev, n, sumsq_dev, stream=None):
        _check(lib().spmvb_engine_scale_rsqrt(self.h, src_dev, dst_dev, n, sumsq_dev, stream))

    def sumsq(self, src_dev, n, out_dev, stream=None):
        _check(lib().spmvb_engine_sumsq(self.h, src_dev, n, out_dev, stream))

    def free(self):
        if self.h:
            lib().spmvb_engine_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Group:
    """Engines over one matrix on several GPUs (include/spmvb.h: spmvb_group_*): rows owned in contiguous ranges
    balanced by non-zeros, x replicated.  Group.create drives n GPUs from this process; Group.create_rank makes this
    process one rank of a multi-process group (the NCCL unique id comes from Group.unique_id() on one rank)."""

    def __init__(self, handle, is_double, rows, cols):
        self.h = _vp(handle)
        self.is_double, self.rows, self.cols = bool(is_double), rows, cols
        L = lib()
        self.world = L.spmvb_group_world(self.h)
        self.local_count = L.spmvb_group_local_count(self.h)
        b = np.zeros(self.world + 1, np.uint32)
        _check(L.spmvb_group_bounds(self.h, _ptr(b)))
        self.bounds = b
        self.ranks = [L.spmvb_group_rank(self.h, i) for i in range(self.local_count)]

    @staticmethod
    def create(rows, cols, row_ptr, col_ind, values, is_double=True, devices=(0,), variant=VARIANT_AUTO):
        rp = np.ascontiguousarray(row_ptr, np.uint64)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        dv = np.ascontiguousarray(list(devices), np.int32)
        out = _vp()
        _check(lib().spmvb_group_create(rows, cols, _ptr(rp), _ptr(ci), _ptr(va), int(is_double), len(dv), _ptr(dv), variant,
                                        ctypes.byref(out)))
        return Group(out.value, is_double, rows, cols)

    @staticmethod
    def unique_id():
        out = np.zeros(128, np.uint8)
        _check(lib().spmvb_group_unique_id(_ptr(out)))
        return out

    @staticmethod
    def create_rank(global_rows, cols, bounds, row_ptr_local, col_ind, values, is_double, device, unique_id, rank, world,
                    variant=VARIANT_AUTO):
        b = np.ascontiguousarray(bounds, np.uint32)
        rp = np.ascontiguousarray(row_ptr_local, np.uint64)
        ci = np.ascontiguousarray(col_ind, np.uint32)
        va = np.ascontiguousarray(values, vdtype(is_double))
        uid = np.ascontiguousarray(unique_id, np.uint8) if unique_id is not None else None
        out = _vp()
        _check(lib().spmvb_group_create_rank(global_rows, cols, _ptr(b), _ptr(rp), _ptr(ci), _ptr(va), int(is_double), device,
                                             variant, _ptr(uid), rank, world, ctypes.byref(out)))
        return Group(out.value, is_double, global_rows, cols)

    @staticmethod
    def adopt(engine, global_rows, cols, bounds, device, unique_id, rank, world):
        """One rank of a multi-process group around an Engine this process created (and keeps)."""
        b = np.ascontiguousarray(bounds, np.uint32)
        uid = np.ascontiguousarray(unique_id, np.uint8) if unique_id is not None else None
        out = _vp()
        _check(lib().spmvb_group_adopt_engine(global_rows, cols, _ptr(b), engine.h, int(engine.is_double), device, _ptr(uid),
                                              rank, world, ctypes.byref(out)))
        return Group(out.value, engine.is_double, global_rows, cols)

    def spmv_host_rows(self, x, y_rows, accumulate=True):
        """spmv_hw over the group with y_rows = this process's rows only (x, y_rows numpy arrays or raw host pointers)."""
        if isinstance(x, np.ndarray):
            assert x.dtype == vdtype(self.is_double) and x.flags.c_contiguous
            xp, n = _ptr(x), len(x)
        else:
            xp, n = x
        yp = _ptr(y_rows) if isinstance(y_rows, np.ndarray) else y_rows
        _check(lib().spmvb_group_spmv_host_rows(self.h, xp, n, yp, int(accumulate)))
        return y_rows

    @property
    def x_over_links(self):
        """1: spmv_host uploads 1/world of x per GPU and all-gathers over NVLink; 0: per-GPU uploads; -1: not decided yet."""
        return lib().spmvb_group_x_over_links(self.h)

    def engine_handle(self, i=0):
        return lib().spmvb_group_engine(self.h, i)

    def launches(self):
        return sum(int(lib().spmvb_engine_launches(_vp(self.engine_handle(i)))) for i in range(self.local_count))

    def set_x(self, x):
        xx = np.ascontiguousarray(x, vdtype(self.is_double))
        _check(lib().spmvb_group_set_x(self.h, _ptr(xx), len(xx)))

    def get_x(self):
        out = np.zeros(self.cols, vdtype(self.is_double))
        _check(lib().spmvb_group_get_x(self.h, _ptr(out), len(out)))
        return out

    def get_y(self):
        """Full-length y with the rows of this process's GPUs filled in (the other rows stay 0)."""
        out = np.zeros(self.rows, vdtype(self.is_double))
        _check(lib().spmvb_group_get_y(self.h, _ptr(out)))
        return out

    def spmv_host(self, x, y, accumulate=True):
        assert x.dtype == vdtype(self.is_double) and y.dtype == x.dtype and len(y) == self.rows
        _check(lib().spmvb_group_spmv_host(self.h, _ptr(x), len(x), _ptr(y), int(accumulate)))
        return y

    def power_iter(self, iters):
        nrm = ctypes.c_double()
        _check(lib().spmvb_group_power_iter(self.h, iters, ctypes.byref(nrm)))
        return nrm.value

    @property
    def phase_ms(self):
        """Last iteration of the last power_iter call: [SpMV + sum of squares, norm all-reduce, normalise / store, gather]."""
        out = (ctypes.c_float * 4)()
        _check(lib().spmvb_group_phase_ms(self.h, out))
        return [float(v) for v in out]

    def ipc_handle(self):
        """cudaIpcMemHandle_t (64 bytes) of this rank's x, for the other ranks of a multi-process group."""
        out = np.zeros(64, np.uint8)
        _check(lib().spmvb_group_ipc_handle(self.h, _ptr(out)))
        return out

    def set_peer_handles(self, handles, mode=2):
        """handles: (world, 64) uint8, every rank's ipc_handle() in rank order; mode: see set_exchange."""
        hh = np.ascontiguousarray(handles, np.uint8).reshape(-1)
        assert len(hh) == 64 * self.world
        _check(lib().spmvb_group_set_peer_handles(self.h, _ptr(hh), mode))

    def set_exchange(self, mode):
        """0 NCCL broadcasts, 1 peer stores to every GPU, 2 peer stores to the forwarding GPU + all-gather."""
        _check(lib().spmvb_group_set_exchange(self.h, mode))

    @property
    def exchange(self):
        return lib().spmvb_group_exchange(self.h)

    @property
    def last_iter_ms(self):
        return float(lib().spmvb_group_last_iter_ms(self.h))

    def free(self):
        if self.h:
            lib().spmvb_group_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
