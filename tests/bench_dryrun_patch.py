"""TEST INFRASTRUCTURE ONLY: lets bench.py's plumbing (argument handling, multi-rank reductions, the JSON line) run on a
machine without a GPU.  Importing this module replaces the CUDA engine by a stand-in that does no arithmetic and
returns made-up timings, and points torch.distributed at gloo.  Nothing here is reachable from the product or from a
real bench run; tests/test_bench_dryrun.py is the only user.  The numbers such a run prints mean nothing."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "spmv-fpga_b200"))
import spmvb  # noqa: E402

torch.cuda.is_available = lambda: True
torch.cuda.set_device = lambda *_a, **_k: None
torch.cuda.synchronize = lambda *_a, **_k: None
torch.Tensor.pin_memory = lambda self, *_a, **_k: self

_real_tensor = torch.tensor
_real_zeros = torch.zeros


def _cpu_only(fn):
    def wrapped(*a, **k):
        k.pop("device", None)
        return fn(*a, **k)
    return wrapped


class _Stream:
    cuda_stream = 0


class _Event:
    def __init__(self, *a, **k): pass
    def record(self, *a, **k): pass
    def elapsed_time(self, other): return 1.0


torch.cuda.Stream = _Stream
torch.cuda.set_stream = lambda *_a, **_k: None
torch.cuda.current_stream = lambda *_a, **_k: _Stream()
torch.cuda.Event = _Event
torch.tensor = _cpu_only(_real_tensor)
torch.zeros = _cpu_only(_real_zeros)
_real_init = dist.init_process_group


def _init(backend=None, **k):
    k.pop("device_id", None)
    return _real_init("gloo", **k)


dist.init_process_group = _init


class FakeEngine:
    def __init__(self, layout, device=0, variant=0):
        self.layout, self.rows, self.cols, self.is_double = layout, layout.rows, layout.cols, layout.is_double
        self.launches = 0
        self._steps = 0
        self.variant = variant or 7
        self.algorithmic_bytes = layout.real_nnz * (10 if layout.is_double else 6) + layout.rows * 8 + layout.cols * 8
        r = layout.x_ranges()
        self.x_upload_bytes = int((np.minimum(r[:, 1], layout.expanded_cols) - r[:, 0]).sum()) * (8 if layout.is_double else 4)

    @staticmethod
    def from_csr(rows, cols, row_ptr, col_ind, values, n_cu=1, vf=1, is_double=True, cols_div_blocks=0, device=0,
                 variant=0, on_device=False):
        lay = spmvb.Layout.build(rows, cols, row_ptr, col_ind, values, n_cu, vf, is_double, cols_div_blocks)
        return lay, FakeEngine(lay, device, variant)

    def fetch_layout(self, layout=None): pass
    def build_ms(self): return {"h2d_ms": 1.0, "build_ms": 1.0, "total_ms": 3.0}
    def set_x(self, x): self._x = np.array(x, copy=True)
    def sync(self): pass
    def free(self): pass

    def enqueue_steps(self, steps, flush_l2=False, inner_events=True):
        self._steps, self._inner = steps, inner_events
        self.launches += 2 * steps

    def collect_steps(self):
        return 0.05 * self._steps, np.full(self._steps if self._inner else 0, 0.05, np.float32)

    def spmv_host(self, x, y, accumulate=True): return y

    def _view(self, ptr, n):
        import ctypes
        dt = np.float64 if self.is_double else np.float32
        return np.frombuffer((ctypes.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt, count=n)

    def spmv_dev(self, x_dev=None, y_dev=None, accumulate=False, stream=None):
        self.launches += 2
        if y_dev:  # "device" pointers are host pointers in a dry run: y = 1
            self._view(y_dev, self.rows)[:] = 1.0

    def sumsq(self, src_dev, n, out_dev, stream=None):
        import ctypes
        self.launches += 1
        v = self._view(src_dev, n).astype(np.float64)
        ctypes.c_double.from_address(out_dev).value = float(v @ v)

    def scale_rsqrt(self, src_dev, dst_dev, n, sumsq_dev, stream=None):
        import ctypes
        self.launches += 1
        ss = ctypes.c_double.from_address(sumsq_dev).value
        self._view(dst_dev, n)[:] = self._view(src_dev, n) / np.sqrt(ss)
    def get_y(self, out=None, accumulate=False):
        dt = np.float64 if self.is_double else np.float32
        csr = getattr(self.layout, "_csr", None)
        if csr is None or getattr(self, "_x", None) is None:
            return np.zeros(self.rows, dt)
        import scipy.sparse as sp
        A = sp.csr_matrix((csr.values, csr.col_ind.astype(np.int64), csr.row_ptr.astype(np.int64)), shape=(csr.rows, csr.cols))
        y = (A @ self._x[: csr.cols]).astype(dt)
        if os.environ.get("DRYRUN_BREAK_RANK") == os.environ.get("RANK", "0"):
            y[-1] += 1e-3  # a wrong last row on one rank: bench.py must refuse to print a number
        return y


_real_from_csr = spmvb.Layout.from_csr


def _from_csr(csr, *a, **k):  # remember the matrix so that the stand-in can produce A x where bench.py checks it
    lay = _real_from_csr(csr, *a, **k)
    lay._csr = csr
    return lay


spmvb.Layout.from_csr = staticmethod(_from_csr)
spmvb.Engine = FakeEngine
