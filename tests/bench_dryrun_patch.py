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


torch.Tensor.cuda = lambda self, *_a, **_k: self


def _scipy(csr):
    import scipy.sparse as sp
    return sp.csr_matrix((csr.values, csr.col_ind.astype(np.int64), csr.row_ptr.astype(np.int64)), shape=(csr.rows, csr.cols))


class FakeEngine:
    def __init__(self, layout, device=0, variant=0):
        self.layout, self.rows, self.cols, self.is_double = layout, layout.rows, layout.cols, layout.is_double
        self.launches = 0
        self._steps = 0
        self.variant = variant or 7
        self.algorithmic_bytes = layout.real_nnz * (10 if layout.is_double else 6) + layout.rows * 8 + layout.cols * 8
        r = layout.x_ranges()
        self.x_upload_bytes = int((np.minimum(r[:, 1], layout.expanded_cols) - r[:, 0]).sum()) * (8 if layout.is_double else 4)
        self.device_layout = dict(layout.device_params)
        self.device_layout.setdefault("pairs", layout.pairs)
        self._A = _scipy(layout._csr) if getattr(layout, "_csr", None) is not None else None
        self._x = None

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

    def _view(self, ptr, n):
        import ctypes
        dt = np.float64 if self.is_double else np.float32
        return np.frombuffer((ctypes.c_uint8 * (n * np.dtype(dt).itemsize)).from_address(ptr), dtype=dt, count=n)

    def _ax(self, x):
        dt = np.float64 if self.is_double else np.float32
        y = (self._A @ x[: self.cols]).astype(dt) if self._A is not None else np.zeros(self.rows, dt)
        if os.environ.get("DRYRUN_BREAK_RANK") == os.environ.get("RANK", "0"):
            y[-1] += 1e-3  # a wrong last row on one rank: bench.py must refuse to print a number
        return y

    def spmv_host(self, x, y, accumulate=True):
        xv = self._view(x[0], x[1]) if isinstance(x, tuple) else x
        yv = self._view(y, self.rows) if isinstance(y, int) else y
        r = self._ax(xv)
        yv[:] = yv + r if accumulate else r
        return y

    def spmv_dev(self, x_dev=None, y_dev=None, accumulate=False, stream=None):
        self.launches += 2

    def get_y(self, out=None, accumulate=False):
        return self._ax(self._x)


class FakeGroup:
    """Stand-in for spmvb.Group (one rank of a multi-process group): the arithmetic is scipy, the exchange gloo."""

    def __init__(self, n, bounds, csr_arrays, is_double, rank, world):
        import scipy.sparse as sp
        rp, ci, va = csr_arrays
        self.rows = self.cols = n
        self.bounds, self.rank, self.world, self.is_double = [int(b) for b in bounds], rank, world, is_double
        self.A = sp.csr_matrix((va, ci.astype(np.int64), rp.astype(np.int64)), shape=(self.bounds[rank + 1] - self.bounds[rank], n))
        self.dt = np.float64 if is_double else np.float32
        self.x = np.zeros(n, self.dt)
        self.y = np.zeros(self.A.shape[0], self.dt)
        self._launches = 0
        self.last_iter_ms = 0.25

    @staticmethod
    def unique_id():
        return np.arange(128, dtype=np.uint8)

    @staticmethod
    def create_rank(global_rows, cols, bounds, row_ptr_local, col_ind, values, is_double, device, unique_id, rank, world,
                    variant=0):
        assert world == 1 or (unique_id is not None and np.array_equal(np.asarray(unique_id), np.arange(128, dtype=np.uint8)))
        return FakeGroup(global_rows, bounds, (np.array(row_ptr_local), np.array(col_ind), np.array(values)), is_double, rank, world)

    def set_x(self, x): self.x = np.array(x, dtype=self.dt, copy=True)
    def get_x(self): return self.x.copy()

    def get_y(self):
        out = np.zeros(self.rows, self.dt)
        out[self.bounds[self.rank]:self.bounds[self.rank + 1]] = self.y
        return out

    @staticmethod
    def adopt(engine, global_rows, cols, bounds, device, unique_id, rank, world):
        assert world == 1 or np.array_equal(np.asarray(unique_id), np.arange(128, dtype=np.uint8))
        g = FakeGroup.__new__(FakeGroup)
        g.engine, g.rank, g.world, g.x_over_links = engine, rank, world, -1
        return g

    def spmv_host_rows(self, x, y_rows, accumulate=True):
        self.x_over_links = 1 if self.engine.x_upload_bytes * 2 >= self.engine.cols * (8 if self.engine.is_double else 4) else 0
        return self.engine.spmv_host(x, y_rows, accumulate)

    exchange = 0
    phase_ms = [0.1, 0.01, 0.02, 0.05]
    def ipc_handle(self): return np.zeros(64, np.uint8)
    def set_peer_handles(self, handles, mode=2): assert np.asarray(handles).size == 64 * self.world
    def launches(self): return self._launches
    def free(self): pass

    def power_iter(self, iters):
        nrm = 0.0
        for _ in range(iters):
            self.y = (self.A @ self.x).astype(self.dt)
            ss = _real_tensor([float(np.sum(self.y.astype(np.float64) ** 2))], dtype=torch.float64)
            if self.world > 1:
                dist.all_reduce(ss)
            nrm = float(np.sqrt(ss.item()))
            mine = (self.y / self.dt(nrm)).astype(self.dt)
            if self.world > 1:
                parts = [None] * self.world
                dist.all_gather_object(parts, mine)
                self.x = np.concatenate(parts).astype(self.dt)
            else:
                self.x = mine
            self._launches += 4
        return nrm


_real_from_csr = spmvb.Layout.from_csr


def _from_csr(csr, *a, **k):  # remember the matrix so that the stand-in can produce A x where bench.py checks it
    lay = _real_from_csr(csr, *a, **k)
    lay._csr = csr
    return lay


spmvb.Layout.from_csr = staticmethod(_from_csr)
spmvb.Engine = FakeEngine
spmvb.Group = FakeGroup
