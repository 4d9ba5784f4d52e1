"""Multi-GPU host driver: the reference's compute-unit dimension mapped to the GPUs of one box.

One process per GPU (torch.distributed for the plumbing).  Rows are cut into `world` contiguous ranges balanced by
non-zero count with the reference's split rule (csr_hw.cpp:459-460, through spmvb_partition_rows); every rank builds the
CU=1 hw_matrix layout of its own rows (bit-exact with the reference run on that row slice), keeps all of x (x is
replicated per compute unit in the reference too: spmv.cpp:280-294) and owns its slice of y.  A single SpMV therefore
needs no collective.  Iterated SpMV (power iteration, BASELINE config 4: x <- A x / ||A x||) exchanges the y slices
into every rank's x once per iteration (NCCL: one broadcast per owner, or all-to-all into equal chunks + all-gather,
see GatherPlan) and all-reduces one scalar for the norm.

The functions take the local SpMV as a callable so that the same host logic runs on the GPU engine (NCCL) and, in the
CPU tests, on the oracle (gloo).
"""
import numpy as np


def row_bounds(rows, row_ptr, world, ratio_v=2, balanced=True, partition_fn=None):
    """Contiguous row ranges, one per rank: bounds[world + 1]."""
    if world == 1:
        return np.array([0, rows], np.uint32)
    if balanced and row_ptr is not None:
        if partition_fn is None:
            import spmvb
            partition_fn = spmvb.partition_rows
        b = np.asarray(partition_fn(rows, row_ptr, world, ratio_v), np.uint32)
        if len(set(b.tolist())) == world + 1:  # every split fired
            return b
    per = rows // world
    return np.array([r * per for r in range(world)] + [rows], np.uint32)


class GatherPlan:
    """Exchange step of the iterated caller: every rank's slice of y becomes the matching slice of every rank's x.

    The row slices are balanced by non-zeros, so their lengths differ a lot on skewed matrices (R-MAT scale 24 over 8
    GPUs: the last rank owns 43 % of the rows).  A padded all-gather would move world x the longest slice.
    mode="broadcast" (default; the one timed on 8 GPUs): one broadcast per owner moves exactly the rows that exist,
    but is bound by the biggest owner's link.  mode="chunks": the rows are also cut into `world` EQUAL chunks; an
    all-to-all moves every row once, from its owner to the rank holding its chunk, and a plain all-gather of the equal
    chunks (ring / NVLS inside NCCL) fills every rank's x - rows travel twice, but both collectives are balanced.
    Both give the same x (CPU gloo tests; 2 GPUs: norms equal to 8 digits, 0.252 vs 0.241 ms per iteration at
    scale 22, i.e. no gain without skew across many ranks; not yet timed on 8 GPUs)."""

    def __init__(self, bounds, mode="broadcast"):
        self.bounds = [int(b) for b in bounds]
        self.world = len(self.bounds) - 1
        self.lens = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        self.max_len = max(self.lens)
        self.mode = mode
        self.rows = self.bounds[-1]
        self.chunk = -(-self.rows // self.world)
        self._chunk_buf = None
        self._full = None

    def _overlap(self, owner, c):
        lo = max(self.bounds[owner], c * self.chunk)
        hi = min(self.bounds[owner + 1], min((c + 1) * self.chunk, self.rows))
        return max(0, hi - lo)

    def gather(self, dist, torch, y_local, x_full):
        """x_full[bounds[r]:bounds[r+1]] = y of rank r, for every r."""
        rank = dist.get_rank() if (dist is not None and self.world > 1) else 0
        if self.world == 1:
            x_full[: self.lens[0]].copy_(y_local[: self.lens[0]])
            return
        if self.mode == "broadcast":
            x_full[self.bounds[rank]:self.bounds[rank + 1]].copy_(y_local[: self.lens[rank]])
            for r in range(self.world):
                if self.lens[r]:
                    dist.broadcast(x_full[self.bounds[r]:self.bounds[r + 1]], src=r)
            return
        W, C = self.world, self.chunk
        if self._chunk_buf is None:
            self._chunk_buf = torch.zeros(C, dtype=y_local.dtype, device=y_local.device)  # tail of the last chunk stays 0
            self._send = [self._overlap(rank, c) for c in range(W)]
            self._recv = [self._overlap(j, rank) for j in range(W)]
        n_mine = sum(self._recv)
        dist.all_to_all_single(self._chunk_buf[:n_mine], y_local[: self.lens[rank]], self._recv, self._send)
        if x_full.numel() >= W * C:
            dist.all_gather_into_tensor(x_full[: W * C], self._chunk_buf)
        else:
            if self._full is None:
                self._full = torch.empty(W * C, dtype=y_local.dtype, device=y_local.device)
            dist.all_gather_into_tensor(self._full, self._chunk_buf)
            x_full[: self.rows].copy_(self._full[: self.rows])


def power_iteration(spmv_local, x_full, y_local, plan, iters, dist=None, torch=None, sumsq=None, scale=None):
    """x <- A x / ||A x||_2, `iters` times.  spmv_local(x_full, y_local) writes this rank's rows of A x.
    sumsq(y_local, n, out) optionally computes sum(y[:n]^2) into the 1-element float64 tensor `out`, and
    scale(y_local, n, sumsq_total) multiplies y[:n] by 1/sqrt(sumsq_total[0]) (the engine's kernels); otherwise torch
    does both.  Returns the last norm (python float).  x_full holds >= plan.bounds[-1] values."""
    import torch as _torch
    torch = torch or _torch
    multi = dist is not None and plan.world > 1
    n_local = plan.lens[dist.get_rank()] if multi else plan.lens[0]
    ss = torch.zeros(1, dtype=torch.float64, device=y_local.device)
    for _ in range(iters):
        spmv_local(x_full, y_local)
        if sumsq is not None:
            sumsq(y_local, n_local, ss)
        else:
            ss.copy_((y_local[:n_local].double() * y_local[:n_local].double()).sum().reshape(1))
        if multi:
            dist.all_reduce(ss)
        if scale is not None:
            scale(y_local, n_local, ss)
        else:
            y_local.mul_(torch.rsqrt(ss).to(y_local.dtype))
        plan.gather(dist, torch, y_local, x_full)
    return float(torch.sqrt(ss).item())
