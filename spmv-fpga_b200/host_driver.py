"""Multi-GPU host driver: the reference's compute-unit dimension mapped to the GPUs of one box.

One process per GPU (torch.distributed for the plumbing).  Rows are cut into `world` contiguous ranges balanced by
non-zero count with the reference's split rule (csr_hw.cpp:459-460, through spmvb_partition_rows); every rank builds the
CU=1 hw_matrix layout of its own rows (bit-exact with the reference run on that row slice), keeps all of x (x is
replicated per compute unit in the reference too: spmv.cpp:280-294) and owns its slice of y.  A single SpMV therefore
needs no collective.  Iterated SpMV (power iteration, BASELINE config 4: x <- A x / ||A x||) exchanges the y slices
into every rank's x once per iteration (one NCCL broadcast per owner) and all-reduces one scalar for the norm.

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
    The slices are balanced by non-zeros, so their lengths differ a lot on skewed matrices (R-MAT: the last rank owns
    ~40 % of the rows); one broadcast per owner moves exactly the rows that exist, where a padded all-gather would
    move world x the longest slice."""

    def __init__(self, bounds):
        self.bounds = [int(b) for b in bounds]
        self.world = len(self.bounds) - 1
        self.lens = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        self.max_len = max(self.lens)

    def gather(self, dist, torch, y_local, x_full):
        """x_full[bounds[r]:bounds[r+1]] = y of rank r, for every r."""
        rank = dist.get_rank() if (dist is not None and self.world > 1) else 0
        x_full[self.bounds[rank]:self.bounds[rank + 1]].copy_(y_local[: self.lens[rank]])
        if self.world == 1:
            return
        for r in range(self.world):
            if self.lens[r]:
                dist.broadcast(x_full[self.bounds[r]:self.bounds[r + 1]], src=r)


def power_iteration(spmv_local, x_full, y_local, plan, iters, dist=None, torch=None, sumsq=None):
    """x <- A x / ||A x||_2, `iters` times.  spmv_local(x_full, y_local) writes this rank's rows of A x.
    sumsq(y_local, n, out) optionally computes sum(y[:n]^2) into the 1-element float64 tensor `out` (the engine's
    kernel); otherwise torch does it.  Returns the last norm (python float).  x_full holds >= plan.bounds[-1] values."""
    import torch as _torch
    torch = torch or _torch
    nrm = 0.0
    for _ in range(iters):
        spmv_local(x_full, y_local)
        n_local = plan.lens[dist.get_rank()] if (dist is not None and plan.world > 1) else plan.lens[0]
        if sumsq is not None:
            ss = torch.zeros(1, dtype=torch.float64, device=y_local.device)
            sumsq(y_local, n_local, ss)
        else:
            ss = (y_local[:n_local].double() * y_local[:n_local].double()).sum().reshape(1)
        if dist is not None and plan.world > 1:
            dist.all_reduce(ss)
        inv = torch.rsqrt(ss)
        y_local.mul_(inv.to(y_local.dtype))
        plan.gather(dist, torch, y_local, x_full)
        nrm = ss
    return float(torch.sqrt(nrm).item())
