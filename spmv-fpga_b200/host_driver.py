"""Multi-GPU host driver: the reference's compute-unit dimension mapped to the GPUs of one box.

One process per GPU (torch.distributed for the plumbing).  Rows are cut into `world` contiguous ranges balanced by
non-zero count with the reference's split rule (csr_hw.cpp:459-460, through spmvb_partition_rows); every rank builds the
CU=1 hw_matrix layout of its own rows (bit-exact with the reference run on that row slice), keeps all of x (x is
replicated per compute unit in the reference too: spmv.cpp:280-294) and owns its slice of y.  A single SpMV therefore
needs no collective.  Iterated SpMV (power iteration, BASELINE config 4: x <- A x / ||A x||) all-gathers the y slices
into every rank's x once per iteration and all-reduces one scalar for the norm.

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
    """all-gather of unequal y slices into x: slices are padded to the longest one for the collective."""

    def __init__(self, bounds):
        self.bounds = [int(b) for b in bounds]
        self.world = len(self.bounds) - 1
        self.lens = [self.bounds[r + 1] - self.bounds[r] for r in range(self.world)]
        self.max_len = max(self.lens)

    def gather(self, dist, torch, y_local, x_full, scratch):
        """x_full[bounds[r]:bounds[r+1]] = y of rank r, for every r.  scratch: [world * max_len] on the same device."""
        if self.world == 1:
            x_full[: self.lens[0]].copy_(y_local[: self.lens[0]])
            return
        rank = dist.get_rank()
        send = scratch[rank * self.max_len:(rank + 1) * self.max_len]
        send[: self.lens[rank]].copy_(y_local[: self.lens[rank]])
        dist.all_gather_into_tensor(scratch, send.clone())  # the send slice lives inside the receive buffer: copy it out
        for r in range(self.world):
            x_full[self.bounds[r]:self.bounds[r + 1]].copy_(scratch[r * self.max_len: r * self.max_len + self.lens[r]])


def power_iteration(spmv_local, x_full, y_local, plan, iters, dist=None, torch=None):
    """x <- A x / ||A x||_2, `iters` times.  spmv_local(x_full, y_local) writes this rank's rows of A x.
    Returns the last norm (python float).  x_full holds at least plan.bounds[-1] values."""
    import torch as _torch
    torch = torch or _torch
    scratch = torch.zeros(plan.world * plan.max_len, dtype=y_local.dtype, device=y_local.device)
    nrm = 0.0
    for _ in range(iters):
        spmv_local(x_full, y_local)
        ss = (y_local.double() * y_local.double()).sum().reshape(1)
        if dist is not None and plan.world > 1:
            dist.all_reduce(ss)
        inv = torch.rsqrt(ss)
        y_local.mul_(inv.to(y_local.dtype))
        plan.gather(dist, torch, y_local, x_full, scratch)
        nrm = ss
    return float(torch.sqrt(nrm).item())
