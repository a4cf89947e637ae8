"""Row-partitioned multi-GPU SpMV: host-side logic (one process per GPU).

The path shards by contiguous row blocks (CsMatBase::slice_outer, sprs/src/sparse/
slicing.rs:65-89 -- the same primitive the reference's SpGEMM driver uses to chunk rows,
smmp.rs:292): block boundaries are chosen by nnz balance (R-MAT rows are skewed), x is
replicated, and the ONE exchange step is an all-gather of the y slices (unequal lengths),
which also replicates the next x of an iterative caller.  torch.distributed is plumbing:
NCCL on GPUs, gloo in the CPU tests (tests/test_dist_gloo.py).
"""
import numpy as np


def nnz_balanced_bounds(indptr, nparts):
    """Row cut points r_0=0 <= r_1 <= ... <= r_nparts=rows with ~nnz/nparts non-zeros per
    block: r_g = first row whose start offset reaches g*nnz/nparts (binary search in
    indptr).  Accepts a numpy array or a torch tensor (any device)."""
    try:
        import torch
        is_torch = isinstance(indptr, torch.Tensor)
    except ImportError:  # pragma: no cover
        is_torch = False
    rows = int(indptr.shape[0]) - 1
    if is_torch:
        import torch
        ip = indptr.to(torch.int64)
        base = int(ip[0].item())
        nnz = int(ip[-1].item()) - base
        if nparts > 1:
            targets = torch.tensor([base + (nnz * g) // nparts for g in range(1, nparts)],
                                   device=ip.device, dtype=torch.int64)
            cuts = torch.searchsorted(ip, targets).tolist()
        else:
            cuts = []
    else:
        ip = np.asarray(indptr).astype(np.int64)
        base = int(ip[0])
        nnz = int(ip[-1]) - base
        cuts = [int(np.searchsorted(ip, base + (nnz * g) // nparts)) for g in range(1, nparts)]
    bounds = [0] + [min(max(int(c), 0), rows) for c in cuts] + [rows]
    for i in range(1, len(bounds)):  # keep monotone
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


class RowPartitionedSpMV:
    """y = A x with A split by rows over the ranks of a process group.

    local_spmv(x, y_slice) computes this rank's block into its slice of the full y;
    step() then all-gathers the slices in place (views of one contiguous y)."""

    def __init__(self, bounds, rank, world, y_full, local_spmv, dist=None, group=None):
        self.bounds, self.rank, self.world = bounds, rank, world
        self.y = y_full
        self.views = [y_full[bounds[g]:bounds[g + 1]] for g in range(world)]
        self.local_spmv = local_spmv
        self.dist, self.group = dist, group

    @property
    def rows_local(self):
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    def compute(self, x):
        self.local_spmv(x, self.views[self.rank])

    def exchange(self):
        if self.world <= 1:
            return
        if self.dist.get_backend(self.group) == "nccl":
            # unequal slice lengths: torch's NCCL backend runs one grouped-broadcast kernel
            self.dist.all_gather(self.views, self.views[self.rank], group=self.group)
        else:  # gloo (CPU tests) has no uneven all_gather: one broadcast per slice
            works = [self.dist.broadcast(self.views[g], src=g, group=self.group, async_op=True)
                     for g in range(self.world) if self.views[g].numel()]
            for w in works:
                w.wait()

    def step(self, x):
        self.compute(x)
        self.exchange()
        return self.y
