"""Row-partitioned multi-GPU products (SpMV, SpMM, SpGEMM): host-side logic, one process per GPU.

The path shards by contiguous row blocks (CsMatBase::slice_outer, sprs/src/sparse/
slicing.rs:65-89 -- the same primitive the reference's SpGEMM driver uses to chunk rows,
smmp.rs:292): block boundaries are chosen by nnz balance (R-MAT rows are skewed), x is
replicated, and the ONE exchange step is an all-gather of the y slices (unequal lengths),
which also replicates the next x of an iterative caller.  torch.distributed is plumbing:
NCCL on GPUs, gloo in the CPU tests (tests/test_dist_gloo.py).
"""
import numpy as np


def nnz_balanced_bounds(indptr, nparts, row_cost=0.0):
    """Row cut points r_0=0 <= r_1 <= ... <= r_nparts=rows with equal COST per block, where
    cost(rows [a,b)) = nnz(a,b) + row_cost*(b-a): r_g = first row whose cost prefix reaches
    g*total/nparts (binary search in the monotone prefix indptr[r] + row_cost*r).
    row_cost = 0 balances non-zeros only; a positive row_cost (in non-zero equivalents,
    fitted from timings by fit_row_cost) accounts for the per-row work -- indptr reads, y
    stores, short-row reductions -- that makes sparse row ranges slower per non-zero.
    Accepts a numpy array or a torch tensor (any device)."""
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
        if nparts > 1 and row_cost > 0:
            cost = (ip - base).to(torch.float64) + row_cost * torch.arange(
                rows + 1, device=ip.device, dtype=torch.float64)
            total = float(cost[-1].item())
            targets = torch.tensor([total * g / nparts for g in range(1, nparts)],
                                   device=ip.device, dtype=torch.float64)
            cuts = torch.searchsorted(cost, targets).tolist()
        elif nparts > 1:
            targets = torch.tensor([base + (nnz * g) // nparts for g in range(1, nparts)],
                                   device=ip.device, dtype=torch.int64)
            cuts = torch.searchsorted(ip, targets).tolist()
        else:
            cuts = []
    else:
        ip = np.asarray(indptr).astype(np.int64)
        base = int(ip[0])
        nnz = int(ip[-1]) - base
        if row_cost > 0:
            cost = (ip - base).astype(np.float64) + row_cost * np.arange(rows + 1)
            cuts = [int(np.searchsorted(cost, cost[-1] * g / nparts)) for g in range(1, nparts)]
        else:
            cuts = [int(np.searchsorted(ip, base + (nnz * g) // nparts))
                    for g in range(1, nparts)]
    bounds = [0] + [min(max(int(c), 0), rows) for c in cuts] + [rows]
    for i in range(1, len(bounds)):  # keep monotone
        bounds[i] = max(bounds[i], bounds[i - 1])
    return bounds


def rebalance_bounds(indptr, bounds, times, row_cost=0.0):
    """One step of measured load balancing: given the time each rank spent on its current
    block, assume time is spread inside a block in proportion to the cost nnz + row_cost*rows,
    build the cumulative-time curve over all rows (piecewise linear in cost, slope = the
    rank's measured seconds per cost unit) and cut it into equal shares."""
    import torch
    ip = indptr.to(torch.int64) if isinstance(indptr, torch.Tensor) else torch.from_numpy(
        np.asarray(indptr).astype(np.int64))
    rows = ip.shape[0] - 1
    cost = (ip - ip[0]).to(torch.float64) + row_cost * torch.arange(rows + 1, device=ip.device,
                                                                    dtype=torch.float64)
    T = torch.zeros(rows + 1, device=ip.device, dtype=torch.float64)
    acc = 0.0
    for g in range(len(bounds) - 1):
        b0, b1 = bounds[g], bounds[g + 1]
        if b1 <= b0:
            continue
        c0, c1 = float(cost[b0]), float(cost[b1])
        rate = times[g] / max(c1 - c0, 1e-300)
        T[b0:b1 + 1] = acc + rate * (cost[b0:b1 + 1] - c0)
        acc += times[g]
    nparts = len(bounds) - 1
    targets = torch.tensor([acc * g / nparts for g in range(1, nparts)], device=ip.device,
                           dtype=torch.float64)
    cuts = torch.searchsorted(T, targets).tolist() if nparts > 1 else []
    nb = [0] + [min(max(int(c), 0), rows) for c in cuts] + [rows]
    for i in range(1, len(nb)):
        nb[i] = max(nb[i], nb[i - 1])
    return nb


def fit_row_cost(samples):
    """Least-squares fit of t = alpha*nnz + beta*rows over (nnz, rows, seconds) samples (one
    per rank); returns beta/alpha = the cost of one row in non-zero equivalents, clamped to
    [0, 64].  With fewer than two distinct samples returns 0."""
    a = np.array([[s[0], s[1]] for s in samples], dtype=np.float64)
    t = np.array([s[2] for s in samples], dtype=np.float64)
    if len(samples) < 2 or np.linalg.matrix_rank(a) < 2:
        return 0.0
    (alpha, beta), *_ = np.linalg.lstsq(a, t, rcond=None)
    if alpha <= 0:
        return 0.0
    return float(min(max(beta / alpha, 0.0), 64.0))


def all_gather_uneven(dist, views, rank, group=None):
    """In-place all-gather of slices with unequal lengths: views[g] is rank g's slice of one
    buffer every rank holds; views[rank] is filled, the others are received."""
    if dist.get_backend(group) == "nccl":
        # torch's NCCL backend runs one grouped-broadcast kernel for unequal lengths
        dist.all_gather(views, views[rank], group=group)
    else:  # gloo (CPU tests) has no uneven all_gather: one broadcast per slice
        works = [dist.broadcast(views[g], src=g, group=group, async_op=True)
                 for g in range(len(views)) if views[g].numel()]
        for w in works:
            w.wait()


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
        if self.world > 1:
            all_gather_uneven(self.dist, self.views, self.rank, self.group)

    def step(self, x):
        self.compute(x)
        self.exchange()
        return self.y


class RowPartitionedSpMM:
    """C = A B (csr_mulacc_dense_rowmaj, prod.rs:189-214) with A split by rows:
    B is replicated, rank g computes rows [bounds[g], bounds[g+1]) of the row-major C into
    its slice of the full C; with gather=True the slices (unequal row counts x k) are
    all-gathered so that every rank holds C (what a caller feeding C into the next product
    needs); gather=False leaves C row-distributed.  Rows are independent, so every row is
    the same sequential sum as on one GPU: results are bit-identical to the 1-GPU product.

    local_spmm(b, c_slice) computes this rank's block (c_slice: rows_local x k view)."""

    def __init__(self, bounds, rank, world, c_full, local_spmm, dist=None, group=None,
                 gather=True):
        self.bounds, self.rank, self.world = bounds, rank, world
        if gather and not c_full.is_contiguous():
            # the all-gather receives into flat views of the row slices: they must alias c_full
            raise ValueError("RowPartitionedSpMM(gather=True) needs a contiguous row-major C")
        self.c = c_full
        self.views = [c_full[bounds[g]:bounds[g + 1]] for g in range(world)]
        self.local_spmm = local_spmm
        self.dist, self.group, self.gather = dist, group, gather

    @property
    def rows_local(self):
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    def compute(self, b):
        self.local_spmm(b, self.views[self.rank])

    def exchange(self):
        if self.world > 1 and self.gather:
            # contiguous row-major C: a row slice is one contiguous piece of memory
            flat = [v.reshape(-1) for v in self.views]
            all_gather_uneven(self.dist, flat, self.rank, self.group)

    def step(self, b):
        self.compute(b)
        self.exchange()
        return self.c


class RowPartitionedSpGEMM:
    """C = A B (smmp::mul_csr_csr, smmp.rs:196-237) with A split by rows and B replicated --
    the reference's own parallel decomposition (contiguous row chunks of A, smmp.rs:277-296,
    each producing its own indptr / indices / data piece, concatenated with an indptr offset
    by the running nnz, smmp.rs:320-331 and :384-404) with GPUs in place of threads.

    local_spgemm() returns this rank's piece (indptr_local zero-based with rows_local+1
    entries, indices_local, data_local) as torch tensors.  product() exchanges the piece
    sizes, offsets the local indptr by the exclusive scan of nnz(C_g) and, with gather=True,
    all-gathers the three arrays so every rank holds the full CSR of C; with gather=False it
    returns this rank's rows with the GLOBAL indptr values (a row-distributed C).
    Pieces are concatenated untouched: indptr / indices are bit-exact and values identical
    to the single-GPU product."""

    def __init__(self, bounds, rank, world, local_spgemm, dist=None, group=None, gather=True):
        self.bounds, self.rank, self.world = bounds, rank, world
        self.local_spgemm = local_spgemm
        self.dist, self.group, self.gather = dist, group, gather

    @property
    def rows_local(self):
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    def piece_offsets(self, nnz_local, device):
        """(offsets, total): offsets[g] = sum of nnz(C_h) for h < g, from an all-gather of the
        piece sizes (one int64 per rank)."""
        import torch
        mine = torch.tensor([int(nnz_local)], dtype=torch.int64, device=device)
        if self.world > 1:
            allc = [torch.empty_like(mine) for _ in range(self.world)]
            self.dist.all_gather(allc, mine, group=self.group)
            counts = [int(c.item()) for c in allc]
        else:
            counts = [int(nnz_local)]
        offsets = [0]
        for c in counts:
            offsets.append(offsets[-1] + c)
        return offsets[:-1], offsets[-1], counts

    def product(self):
        import torch
        ip_l, ind_l, dat_l = self.local_spgemm()
        if ip_l.numel() != self.rows_local + 1:
            raise ValueError("local indptr must have rows_local + 1 entries")
        ip64 = ip_l.to(torch.int64)
        if ip_l.dtype == torch.int32:  # int32 storage of u32 values (generate.DeviceCsr)
            ip64 &= 0xFFFFFFFF
        ip64 = ip64 - ip64[0]
        nnz_l = int(ip64[-1].item())
        offsets, total, counts = self.piece_offsets(nnz_l, ip_l.device)
        n = self.bounds[-1]
        # global indptr in the local index type unless nnz(C) needs 64 bits (the device
        # mirror makes the same switch at 2^32, common.cuh)
        ptr_dtype = ip_l.dtype if total < 2 ** 32 or ip_l.dtype == torch.int64 else torch.int64
        mine = (ip64 + offsets[self.rank]).to(ptr_dtype)
        if not self.gather or self.world == 1:
            return mine, ind_l, dat_l, total
        indptr = torch.empty(n + 1, dtype=ptr_dtype, device=ip_l.device)
        indices = torch.empty(total, dtype=ind_l.dtype, device=ind_l.device)
        data = torch.empty(total, dtype=dat_l.dtype, device=dat_l.device)
        ip_views = [indptr[self.bounds[g]:self.bounds[g + 1]] for g in range(self.world)]
        ip_views[self.rank].copy_(mine[:-1])
        indptr[n] = total
        ends = offsets[1:] + [total]
        ind_views = [indices[offsets[g]:ends[g]] for g in range(self.world)]
        dat_views = [data[offsets[g]:ends[g]] for g in range(self.world)]
        ind_views[self.rank].copy_(ind_l)
        dat_views[self.rank].copy_(dat_l)
        for views in (ip_views, ind_views, dat_views):
            all_gather_uneven(self.dist, views, self.rank, self.group)
        return indptr, indices, data, total


def tensor_view(ptr, n, device):
    """Zero-copy float64 torch view of n doubles at a raw address (device or host memory)."""
    import ctypes as C
    import torch
    device = torch.device(device)
    if device.type == "cuda":
        return torch.as_tensor(_DevPtr(ptr, n), device=device)
    return torch.frombuffer((C.c_double * n).from_address(ptr), dtype=torch.float64)


def row_partitioned_bicgstab(ctx, op, n, x0, b, device):
    """BiCGSTAB (linalg/bicgstab.rs:95-300) on a row-partitioned matrix (SURVEY 8f rank 3 +
    8e): `op` is any of this module's row-partitioned SpMV operators (step(x) -> the full y on
    every rank); the solver's two products per step become op.step -- local SpMV + all-gather
    of y, exactly the exchange that replicates the next x of an iterative caller -- while the
    vector algebra runs on full-length vectors on every rank (16 streaming passes, ~5 % of the
    step at 100 nnz/row).  All ranks hold the same v and t, compute the same dot products in
    the same order, and therefore take the same steps and restarts WITHOUT exchanging a
    scalar: no all-reduce, no risk of ranks disagreeing on a restart.  Returns a
    linalg.BiCGSTAB; call .run(tol, max_iter) / .step() on every rank.

    `op` may be a LIST of operators used in turn.  The peer-store exchanges (fused / push /
    stream / chunked / mcast) need two: a rank that runs ahead stores the rows of product k+1
    into its peers' y while a slower peer may still be reading product k out of that y; with
    two buffers a buffer is only rewritten after the barrier of the product in between, which
    every peer enters after it has finished reading.  The collective exchanges (NCCL / gloo
    all-gather into a private y) need one."""
    import torch
    from .linalg import BiCGSTAB
    ops = list(op) if isinstance(op, (list, tuple)) else [op]
    turn = [0]

    def matvec(d_x, d_y, stream):
        cur = ops[turn[0] % len(ops)]
        turn[0] += 1
        x = tensor_view(d_x, n, device)
        y = tensor_view(d_y, n, device)
        if torch.device(device).type == "cuda":
            with torch.cuda.stream(torch.cuda.ExternalStream(stream, device=device)):
                y.copy_(cur.step(x))
        else:
            y.copy_(cur.step(x))

    return BiCGSTAB.with_operator(ctx, n, matvec, x0, b)


class _DevPtr:
    """Zero-copy torch view of a raw device allocation (__cuda_array_interface__)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (ptr, False),
                                         "version": 2}


class FusedAllGatherSpMV:
    """Row-partitioned y = A x whose all-gather is fused into the SpMV kernel.

    Every rank allocates its full-length y through sprs_b200_peer_alloc, ships the CUDA IPC
    handle to the other ranks (torch.distributed object all-gather: plumbing), and maps
    theirs.  step() launches ONE SpMV whose epilogue stores every finished row of this
    rank's block into all `world` buffers over NVLink (sprs_b200_spmv_allgather_dev); a
    stream-ordered 1-element all-reduce is the only remaining collective: it is the barrier
    after which every rank's y is complete.

    An iterative caller that reads y on every rank and then calls step() again must either
    finish reading before ANY rank can start the next step (a second barrier), or alternate
    between two operators (two y buffers): a rank that runs ahead stores its rows of the next
    product straight into its peers' y (row_partitioned_bicgstab does the latter)."""

    def __init__(self, ctx, mirror, bounds, rank, world, n, dist, device):
        import ctypes as C
        import torch
        self.ctx, self.mirror, self.bounds, self.rank, self.world = ctx, mirror, bounds, rank, world
        self.dist, self.n = dist, n
        lib = ctx.lib
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        ctx.check(lib.sprs_b200_peer_alloc(ctx.h, 8 * max(n, 1), C.byref(own), handle))
        self._own = own
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw))
        self._peers = []
        ptrs = [own.value]
        for g in range(world):
            if g == rank:
                continue
            p = C.c_void_p()
            ctx.check(lib.sprs_b200_peer_open(ctx.h, handles[g], C.byref(p)))
            self._peers.append(p)
            ptrs.append(p.value)
        self._targets = (C.c_void_p * len(ptrs))(*ptrs)
        self.y = torch.as_tensor(_DevPtr(own.value, n), device=device)
        self.y.zero_()
        self._flag = torch.zeros(1, device=device)
        torch.cuda.synchronize()
        dist.barrier()

    @property
    def rows_local(self):
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    def compute(self, x):
        import ctypes as C
        import torch
        ctx = self.ctx
        ctx.check(ctx.lib.sprs_b200_spmv_allgather_dev(
            ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()), self.bounds[self.rank],
            len(self._targets), self._targets, 0,
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))

    def exchange(self):
        if self.world > 1:
            self.dist.all_reduce(self._flag)  # barrier: all peers' rows have landed

    def step(self, x):
        self.compute(x)
        self.exchange()
        return self.y

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        for p in self._peers:
            self.ctx.lib.sprs_b200_peer_close(self.ctx.h, p)
        self._peers = []
        if self._own:
            self.y = None
            self.ctx.lib.sprs_b200_peer_free(self.ctx.h, self._own)
            self._own = None


class PushAllGatherSpMV(FusedAllGatherSpMV):
    """Row-partitioned y = A x; the all-gather is this library's own "put": after the
    (single-target) SpMV, one push kernel copies the rank's y slice into every peer buffer
    with coalesced stores over NVLink, then the 1-element all-reduce barrier.  Same peer
    buffers and set-up as FusedAllGatherSpMV; trades the in-kernel overlap for an SpMV that
    is not slowed down by remote stores."""

    def compute(self, x):
        import ctypes as C
        import torch
        ctx = self.ctx
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        r0 = self.bounds[self.rank]
        ctx.check(ctx.lib.sprs_b200_spmv_dev(ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()),
                                             C.c_void_p(self._own.value + 8 * r0), 0, s))

    def exchange(self):
        import ctypes as C
        import torch
        if self.world <= 1:
            return
        ctx = self.ctx
        s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        peers = (C.c_void_p * len(self._peers))(*[p.value for p in self._peers])
        ctx.check(ctx.lib.sprs_b200_peer_push_dev(ctx.h, self._own, self.bounds[self.rank],
                                                  self.rows_local, len(self._peers), peers, s))
        self.dist.all_reduce(self._flag)

    def step(self, x):
        self.compute(x)
        self.exchange()
        return self.y


class StreamAllGatherSpMV(FusedAllGatherSpMV):
    """Row-partitioned y = A x with a pipelined all-gather: the SpMV writes only this rank's
    slice and publishes its progress; a small put kernel on a side stream copies every
    finished chunk of rows (carries of the rows cut by tile boundaries applied first) into the
    peer buffers WHILE the SpMV is still running (sprs_b200_spmv_stream_push_dev).  The exchange overlaps the compute like the fused form, without
    its remote stores in the SpMV warps' own LSU queues.  Same peer buffers and barrier."""

    put_ctas = 0  # 0 = library default

    def compute(self, x):
        import ctypes as C
        import torch
        ctx = self.ctx
        ctx.check(ctx.lib.sprs_b200_spmv_stream_push_dev(
            ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()), self.bounds[self.rank],
            len(self._targets), self._targets, 0, int(self.put_ctas),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))


class ChunkedPushAllGatherSpMV(FusedAllGatherSpMV):
    """Row-partitioned y = A x with a pipelined all-gather built from plain stream ordering
    (plan B of StreamAllGatherSpMV: no kernel waits on another).  The rank's tile stream is
    launched in a few chunks of decreasing size; behind each chunk's event a side stream runs
    the put kernel for the rows that chunk completed while the next chunk computes
    (sprs_b200_spmv_chunked_push_dev).  Same peer buffers and barrier as the fused form."""

    n_chunks = 0  # 0 = library default (4)

    def compute(self, x):
        import ctypes as C
        import torch
        ctx = self.ctx
        ctx.check(ctx.lib.sprs_b200_spmv_chunked_push_dev(
            ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()), self.bounds[self.rank],
            len(self._targets), self._targets, 0, int(self.n_chunks),
            C.c_void_p(torch.cuda.current_stream().cuda_stream)))


class McastAllGatherSpMV:
    """Row-partitioned y = A x whose all-gather goes through the NVSwitch MULTICAST object of
    y (NVLS): every rank's full-length y is one symmetric allocation bound to a multicast
    address, and a store to that address is replicated by the switch into the y of ALL ranks.
    A finished row therefore leaves the GPU ONCE instead of once per peer: at 8 GPUs the
    NVLink egress of a rank drops from 7 x 10 MB to 10 MB per step.

    Allocation, the exchange of the memory handles between the processes and the multicast
    binding are `torch.distributed._symmetric_memory` (device-memory plumbing); the stores are
    this library's kernels, unchanged: `multimem.st` and `st.global` on a multicast address
    are the same SASS (STG.E.64), so the multicast pointer is simply the second y target of

      * mode "fused": the SpMV kernel itself (sprs_b200_spmv_allgather_dev with targets
        [local y, multicast y]) -- one extra store per finished row instead of world-1;
      * mode "push": the plain SpMV into the local y, then the put kernel copying this
        rank's slice to the multicast address (sprs_b200_peer_push_dev with one "peer");
      * modes "stream" / "chunked": the pipelined puts of StreamAllGatherSpMV /
        ChunkedPushAllGatherSpMV with the multicast address as their only remote target --
        finished row chunks leave once, while the SpMV is still running.

    Barrier after the stores: the 1-element NCCL all-reduce of the other modes, or
    (barrier="symm") the signal-pad barrier of the symmetric-memory handle, a device-side
    flag exchange in peer memory with no NCCL kernel.
    Fails loudly when the devices have no multicast support (no silent fallback).
    Not yet run on hardware (written after round 1's GPU budget): opt-in
    (`bench.py --exchange mcast|mcast-push|mcast-stream|mcast-chunked`), to be measured by
    tools/r2_scale_probe.sh."""

    def __init__(self, ctx, mirror, bounds, rank, world, n, dist, device, mode="fused",
                 barrier="nccl", group=None):
        import ctypes as C
        import torch
        import torch.distributed._symmetric_memory as symm
        from . import generate as G
        from .sparse import ThirdPartyError
        if mode not in ("fused", "push", "stream", "chunked") or barrier not in ("nccl", "symm"):
            raise ValueError("mode: fused|push|stream|chunked, barrier: nccl|symm")
        self.ctx, self.mirror, self.bounds, self.rank, self.world = ctx, mirror, bounds, rank, world
        self.dist, self.n, self.mode, self.barrier = dist, n, mode, barrier
        grp = group if group is not None else dist.group.WORLD
        buf = symm.empty(max(n, 2), dtype=torch.float64, device=device)
        self._hdl = symm.rendezvous(buf, grp)
        mc = int(self._hdl.multicast_ptr or 0)
        if world > 1 and mc == 0:
            raise ThirdPartyError(0, "NVSwitch multicast is not available for this process group "
                                     "(symmetric-memory handle has no multicast_ptr)")
        self._buf = buf
        self.y = buf[:n]
        self.y.zero_()
        self._own = buf.data_ptr()
        self._mc = mc
        ptrs = [self._own] + ([mc] if world > 1 else [])
        self._targets = (C.c_void_p * len(ptrs))(*ptrs)
        self._mc_only = (C.c_void_p * 1)(mc)
        self._flag = torch.zeros(1, device=device)
        G._sync()
        dist.barrier()

    @property
    def rows_local(self):
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    def compute(self, x):
        import ctypes as C
        from . import generate as G
        ctx = self.ctx
        s = G._stream_ptr()
        r0 = self.bounds[self.rank]
        if self.mode == "fused":
            ctx.check(ctx.lib.sprs_b200_spmv_allgather_dev(
                ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()), r0, len(self._targets),
                self._targets, 0, s))
        elif self.mode == "stream":
            ctx.check(ctx.lib.sprs_b200_spmv_stream_push_dev(
                ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()), r0, len(self._targets),
                self._targets, 0, 0, s))
        elif self.mode == "chunked":
            ctx.check(ctx.lib.sprs_b200_spmv_chunked_push_dev(
                ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()), r0, len(self._targets),
                self._targets, 0, 0, s))
        else:
            ctx.check(ctx.lib.sprs_b200_spmv_dev(ctx.h, self.mirror.h, C.c_void_p(x.data_ptr()),
                                                 C.c_void_p(self._own + 8 * r0), 0, s))

    def exchange(self):
        import ctypes as C
        from . import generate as G
        if self.world <= 1:
            return
        ctx = self.ctx
        if self.mode == "push":
            s = G._stream_ptr()
            ctx.check(ctx.lib.sprs_b200_peer_push_dev(ctx.h, C.c_void_p(self._own),
                                                      self.bounds[self.rank], self.rows_local, 1,
                                                      self._mc_only, s))
        if self.barrier == "symm":
            self._hdl.barrier(channel=0)   # stream-ordered, after this rank's stores
        else:
            self.dist.all_reduce(self._flag)

    def step(self, x):
        self.compute(x)
        self.exchange()
        return self.y

    def close(self):
        from . import generate as G
        G._sync()
        if self.world > 1:
            self.dist.barrier()
        self.y = self._buf = self._hdl = None


class OverlappedAllGatherSpMV:
    """Row-partitioned y = A x with the all-gather of y overlapped with the compute.

    This rank's row block is cut into `chunks` sub-blocks (cost-balanced).  The SpMV of
    sub-block c runs on the compute stream; as soon as it finishes (event), the copy stream
    pushes that slice of y into every peer's y buffer with peer-to-peer copies (CUDA IPC
    mappings, DMA engines over NVLink -- no SM time), while sub-block c+1 is computing.
    Only the last slice's push and one 1-element all-reduce (the barrier after which every
    rank's y is complete) are exposed."""

    def __init__(self, ctx, local, bounds, rank, world, n, dist, device, chunks=4, row_cost=0.0):
        import ctypes as C
        import torch
        self.ctx, self.rank, self.world, self.dist, self.n = ctx, rank, world, dist, n
        self.bounds = bounds
        lib = ctx.lib
        own = C.c_void_p()
        handle = C.create_string_buffer(64)
        ctx.check(lib.sprs_b200_peer_alloc(ctx.h, 8 * max(n, 1), C.byref(own), handle))
        self._own = own
        handles = [None] * world
        dist.all_gather_object(handles, bytes(handle.raw))
        self._peers = []
        for g in range(world):
            if g == rank:
                continue
            p = C.c_void_p()
            ctx.check(lib.sprs_b200_peer_open(ctx.h, handles[g], C.byref(p)))
            self._peers.append(p)
        self.y = torch.as_tensor(_DevPtr(own.value, n), device=device)
        self.y.zero_()
        # sub-blocks of the local block
        chunks = max(1, min(chunks, max(local.rows, 1)))
        cb = nnz_balanced_bounds(local.indptr, chunks, row_cost=row_cost)
        self.sub = []
        r0 = bounds[rank]
        for c in range(chunks):
            if cb[c + 1] > cb[c]:
                blk = local if chunks == 1 else local.slice_rows(cb[c], cb[c + 1])
                self.sub.append((r0 + cb[c], r0 + cb[c + 1], blk))
        self.copy_stream = torch.cuda.Stream(device=device)
        self.events = [torch.cuda.Event() for _ in self.sub]
        self.done = torch.cuda.Event()
        self._flag = torch.zeros(1, device=device)
        torch.cuda.synchronize()
        dist.barrier()

    @property
    def rows_local(self):
        return self.bounds[self.rank + 1] - self.bounds[self.rank]

    def step(self, x):
        import ctypes as C
        import torch
        ctx, lib = self.ctx, self.ctx.lib
        cur = torch.cuda.current_stream()
        cs = self.copy_stream
        xp = C.c_void_p(x.data_ptr())
        for i, (a0, a1, blk) in enumerate(self.sub):
            ctx.check(lib.sprs_b200_spmv_dev(ctx.h, blk.mirror.h, xp,
                                             C.c_void_p(self._own.value + 8 * a0), 0,
                                             C.c_void_p(cur.cuda_stream)))
            if self.world > 1:
                self.events[i].record(cur)
                cs.wait_event(self.events[i])
                for p in self._peers:
                    ctx.check(lib.sprs_b200_copy_dev(
                        ctx.h, C.c_void_p(p.value + 8 * a0), C.c_void_p(self._own.value + 8 * a0),
                        8 * (a1 - a0), C.c_void_p(cs.cuda_stream)))
        if self.world > 1:
            self.done.record(cs)
            cur.wait_event(self.done)
            self.dist.all_reduce(self._flag)  # barrier: every peer's pushes have landed
        return self.y

    def compute(self, x):  # bench.py times compute and exchange together for this mode
        self.step(x)

    def exchange(self):
        pass

    def close(self):
        import torch
        torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
        for p in self._peers:
            self.ctx.lib.sprs_b200_peer_close(self.ctx.h, p)
        self._peers = []
        if self._own:
            self.y = None
            self.ctx.lib.sprs_b200_peer_free(self.ctx.h, self._own)
            self._own = None
