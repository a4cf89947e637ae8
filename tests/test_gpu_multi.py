"""2-GPU test of the row-partitioned SpMV (needs >= 2 GPUs; skipped on a 1-GPU box): all
exchange modes -- NCCL all_gather, the all-gather fused into the kernel through CUDA IPC peer
stores, the chunked compute / peer-copy overlap, the own put kernel and the pipelined put
(stream) -- must reproduce the single-GPU result on every rank."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import sprs_b200 as sp
    from sprs_b200 import generate as G
    from sprs_b200.dist import (FusedAllGatherSpMV, OverlappedAllGatherSpMV, PushAllGatherSpMV,
                                ChunkedPushAllGatherSpMV, RowPartitionedSpMV,
                                StreamAllGatherSpMV, nnz_balanced_bounds)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        ctx = sp.Context.default(rank)
        n = 200_000
        full = G.rmat_csr(ctx, n, 40, seed=7)
        x = G.normal_vector(ctx, n, 9)
        ref = torch.empty(n, device=dev, dtype=torch.float64)
        G.spmv(ctx, full, x, ref)
        bounds = nnz_balanced_bounds(full.indptr, world)
        a = full.slice_rows(bounds[rank], bounds[rank + 1])
        y = torch.full((n,), float("nan"), device=dev, dtype=torch.float64)
        op = RowPartitionedSpMV(bounds, rank, world, y, lambda xv, ys: G.spmv(ctx, a, xv, ys),
                                dist=dist)
        got = op.step(x)
        torch.cuda.synchronize()
        scale = ref.abs().max().item()
        ok_nccl = bool(((got - ref).abs() <= 1e-9 * scale).all())
        fop = FusedAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev)
        oks = []
        for _ in range(3):
            fop.y.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            g2 = fop.step(x)
            torch.cuda.synchronize()
            oks.append(bool(((g2 - ref).abs() <= 1e-9 * scale).all()))
            dist.barrier()
        fop.close()
        oop = OverlappedAllGatherSpMV(ctx, a, bounds, rank, world, n, dist, dev, chunks=3)
        for _ in range(3):
            oop.y.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            g3 = oop.step(x)
            torch.cuda.synchronize()
            oks.append(bool(((g3 - ref).abs() <= 1e-9 * scale).all()))
            dist.barrier()
        oop.close()
        pop = PushAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev)
        for _ in range(3):
            pop.y.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            g4 = pop.step(x)
            torch.cuda.synchronize()
            oks.append(bool(((g4 - ref).abs() <= 1e-9 * scale).all()))
            dist.barrier()
        pop.close()
        cop = ChunkedPushAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev)
        for _ in range(3):
            cop.y.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            g6 = cop.step(x)
            torch.cuda.synchronize()
            oks.append(bool(((g6 - ref).abs() <= 1e-9 * scale).all()))
            dist.barrier()
        cop.close()
        sop = StreamAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev)
        for _ in range(3):
            sop.y.fill_(float("nan"))
            torch.cuda.synchronize()
            dist.barrier()
            g5 = sop.step(x)
            torch.cuda.synchronize()
            oks.append(bool(((g5 - ref).abs() <= 1e-9 * scale).all()))
            dist.barrier()
        sop.close()
        q.put((rank, ok_nccl, all(oks)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_row_partitioned_spmv_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert all(a and b for _, a, b in res), res
