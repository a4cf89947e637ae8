"""2-GPU test of the row-partitioned SpMV (needs >= 2 GPUs; skipped on a 1-GPU box): the
exchange modes measured in round 1 -- NCCL all_gather, the all-gather fused into the kernel
through CUDA IPC peer stores, the chunked compute / peer-copy overlap, the own put kernel --
must reproduce the single-GPU result on every rank.  The modes written after the last GPU
session (pipelined put `stream`, `chunked` push, NVSwitch multicast) run in a second test that
is opt-in on hardware (SPRS_B200_TEST_STREAM_PUSH=1, set by tools/r2_first_call.sh): a trap or
hang there must not hide the validated modes under `pytest -x`."""
import os
import socket
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q, late_modes=False):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    import sprs_b200 as sp
    from sprs_b200 import generate as G
    from sprs_b200.dist import (FusedAllGatherSpMV, OverlappedAllGatherSpMV, PushAllGatherSpMV,
                                ChunkedPushAllGatherSpMV, McastAllGatherSpMV, RowPartitionedSpMV,
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
        oks = []

        def check_mode(make):
            o = make()
            for _ in range(3):
                o.y.fill_(float("nan"))
                torch.cuda.synchronize()
                dist.barrier()
                g = o.step(x)
                torch.cuda.synchronize()
                oks.append(bool(((g - ref).abs() <= 1e-9 * scale).all()))
                dist.barrier()
            o.close()

        if not late_modes:
            check_mode(lambda: FusedAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev))
            check_mode(lambda: OverlappedAllGatherSpMV(ctx, a, bounds, rank, world, n, dist, dev,
                                                       chunks=3))
            check_mode(lambda: PushAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev))
        if late_modes:
            # the modes whose put kernel WAITS on the SpMV (stream, mcast-stream) run last: they
            # are the only ones that can trap, and a trap loses the CUDA context
            check_mode(lambda: ChunkedPushAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev))
            import torch.distributed._symmetric_memory as symm
            from torch._C._autograd import DeviceType
            has_mc = symm._SymmetricMemory.has_multicast_support(DeviceType.CUDA, dev.index)
            if has_mc:
                for mode in ("fused", "push", "chunked"):
                    for barrier in ("nccl", "symm"):
                        check_mode(lambda: McastAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n,
                                                              dist, dev, mode=mode, barrier=barrier))
            else:
                print("no multicast support on this box: mcast modes not exercised", flush=True)
            # SURVEY 8e "next": row-partitioned SpMM and SpGEMM reproduce the 1-GPU products
            from sprs_b200.dist import RowPartitionedSpGEMM, RowPartitionedSpMM
            k = 16
            b = torch.randn(n, k, device=dev, dtype=torch.float64,
                            generator=torch.Generator(device=dev).manual_seed(3))
            c_ref = torch.zeros(n, k, device=dev, dtype=torch.float64)
            G.spmm_rowmaj(ctx, full, b, c_ref)
            c = torch.full((n, k), float("nan"), device=dev, dtype=torch.float64)
            mm = RowPartitionedSpMM(bounds, rank, world, c,
                                    lambda bt, cs: G.spmm_rowmaj(ctx, a, bt, cs), dist=dist)
            c_got = mm.step(b)
            torch.cuda.synchronize()
            oks.append(bool(torch.equal(c_got, c_ref)))
            small = G.rmat_csr(ctx, 30_000, 8, seed=11)
            sb = nnz_balanced_bounds(small.indptr, world)
            blk = small.slice_rows(sb[rank], sb[rank + 1])
            keep = []

            def local_spgemm():
                m, ip_t, ind_t, dat_t = G.spgemm(ctx, blk, small)
                keep.append(m)
                return ip_t, ind_t, dat_t

            gip, gind, gdat, total = RowPartitionedSpGEMM(sb, rank, world, local_spgemm,
                                                          dist=dist).product()
            m1, ip1, ind1, dat1 = G.spgemm(ctx, small, small)
            torch.cuda.synchronize()
            oks.append(total == ind1.numel() and bool(torch.equal(gip, ip1)) and
                       bool(torch.equal(gind, ind1)))
            # values: rows above 128 entries accumulate with shared-memory atomics (order not
            # fixed), so they agree to rounding, like any two runs of the 1-GPU product
            oks.append(bool(((gdat - dat1).abs() <= 1e-9 * dat1.abs().max()).all()))
            del keep, m1
            # row-partitioned BiCGSTAB (operator form over the put exchange): same bits on
            # every rank, the solution of the 1-GPU solver to rounding
            import numpy as np
            import scipy.sparse as sparse
            from sprs_b200.dist import row_partitioned_bicgstab
            rs = np.random.default_rng(5)
            ns = 20_000
            S = sparse.random(ns, ns, density=8 / ns, random_state=rs, format="csr")
            S = (S + sparse.diags(np.asarray(abs(S).sum(axis=1)).ravel() + 1.0)).tocsr()
            S.sort_indices()
            sip, sind, sd = S.indptr.astype(np.uint32), S.indices.astype(np.uint32), S.data.copy()
            rhs = rs.standard_normal(ns)
            bb = nnz_balanced_bounds(sip, world)
            q0, q1 = bb[rank], bb[rank + 1]
            blkm = sp.CsMat((q1 - q0, ns), sip[q0:q1 + 1] - sip[q0], sind[sip[q0]:sip[q1]],
                            sd[sip[q0]:sip[q1]], ctx=ctx)
            pops = [PushAllGatherSpMV(ctx, blkm.device(), bb, rank, world, ns, dist, dev)
                    for _ in range(2)]   # two y buffers: see row_partitioned_bicgstab
            sol = row_partitioned_bicgstab(ctx, pops, ns, np.zeros(ns), rhs, dev).run(1e-9, 200)
            one = sp.linalg.BiCGSTAB.solve(sp.CsMat((ns, ns), sip, sind, sd, ctx=ctx), np.zeros(ns),
                                           rhs, 1e-9, 200)
            xs = sol.x()
            oks.append(bool(np.allclose(xs, one.x(), rtol=1e-7, atol=1e-10)))
            oks.append(float(np.linalg.norm(rhs - S @ xs)) < 1.001e-9)
            mine = torch.from_numpy(xs).to(dev)
            other = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(other, mine)
            oks.append(all(bool(torch.equal(o, mine)) for o in other))
            del sol
            for po in pops:
                po.close()
            check_mode(lambda: StreamAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist, dev))
            if has_mc:
                check_mode(lambda: McastAllGatherSpMV(ctx, a.mirror, bounds, rank, world, n, dist,
                                                      dev, mode="stream", barrier="nccl"))
        q.put((rank, ok_nccl, all(oks)))
    finally:
        dist.destroy_process_group()


@pytest.mark.gpu
def test_row_partitioned_spmv_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_two_ranks(False)


def _run_two_ranks(late_modes):
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_worker, args=(r, 2, port, q, late_modes)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=600) for _ in procs]
    finally:
        for p in procs:
            p.join(timeout=120)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    assert all(a and b for _, a, b in res), res


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("SPRS_B200_TEST_STREAM_PUSH") != "1",
                    reason="exchange modes without a hardware run yet: opt-in "
                           "(SPRS_B200_TEST_STREAM_PUSH=1, tools/r2_first_call.sh)")
def test_row_partitioned_spmv_two_gpus_late_modes():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    _run_two_ranks(True)
