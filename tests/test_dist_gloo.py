"""CPU tests (gloo, world_size 2) of the multi-GPU host logic in sprs_b200/dist.py: the
nnz-balanced row partition and the unequal-slice all-gather of y.  The local product is
done by the oracle here (the CUDA kernels need a GPU; tests may use the oracle)."""
import os
import socket
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nnz_balanced_bounds_numpy_and_torch():
    import torch
    from sprs_b200.dist import nnz_balanced_bounds
    rng = np.random.default_rng(0)
    lens = (rng.pareto(1.1, 5000) * 20).astype(np.int64)  # skewed rows
    ip = np.zeros(5001, dtype=np.int64)
    np.cumsum(lens, out=ip[1:])
    for parts in (1, 2, 3, 8):
        b = nnz_balanced_bounds(ip, parts)
        assert b[0] == 0 and b[-1] == 5000 and len(b) == parts + 1
        assert all(b[i] <= b[i + 1] for i in range(parts))
        assert b == nnz_balanced_bounds(torch.from_numpy(ip), parts)
        per = [ip[b[i + 1]] - ip[b[i]] for i in range(parts)]
        # every block within one (max) row of the ideal share
        assert max(per) <= ip[-1] / parts + lens.max()
    # non-zero-based indptr (a slice view) and degenerate shapes
    assert nnz_balanced_bounds(ip[100:201], 2)[-1] == 100
    assert nnz_balanced_bounds(np.zeros(1, np.int64), 4) == [0, 0, 0, 0, 0]
    assert nnz_balanced_bounds(np.zeros(11, np.int64), 2) == [0, 0, 10]


def test_row_cost_partition_and_fit():
    from sprs_b200.dist import fit_row_cost, nnz_balanced_bounds
    import torch
    # first half: 10 heavy rows; second half: 10000 light rows with the same total nnz
    lens = np.concatenate([np.full(10, 1000), np.ones(10000, dtype=np.int64)])
    ip = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=ip[1:])
    assert nnz_balanced_bounds(ip, 2)[1] == 10
    b = nnz_balanced_bounds(ip, 2, row_cost=4.0)  # rows cost 4 nnz each: cut moves right
    assert b[1] > 10 and b == nnz_balanced_bounds(torch.from_numpy(ip), 2, row_cost=4.0)
    cost = lambda r0, r1: (ip[r1] - ip[r0]) + 4.0 * (r1 - r0)
    assert abs(cost(0, b[1]) - cost(b[1], len(lens))) <= 1000 + 4
    # fit: t = 2e-9*nnz + 1e-8*rows  ->  row cost 5 nnz
    samples = [(5e8, 1e5, 2e-9 * 5e8 + 1e-8 * 1e5), (5e8, 9e6, 2e-9 * 5e8 + 1e-8 * 9e6)]
    assert abs(fit_row_cost(samples) - 5.0) < 1e-6
    assert fit_row_cost(samples[:1]) == 0.0
    assert fit_row_cost([(1e6, 10, 1.0), (1e6, 20, 0.5)]) == 0.0  # negative beta clamps to 0


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from conftest import rand_csr
    from oracle import oracle as O
    from sprs_b200.dist import RowPartitionedSpMV, nnz_balanced_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(123)  # same matrix on every rank
        n = 3000
        ip, ind, d = rand_csr(rng, n, n, 20, skew=True)
        x = rng.standard_normal(n)
        bounds = nnz_balanced_bounds(ip, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        s = int(ip[r0])

        def local_spmv(xv, y_slice):  # slice_outer + proper_indptr, then the product
            out = np.zeros(r1 - r0)
            O.mul_acc_mat_vec_csr((ip[r0:r1 + 1] - s).astype(np.uint32), ind[s:int(ip[r1])],
                                  d[s:int(ip[r1])], xv.numpy(), out)
            y_slice.copy_(torch.from_numpy(out))

        y = torch.full((n,), float("nan"), dtype=torch.float64)
        op = RowPartitionedSpMV(bounds, rank, world, y, local_spmv, dist=dist)
        got = op.step(torch.from_numpy(x)).numpy().copy()
        ref = np.zeros(n)
        O.mul_acc_mat_vec_csr(ip, ind, d, x, ref)
        ok = bool(np.array_equal(got, ref))  # same sequential sums on every rank
        # iterate: y of step k is the x of step k+1 on every rank (square matrix)
        got2 = op.step(torch.from_numpy(got)).numpy().copy()
        ref2 = np.zeros(n)
        O.mul_acc_mat_vec_csr(ip, ind, d, ref, ref2)
        ok = ok and bool(np.array_equal(got2, ref2))
        q.put((rank, ok, bounds))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_spmv_gloo_world2():
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    assert res[0][2] == res[1][2] and 0 < res[0][2][1] < 3000


def _worker_spmm_spgemm(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    from conftest import rand_csr
    from oracle import oracle as O
    from sprs_b200.dist import RowPartitionedSpGEMM, RowPartitionedSpMM, nnz_balanced_bounds
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(77)  # same operands on every rank
        n, m, k = 1500, 900, 24
        ip, ind, d = rand_csr(rng, n, m, 12, skew=True, empty_frac=0.2)
        bounds = nnz_balanced_bounds(ip, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        s0, s1 = int(ip[r0]), int(ip[r1])
        lip, lind, ld = (ip[r0:r1 + 1] - s0).astype(np.uint32), ind[s0:s1], d[s0:s1]

        # ---- SpMM: rows of C are independent -> bit-identical to the one-process product
        b = rng.standard_normal((m, k))
        ref = np.zeros((n, k))
        O.csr_mulacc_dense_rowmaj(ip, ind, d, b, ref)

        def local_spmm(bt, c_slice):
            out = np.zeros((r1 - r0, k))
            O.csr_mulacc_dense_rowmaj(lip, lind, ld, bt.numpy(), out)
            c_slice.copy_(torch.from_numpy(out))

        c = torch.full((n, k), float("nan"), dtype=torch.float64)
        got = RowPartitionedSpMM(bounds, rank, world, c, local_spmm, dist=dist).step(
            torch.from_numpy(b)).numpy()
        ok = bool(np.array_equal(got, ref))
        c2 = torch.full((n, k), float("nan"), dtype=torch.float64)
        part = RowPartitionedSpMM(bounds, rank, world, c2, local_spmm, dist=dist,
                                  gather=False).step(torch.from_numpy(b)).numpy()
        ok = ok and bool(np.array_equal(part[r0:r1], ref[r0:r1]))
        ok = ok and bool(np.isnan(np.delete(part, np.s_[r0:r1], axis=0)).all())

        # ---- SpGEMM: pieces concatenated with an indptr offset (smmp.rs:320-331, 384-404)
        bip, bind, bd = rand_csr(rng, m, 1100, 9, empty_frac=0.1)
        cip, cind, cd = O.mul_csr_csr((n, m), (ip, ind, d), (m, 1100), (bip, bind, bd), threads=1)

        def local_spgemm():
            pip, pind, pd = O.mul_csr_csr((r1 - r0, m), (lip, lind, ld), (m, 1100), (bip, bind, bd), threads=1)
            return (torch.from_numpy(pip.astype(np.int32)), torch.from_numpy(pind.astype(np.int32)),
                    torch.from_numpy(pd))

        gip, gind, gd, total = RowPartitionedSpGEMM(bounds, rank, world, local_spgemm,
                                                    dist=dist).product()
        ok = ok and total == int(cip[-1])
        ok = ok and bool(np.array_equal(gip.numpy(), cip.astype(np.int32)))
        ok = ok and bool(np.array_equal(gind.numpy(), cind.astype(np.int32)))
        ok = ok and bool(np.array_equal(gd.numpy().view(np.uint64), cd.view(np.uint64)))
        # row-distributed form: this rank's rows with global indptr values
        pip, pind, pd, total2 = RowPartitionedSpGEMM(bounds, rank, world, local_spgemm, dist=dist,
                                                     gather=False).product()
        ok = ok and total2 == total
        ok = ok and bool(np.array_equal(pip.numpy(), cip[r0:r1 + 1].astype(np.int32)))
        ok = ok and bool(np.array_equal(pind.numpy(), cind[int(cip[r0]):int(cip[r1])].astype(np.int32)))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_spmm_spgemm_gloo_world2():
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_spmm_spgemm, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res), res


def _worker_bicgstab(rank, world, port, q):
    """Row-partitioned BiCGSTAB end to end on the CPU: the solver and the local SpMV are the
    library's kernels on the emulator (tests/emu, test infrastructure), the exchange is gloo."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    import ctypes as C
    import scipy.sparse as sparse
    import torch
    import torch.distributed as dist
    import sprs_b200 as sp
    from conftest import emu_library
    from oracle import oracle as O
    from sprs_b200.dist import RowPartitionedSpMV, nnz_balanced_bounds, row_partitioned_bicgstab
    sp._lib.LIB_PATH = emu_library()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(2024)  # same system on every rank
        n = 1200
        A = sparse.random(n, n, density=10 / n, random_state=rng, format="csr")
        A = (A + sparse.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)).tocsr()
        A.sort_indices()
        ip, ind, d = A.indptr.astype(np.uint32), A.indices.astype(np.uint32), A.data.copy()
        b = rng.standard_normal(n)
        bounds = nnz_balanced_bounds(ip, world)
        r0, r1 = bounds[rank], bounds[rank + 1]
        s0, s1 = int(ip[r0]), int(ip[r1])
        ctx = sp.Context.default()
        blk = sp.CsMat((r1 - r0, n), ip[r0:r1 + 1] - s0, ind[s0:s1], d[s0:s1], ctx=ctx)
        mirror = blk.device()

        def local_spmv(xv, y_slice):  # the library's SpMV kernel on this rank's row block
            out = torch.empty(r1 - r0, dtype=torch.float64)
            ctx.check(ctx.lib.sprs_b200_spmv_dev(ctx.h, mirror.h, C.c_void_p(xv.data_ptr()),
                                                 C.c_void_p(out.data_ptr()), 0, None))
            y_slice.copy_(out)

        y = torch.zeros(n, dtype=torch.float64)
        op = RowPartitionedSpMV(bounds, rank, world, y, local_spmv, dist=dist)
        tol = 1e-10
        solver = row_partitioned_bicgstab(ctx, op, n, np.zeros(n), b, "cpu").run(tol, 200)
        x = solver.x()
        ok_ref, ref = O.BiCGSTAB.solve((ip, ind, d), np.zeros(n), b, tol, 200)
        ok = ok_ref and bool(np.allclose(x, ref.x(), rtol=1e-7, atol=1e-10))
        ok = ok and float(np.linalg.norm(b - A @ x)) < tol * 1.001
        ok = ok and solver.hard_restart_count() >= 1
        # an exception inside the operator surfaces as itself, not as a status code
        def broken(d_x, d_y, stream):
            raise KeyError("operator failed")
        try:
            sp.linalg.BiCGSTAB.with_operator(ctx, n, broken, np.zeros(n), b)
            ok = False
        except KeyError:
            pass
        q.put((rank, ok, solver.iteration_count(), x.tobytes()))
    finally:
        dist.destroy_process_group()


def test_row_partitioned_bicgstab_gloo_world2():
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker_bicgstab, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _, _ in res), [(r, ok, it) for r, ok, it, _ in res]
    # every rank took the same steps and holds the same bits without exchanging a scalar
    assert res[0][2] == res[1][2] and res[0][3] == res[1][3]
