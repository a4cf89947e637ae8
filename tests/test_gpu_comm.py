"""Multi-rank path through the library's communicator (include/sprs_b200.h: comm_init_rank,
symm_alloc, spmv_rowpart, mul_mat_vec_rowpart) on hardware.  The ranks are separate processes
that meet through the 64-byte id only (no torch.distributed); rank r uses device r % n_devices,
so on a single-GPU box both ranks share device 0 (CUDA IPC between two processes on one
device) and the whole multi-rank logic still runs; with >= 2 GPUs the same tests take distinct
devices and, where the box has NVSwitch multicast, the multicast binding.
Every rank's WHOLE all-gathered y is compared with the CPU oracle (prod.rs:274-298 restated,
oracle/sprs_oracle.cpp), gate |d| <= 1e-6 * sum|terms|."""
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _n_devices():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("world", [2, 3])
def test_comm_ranks_through_the_c_header_only(world):
    """tests/cpp/test_comm_ranks.cpp: `world` processes driven through sprs_b200.h alone --
    rendezvous, partition, symmetric x / y, every exchange mode of spmv_rowpart twice from a
    NaN-poisoned y, the host-slice form -- each rank checking the full product."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_comm_ranks")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    env = dict(os.environ, SPRS_TEST_NDEV=str(max(1, _n_devices())))
    r = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "OK " in r.stdout, r.stdout + r.stderr


def _rank_main(rank, world, q_id, q_out, ndev):
    sys.path.insert(0, ROOT)
    import torch
    import sprs_b200 as sp
    from oracle import oracle as O
    from sprs_b200 import generate as G
    from sprs_b200.dist import Comm, CommHostSpMV, CommSpMV, nnz_balanced_bounds
    try:
        device = rank % ndev
        torch.cuda.set_device(device)
        dev = torch.device("cuda", device)
        ctx = sp.Context.default(device)
        comm = Comm(ctx, q_id.get(timeout=120), rank, world)
        n = 120_000
        full = G.rmat_csr(ctx, n, 40, seed=7)       # every rank builds the same matrix
        x = G.normal_vector(ctx, n, 9)
        hip, hind, hdat = full.to_host()
        hx = x.cpu().numpy()
        ref, bound = np.zeros(n), np.zeros(n)
        O.mul_acc_mat_vec_csr(hip, hind, hdat, hx, ref)
        O.mul_acc_mat_vec_csr(hip, hind, np.abs(hdat), np.abs(hx), bound)
        bounds = nnz_balanced_bounds(full.indptr, world, row_cost=8.0)
        r0, r1 = bounds[rank], bounds[rank + 1]
        a = full.slice_rows(r0, r1)
        report = {"multicast_supported": comm.multicast, "modes": {}}
        for mode in ("push", "fused", "auto"):
            for mc in (False, True):
                op = CommSpMV(comm, a.mirror, bounds, n, dev, exchange=mode, multicast=mc)
                worst = 0.0
                for _ in range(2):
                    op.y.fill_(float("nan"))      # a row that never arrives fails the gate
                    torch.cuda.synchronize()
                    comm.barrier_host()
                    got = op.step(x).cpu().numpy()
                    comm.check()
                    err = np.abs(got - ref) / (1e-6 * bound + 1e-300)
                    worst = max(worst, float(np.nanmax(err)) if np.all(np.isfinite(got)) else np.inf)
                    comm.barrier_host()
                report["modes"]["%s%s" % (mode, "+mc" if op.multicast else "")] = worst
                op.close()
        hop = CommHostSpMV(comm, a.mirror, bounds, n, multicast=True)
        hx_slice = torch.from_numpy(hx[r0:r1].copy()).pin_memory()
        hy = torch.empty(max(r1 - r0, 1), dtype=torch.float64).pin_memory()
        for _ in range(2):
            hop.step(hx_slice.data_ptr(), hy.data_ptr())
        err = np.abs(hy[:r1 - r0].numpy() - ref[r0:r1]) / (1e-6 * bound[r0:r1] + 1e-300)
        report["host_slices"] = float(err.max()) if err.size else 0.0
        hop.close()
        # x uploaded in EQUAL column slices, whatever the row cut (bench.py's e2e form)
        xb = [n * g // world for g in range(world + 1)]
        hop = CommHostSpMV(comm, a.mirror, bounds, n, multicast=True, x_bounds=xb)
        hx_slice = torch.from_numpy(hx[xb[rank]:xb[rank + 1]].copy()).pin_memory()
        hy.fill_(float("nan"))
        for _ in range(2):
            hop.step(hx_slice.data_ptr(), hy.data_ptr())
        err = np.abs(hy[:r1 - r0].numpy() - ref[r0:r1]) / (1e-6 * bound[r0:r1] + 1e-300)
        report["host_equal_x_slices"] = (float(np.nanmax(err)) if np.all(np.isfinite(hy[:r1 - r0].numpy()))
                                         else float("inf")) if err.size else 0.0
        hop.close()
        comm.close()
        q_out.put((rank, report))
    except Exception as e:  # report instead of leaving the parent to time out
        import traceback
        q_out.put((rank, {"error": repr(e), "trace": traceback.format_exc()[-1500:]}))


def test_comm_two_ranks_python_vs_oracle():
    import torch.multiprocessing as mp
    import sprs_b200 as sp
    from sprs_b200.dist import Comm
    world, ndev = 2, max(1, _n_devices())
    mpc = mp.get_context("spawn")
    q_id, q_out = mpc.Queue(), mpc.Queue()
    cid = Comm.unique_id()
    for _ in range(world):
        q_id.put(cid)
    procs = [mpc.Process(target=_rank_main, args=(r, world, q_id, q_out, ndev)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q_out.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
    for r in range(world):
        assert "error" not in res[r], res[r]
        for mode, worst in res[r]["modes"].items():
            assert worst <= 1.0, (r, mode, worst)
        assert res[r]["host_slices"] <= 1.0, (r, res[r])
        assert res[r]["host_equal_x_slices"] <= 1.0, (r, res[r])
    if ndev >= 2 and res[0]["multicast_supported"]:
        assert any(m.endswith("+mc") for m in res[0]["modes"]), res[0]
    assert sp is not None
