"""CPU dry run of bench.py's N=1 control flow with the GPU pieces replaced by fakes: catches
Python-level breakage (names, JSON contract keys) of the headline benchmark without a GPU.
The fake SpMV is scipy; nothing here says anything about performance."""
import json
import sys
import types

import numpy as np
import pytest
import scipy.sparse as sps
import torch


class _FakeEvent:
    def __init__(self, enable_timing=False):
        pass

    def record(self, stream=None):
        pass

    def elapsed_time(self, other):
        return 1.0


def _skew_event(rank):
    """Events for the N > 1 dry runs: rank-dependent, changing "timings" (0.3 or 0.05 per rank,
    alternating every 6 readings) so that bench.py's measured re-cut loop really re-cuts,
    rebuilds its operator and falls back to the best cut."""
    class Ev(_FakeEvent):
        count = 0

        def elapsed_time(self, other):
            Ev.count += 1
            return 1.0 + rank * (0.3 if ((Ev.count - 1) // 6) % 2 == 0 else 0.05)
    return Ev


class _FakeMirror:
    @property
    def h(self):
        return self


class _FakeCsr:
    def __init__(self, m):
        m = m.tocsr()
        m.sort_indices()
        self.m = m
        self.rows, self.cols = m.shape
        self.indptr = torch.from_numpy(m.indptr.astype(np.int32))
        self.indices = torch.from_numpy(m.indices.astype(np.int32))
        self.data = torch.from_numpy(m.data.astype(np.float64))
        self.nnz = int(m.nnz)
        self.mirror = _FakeMirror()
        self.mirror.owner = self

    def slice_rows(self, r0, r1):
        return _FakeCsr(self.m[r0:r1])


class _FakeLib:
    def __init__(self, ctx):
        self.ctx = ctx

    def sprs_b200_mul_mat_vec(self, h, mh, xp, n, yp, rows):
        # y = A x into the caller's (fake-pinned) host buffer, like the real call
        from sprs_b200.dist import tensor_view
        a = mh.owner if hasattr(mh, "owner") else None
        self.ctx.launches += 2
        if a is not None:
            x = tensor_view(xp.value, n, "cpu")
            tensor_view(yp.value, rows, "cpu").copy_(torch.from_numpy(a.m @ x.numpy()))
        return 0

    def sprs_b200_diag_gather_ceiling(self, h, mh, xp, iters, ms_ref, cov_ref):
        ms_ref._obj.value = 0.5
        cov_ref._obj.value = 1000
        return 0


class _FakeCtx:
    def __init__(self):
        self.launches = 0
        self.h = None
        self.device = 0
        self.lib = _FakeLib(self)

    def check(self, st):
        assert st == 0


@pytest.fixture
def fake_gpu(monkeypatch):
    import sprs_b200
    from sprs_b200 import generate as G
    real_device = torch.device
    monkeypatch.setattr(torch, "device", lambda *a, **k: real_device("cpu"))
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "Event", _FakeEvent)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    ctx = _FakeCtx()
    monkeypatch.setattr(sprs_b200.Context, "default", classmethod(lambda cls, device=None: ctx))

    def make_matrix(c, gen, n, npr, seed):
        rng = np.random.default_rng(seed % (1 << 32))
        return _FakeCsr(sps.random(n, n, density=npr / n, format="csr", random_state=rng,
                                   data_rvs=rng.standard_normal))

    def spmv(c, a, x, y, accumulate=False):
        c.launches += 2
        res = torch.from_numpy(a.m @ x.numpy())
        y.copy_(y + res if accumulate else res)
        return y

    monkeypatch.setattr(G, "make_matrix", make_matrix)
    monkeypatch.setattr(G, "rand_csr", lambda c, rows, cols, npr, seed=0: _FakeCsr(
        sps.eye(rows, cols, format="csr") * 2.0))
    monkeypatch.setattr(G, "normal_vector", lambda c, n, seed=1: torch.from_numpy(
        np.random.default_rng(seed).standard_normal(n)))
    monkeypatch.setattr(G, "spmv", spmv)
    return ctx


REQUIRED = ["metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step",
            "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config", "roofline",
            "e2e", "gpu_launches", "clocks"]


def _run_bench(monkeypatch, capsys, argv):
    import bench
    monkeypatch.setitem(bench.WORKLOADS, "spmv_rmat_10m", ("spmv", 4000, 12, "rmat"))
    monkeypatch.setattr(sys, "argv", ["bench.py"] + argv)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    bench.main()
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    assert len(out) == 1, "rank 0 must print exactly one JSON line"
    return json.loads(out[0])


def test_bench_n1_contract_keys(fake_gpu, monkeypatch, capsys):
    line = _run_bench(monkeypatch, capsys, ["--steps", "3", "--warmup", "3"])
    for k in REQUIRED:
        assert k in line, k
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["dtype"] == "f64"
    assert line["metric"] == "csr_spmv_f64_gflops" and line["unit"] == "GFLOP/s"
    assert line["config"]["workload"] == "spmv_rmat_10m" and "model" not in line["config"]
    r = line["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-12
    assert set(["value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"]) <= set(line["e2e"])
    cb = line["cpu_baseline"]  # oracle port, 1 thread
    assert cb["kind"] == "port" and cb["cores"] == 1 and cb["value"] > 0
    assert line["gpu_launches"] == 2 * 3
    assert "spmv_rand_1m" in line["extra"]
    par = line["parity_vs_oracle"]
    assert par["ok"] and par["max_error_over_gate"] <= 1.0 and par["rows_checked_per_rank"] > 1000
    assert line["e2e"]["matches_device_result"]
    assert line["roofline"]["gather_ceiling"]["gnnz_s"] > 0 and "frac_of_gather_ceiling" in line["roofline"]


def test_bench_flags(fake_gpu, monkeypatch, capsys):
    line = _run_bench(monkeypatch, capsys, ["--steps", "2", "--no-cpu-baseline", "--no-extra"])
    assert "cpu_baseline" not in line and "extra" not in line
    assert line["warmup"] >= 3  # W >= 3 whatever the flag says



# ---- N > 1 control flow on two gloo ranks with the GPU pieces faked: the library's
#      communicator is replaced by a gloo-backed stand-in with the same Python face
#      (sprs_b200.dist.Comm / CommSpMV / CommHostSpMV), so that bench.py's and
#      tools/scale_modes.py's multi-rank Python (partition calibration, re-cuts, oracle parity on
#      every rank, the e2e slices, the JSON lines) runs where no GPU is attached.
def _install_fakes(rank, world, port):
    import os
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank), SPRS_BENCH_NO_SAMPLER="1")
    import torch.distributed as dist
    import sprs_b200
    from sprs_b200 import dist as D
    from sprs_b200 import generate as G
    real_device = torch.device
    torch.device = lambda *a, **k: real_device("cpu")
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.Event = _skew_event(rank)
    torch.Tensor.pin_memory = lambda self: self
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: real_init("gloo")
    ctx = _FakeCtx()
    sprs_b200.Context.default = classmethod(lambda cls, device=None: ctx)

    def make_matrix(c, gen, n, npr, seed):
        rng = np.random.default_rng(seed % (1 << 32))
        return _FakeCsr(sps.random(n, n, density=npr / n, format="csr", random_state=rng,
                                   data_rvs=rng.standard_normal))

    def spmv(c, a, x, y, accumulate=False):
        c.launches += 2
        y.copy_(torch.from_numpy(a.m @ x.numpy()))
        return y

    G.make_matrix, G.spmv = make_matrix, spmv
    G.normal_vector = lambda c, n, seed=1: torch.from_numpy(np.random.default_rng(seed).standard_normal(n))

    class FakeComm:
        multicast = False

        @staticmethod
        def unique_id(c=None):
            return b"dryrun".ljust(64, b"\0")

        def __init__(self, c, comm_id, rank_, world_):
            assert len(comm_id) == 64
            self.ctx, self.rank, self.world = c, rank_, world_

        def allgather(self, record):
            out = [None] * self.world
            dist.all_gather_object(out, record)
            return out

        def allgather_f64(self, values):
            rec = np.asarray(values, dtype=np.float64).tobytes()
            return np.stack([np.frombuffer(r, dtype=np.float64) for r in self.allgather(rec)])

        def barrier_host(self):
            dist.barrier()

        def barrier_dev(self, stream=None):
            dist.barrier()

        def check(self, stream=None):
            pass

        def close(self):
            pass

    class FakeCommSpMV(D.RowPartitionedSpMV):
        def __init__(self, comm, mirror, bounds, n, device, exchange="auto", multicast=True):
            assert exchange in D.EXCHANGES
            blk = mirror.owner
            self.multicast = False
            super().__init__(bounds, comm.rank, comm.world, torch.zeros(n, dtype=torch.float64),
                             lambda xv, ys: spmv(comm.ctx, blk, xv, ys), dist=dist)

        def close(self):
            pass

    class _FakeX:
        multicast_ptr = 0

    class FakeCommHostSpMV:
        def __init__(self, comm, mirror, bounds, n, multicast=True, x_bounds=None):
            self.comm, self.blk, self.bounds, self.n, self.x = comm, mirror.owner, bounds, n, _FakeX()
            self.x_bounds = x_bounds if x_bounds is not None else bounds
            assert self.x_bounds[0] == 0 and self.x_bounds[-1] == n

        def step(self, x_ptr, y_ptr):
            r0, r1 = self.bounds[self.comm.rank], self.bounds[self.comm.rank + 1]
            c0, c1 = self.x_bounds[self.comm.rank], self.x_bounds[self.comm.rank + 1]
            xs = D.tensor_view(x_ptr, max(c1 - c0, 1), "cpu")[:c1 - c0]
            parts = [None] * self.comm.world
            dist.all_gather_object(parts, xs.numpy().copy())
            xf = np.concatenate(parts)
            D.tensor_view(y_ptr, max(r1 - r0, 1), "cpu")[:r1 - r0].copy_(torch.from_numpy(self.blk.m @ xf))

        def close(self):
            pass

    D.Comm, D.CommSpMV, D.CommHostSpMV = FakeComm, FakeCommSpMV, FakeCommHostSpMV
    return ctx


def _bench_n2_worker(rank, world, port, q, exchange):
    import contextlib
    import io
    _install_fakes(rank, world, port)
    import bench
    bench.WORKLOADS["spmv_rmat_10m"] = ("spmv", 3000, 10, "rmat")
    sys.argv = ["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--no-cpu-baseline",
                "--no-extra", "--exchange", exchange]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    q.put((rank, buf.getvalue()))


def _spawn2(target, extra=()):
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=target, args=(r, 2, port, q) + tuple(extra)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return res


@pytest.mark.parametrize("exchange", ["auto", "nccl"])
def test_bench_n2_control_flow_gloo(exchange):
    res = _spawn2(_bench_n2_worker, (exchange,))
    lines = [l for l in res[0].splitlines() if l.startswith("{")]
    assert len(lines) == 1 and res[1].strip() == ""       # rank 0 alone reports
    line = json.loads(lines[0])
    for k in REQUIRED:
        assert k in line, k
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    par = line["parity_vs_oracle"]
    assert par["ok"] and par["rows_checked_per_rank"] > 1500 and par["max_error_over_gate"] <= 1.0
    e2e = line["e2e"]
    assert e2e["matches_device_result"] and e2e["h2d_bytes_per_step"] == 8 * 3000 == e2e["d2h_bytes_per_step"]
    assert len(line["roofline"]["kernel_ms_per_rank"]) == 2
    assert ("NCCL" in line["config"]["collective"]) == (exchange == "nccl")
    # the skewed stand-in timings must have driven the measured re-cut loop (operator rebuilt)
    import re
    recuts = int(re.search(r"then (\d+) measured", line["config"]["partition"]).group(1))
    assert recuts >= 1, line["config"]["partition"]


def _scale_modes_worker(rank, world, port, q):
    import contextlib
    import io
    _install_fakes(rank, world, port)
    import scale_modes
    sys.argv = ["scale_modes.py", "--n", "3000", "--npr", "10", "--steps", "2", "--warmup", "1",
                "--modes", "nccl push fused push+mc"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        scale_modes.main()
    q.put((rank, buf.getvalue()))


def test_scale_modes_control_flow_gloo_world2():
    res = _spawn2(_scale_modes_worker)
    lines = [json.loads(l) for l in res[0].splitlines() if l.startswith("{")]
    assert res[1].strip() == ""
    modes = [d["mode"] for d in lines if "mode" in d]
    assert modes == ["nccl", "push", "fused", "push+mc"]
    done = [d for d in lines if "ms_per_step" in d and "mode" in d]
    assert len(done) == 3 and all(d["correct"] and d["speedup_vs_n1"] > 0 for d in done)
    assert [d for d in lines if "skipped" in d][0]["mode"] == "push+mc"   # the stand-in has no multicast
    assert any("partition_round" in d for d in lines) and any("setup_seconds" in d for d in lines)
    assert sum(1 for d in lines if "e2e_host_slices" in d and d.get("correct")) == 3   # same cut x2, pcie cut


def test_extra_child_entry_and_parent_parse(fake_gpu, monkeypatch, capsys):
    """`bench.py --extra-only spmv_rand_1m` (the config-2 extra in a process of its own): the child
    entry prints one JSON object with the entry's keys; the parent takes the last JSON line of a
    successful child and returns None (-> in-process measurement) for anything else."""
    import subprocess
    import bench
    assert bench.extra_child("spmv_rand_1m") == 0
    out = [l for l in capsys.readouterr().out.splitlines() if l.startswith("{")]
    child = json.loads(out[-1])
    assert child["ms"] > 0 and child["parity_vs_oracle"]["ok"] and "child process" in child["process"]

    def fake_run(rc, stdout):
        return lambda *a, **k: subprocess.CompletedProcess(a, rc, stdout=stdout, stderr="")
    monkeypatch.setattr(subprocess, "run", fake_run(0, "noise\n" + out[-1] + "\n"))
    assert bench.run_extra_in_child("spmv_rand_1m")["ms"] == child["ms"]
    monkeypatch.setattr(subprocess, "run", fake_run(1, out[-1]))
    assert bench.run_extra_in_child("spmv_rand_1m") is None
    monkeypatch.setattr(subprocess, "run", fake_run(0, "no json here"))
    assert bench.run_extra_in_child("spmv_rand_1m") is None
    monkeypatch.setattr(subprocess, "run", fake_run(0, '{"error": "x"}'))
    assert bench.run_extra_in_child("spmv_rand_1m") is None
