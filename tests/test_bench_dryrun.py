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


class _FakeMirror:
    h = None


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
        self.ctx.launches += 2
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


def test_bench_flags(fake_gpu, monkeypatch, capsys):
    line = _run_bench(monkeypatch, capsys, ["--steps", "2", "--no-cpu-baseline", "--no-extra"])
    assert "cpu_baseline" not in line and "extra" not in line
    assert line["warmup"] >= 3  # W >= 3 whatever the flag says


# ---- tools/scale_modes.py (all exchange modes in one multi-GPU launch): control-flow dry run
def _scale_modes_worker(rank, world, port, q):
    import os
    import socket  # noqa: F401
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    import io
    import contextlib
    import torch.distributed as dist
    import sprs_b200
    from sprs_b200 import dist as D
    from sprs_b200 import generate as G
    real_device = torch.device
    torch.device = lambda *a, **k: real_device("cpu")
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.empty_cache = lambda: None
    torch.cuda.Event = _FakeEvent
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: real_init("gloo")
    ctx = _FakeCtx()
    sprs_b200.Context.default = classmethod(lambda cls, device=None: ctx)

    def make_matrix(c, gen, n, npr, seed):
        rng = np.random.default_rng(seed % (1 << 32))
        return _FakeCsr(sps.random(n, n, density=npr / n, format="csr", random_state=rng,
                                   data_rvs=rng.standard_normal))

    def spmv(c, a, x, y, accumulate=False):
        y.copy_(torch.from_numpy(a.m @ x.numpy()))
        return y

    G.make_matrix, G.spmv = make_matrix, spmv
    G.normal_vector = lambda c, n, seed=1: torch.from_numpy(np.random.default_rng(seed).standard_normal(n))

    class FakePeerOp(D.RowPartitionedSpMV):  # stands in for every peer-buffer exchange class
        def __init__(self, c, mirror, bounds, rank, world, n, dist_, device, **kw):
            if kw.get("mode") == "chunked" and kw.get("barrier") == "symm":
                raise RuntimeError("pretend this combination is unavailable")  # the skip path
            blk = mirror.owner
            super().__init__(bounds, rank, world, torch.zeros(n, dtype=torch.float64),
                             lambda xv, ys: spmv(c, blk, xv, ys), dist=dist_)

        def close(self):
            pass

    for name in ("PushAllGatherSpMV", "FusedAllGatherSpMV", "StreamAllGatherSpMV",
                 "ChunkedPushAllGatherSpMV", "McastAllGatherSpMV"):
        setattr(D, name, FakePeerOp)
    import scale_modes
    sys.argv = ["scale_modes.py", "--n", "3000", "--npr", "10", "--steps", "2", "--warmup", "1",
                "--modes", "push fused nccl mcast-push mcast-chunked stream"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        scale_modes.main()
    q.put((rank, buf.getvalue()))


def test_scale_modes_control_flow_gloo_world2():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_scale_modes_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lines = [json.loads(l) for l in res[0].splitlines() if l.startswith("{")]
    assert res[1].strip() == ""                      # rank 0 alone reports
    modes = [(d["mode"], d["barrier"]) for d in lines if "mode" in d]
    assert modes == [("push", "nccl"), ("fused", "nccl"), ("nccl", "nccl"), ("mcast-push", "nccl"),
                     ("mcast-push", "symm"), ("mcast-chunked", "nccl"), ("mcast-chunked", "symm"),
                     ("stream", "nccl")]
    done = [d for d in lines if "ms_per_step" in d]
    assert len(done) == 7 and all(d["correct"] and d["speedup_vs_n1"] > 0 for d in done)
    assert [d for d in lines if "skipped" in d][0]["mode"] == "mcast-chunked"
    assert any("partition_round" in d for d in lines) and any("setup_seconds" in d for d in lines)


# ---- bench.py N > 1 control flow (partition calibration, re-cuts, the exchange trial of
#      `--exchange auto`, the JSON line) on two gloo ranks with the GPU pieces faked
def _bench_n2_worker(rank, world, port, q):
    import os
    ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world), LOCAL_RANK=str(rank), SPRS_BENCH_NO_SAMPLER="1",
                      SPRS_B200_AUTO_TRIAL_MIN_GPUS="2")
    import contextlib
    import io
    import torch.distributed as dist
    import torch.distributed._symmetric_memory as symm
    import sprs_b200
    from sprs_b200 import dist as D
    from sprs_b200 import generate as G
    real_device = torch.device
    torch.device = lambda *a, **k: real_device("cpu")
    torch.cuda.set_device = lambda d: None
    torch.cuda.synchronize = lambda *a: None
    torch.cuda.empty_cache = lambda: None
    torch.Tensor.pin_memory = lambda self: self

    class Event(_FakeEvent):
        ms = 1.0

        def elapsed_time(self, other):
            return Event.ms

    torch.cuda.Event = Event
    real_init = dist.init_process_group
    dist.init_process_group = lambda backend, **k: real_init("gloo")
    symm._SymmetricMemory.has_multicast_support = staticmethod(lambda *a: True)
    ctx = _FakeCtx()
    sprs_b200.Context.default = classmethod(lambda cls, device=None: ctx)

    def make_matrix(c, gen, n, npr, seed):
        rng = np.random.default_rng(seed % (1 << 32))
        return _FakeCsr(sps.random(n, n, density=npr / n, format="csr", random_state=rng,
                                   data_rvs=rng.standard_normal))

    def spmv(c, a, x, y, accumulate=False):
        y.copy_(torch.from_numpy(a.m @ x.numpy()))
        return y

    G.make_matrix, G.spmv = make_matrix, spmv
    G.normal_vector = lambda c, n, seed=1: torch.from_numpy(np.random.default_rng(seed).standard_normal(n))
    closed = []

    class FakePeerOp(D.RowPartitionedSpMV):
        speed = 1.0

        def __init__(self, c, mirror, bounds, rank_, world_, n, dist_, device, mode=None, barrier=None):
            self.tag = mode
            if mode == "fused":          # pretend plain `mcast` computes something else
                self.wrong = True
            blk = mirror.owner
            super().__init__(bounds, rank_, world_, torch.zeros(n, dtype=torch.float64),
                             lambda xv, ys: spmv(c, blk, xv, ys), dist=dist_)

        def step(self, xv):
            Event.ms = 0.5 if self.tag == "push" else 1.0   # only mcast-push is "faster"
            out = super().step(xv)
            if getattr(self, "wrong", False):
                out[0] += 1.0
            return out

        def close(self):
            closed.append(self.tag)

    class Plain(FakePeerOp):             # the non-mcast peer classes take no mode/barrier
        def __init__(self, c, mirror, bounds, rank_, world_, n, dist_, device):
            super().__init__(c, mirror, bounds, rank_, world_, n, dist_, device)

    for name in ("PushAllGatherSpMV", "FusedAllGatherSpMV", "StreamAllGatherSpMV",
                 "ChunkedPushAllGatherSpMV"):
        setattr(D, name, Plain)
    D.McastAllGatherSpMV = FakePeerOp
    import bench
    bench.WORKLOADS["spmv_rmat_10m"] = ("spmv", 3000, 10, "rmat")
    sys.argv = ["bench.py", "--gpus", "2", "--steps", "2", "--warmup", "3", "--no-cpu-baseline",
                "--no-extra"]
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        bench.main()
    q.put((rank, buf.getvalue(), closed))


def test_bench_n2_auto_exchange_trial_gloo():
    import socket
    import torch.multiprocessing as mp
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mpc = mp.get_context("spawn")
    q = mpc.Queue()
    procs = [mpc.Process(target=_bench_n2_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = {r: (out, closed) for r, out, closed in (q.get(timeout=240) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lines = [l for l in res[0][0].splitlines() if l.startswith("{")]
    assert len(lines) == 1 and res[1][0].strip() == ""
    line = json.loads(lines[0])
    for k in REQUIRED:
        assert k in line, k
    assert line["n_gpus"] == 2 and line["scaling"] == "strong"
    trial = line["config"]["exchange_trial"]
    # the validated exchange (push at 2 GPUs) was timed, mcast-push beat it and was selected,
    # plain mcast was rejected because its result differed
    # (the fake event reports 1.0 / 0.5 ms for the 5 timed steps)
    assert trial["push"] == 0.2 and trial["mcast-push"] == 0.1 and trial["selected"] == "mcast-push"
    assert trial["mcast"].startswith("rejected")
    assert "multicast" in line["config"]["collective"]
    assert "fused" in res[0][1] and None in res[0][1]      # the loser and the base op were closed
