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
