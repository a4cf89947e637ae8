"""Pipelined all-gather (sprs_b200_spmv_stream_push_dev): the SpMV publishes its progress and
a concurrent put kernel copies finished row chunks into the other buffers.  Single process:
the "peer" buffers are ordinary device allocations, which exercises the whole mechanism
(progress counters, epochs, chunk row ranges, put/SpMV concurrency, carry fix-up to all
targets) on one GPU.  Written after the round's last GPU session (sorts last on purpose)."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import rand_csr

# The put kernel spins on counters the SpMV fills: if the two kernels were ever serialised it
# traps after ~3 s and the CUDA context is lost.  Until the mechanism has had its first run
# on hardware these tests are opt-in there (they always run on the CPU emulator pre-flight).
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("SPRS_B200_TEST_STREAM_PUSH") != "1" and
                                 os.environ.get("SPRS_B200_EMU") != "1",
                                 reason="opt-in until first hardware run: "
                                        "SPRS_B200_TEST_STREAM_PUSH=1")]


@pytest.fixture(scope="module")
def sp():
    import sprs_b200
    sprs_b200.Context.default()
    return sprs_b200


def _run_case(sp, monkeypatch, csr, rows, cols, chunk_shift, n_targets=3, offset=5, reps=3,
              accumulate=False, chunked=None):
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    if chunk_shift is None:
        monkeypatch.delenv("SPRS_B200_PUSH_CHUNK_SHIFT", raising=False)
    else:
        monkeypatch.setenv("SPRS_B200_PUSH_CHUNK_SHIFT", str(chunk_shift))
    a = sp.CsMat((rows, cols), *csr)
    dev = a.device()                      # a fresh mirror: the chunking is fixed at first use
    mirror = dev.h
    tdev = G._device(ctx)
    total = rows + offset + 7             # the local block sits at [offset, offset+rows)
    bufs = [torch.full((total,), -777.0, device=tdev, dtype=torch.float64)
            for _ in range(n_targets)]
    ref = torch.full((rows,), -777.0, device=tdev, dtype=torch.float64)
    ptrs = (C.c_void_p * n_targets)(*[b.data_ptr() for b in bufs])
    for rep in range(reps):
        x = G.normal_vector(ctx, cols, seed=100 + rep)
        if accumulate:
            for b in bufs:
                b[offset:offset + rows] = 0.5
            ref[:] = 0.5
        ctx.check(ctx.lib.sprs_b200_spmv_dev(ctx.h, mirror, C.c_void_p(x.data_ptr()),
                                             C.c_void_p(ref.data_ptr()), int(accumulate),
                                             G._stream_ptr()))
        if chunked is not None:   # plan B: chunks + events (sprs_b200_spmv_chunked_push_dev)
            ctx.check(ctx.lib.sprs_b200_spmv_chunked_push_dev(
                ctx.h, mirror, C.c_void_p(x.data_ptr()), offset, n_targets, ptrs, int(accumulate),
                chunked, G._stream_ptr()))
        else:
            ctx.check(ctx.lib.sprs_b200_spmv_stream_push_dev(
                ctx.h, mirror, C.c_void_p(x.data_ptr()), offset, n_targets, ptrs, int(accumulate),
                4, G._stream_ptr()))
        G._sync()
        want = ref.cpu().numpy()
        for q, b in enumerate(bufs):
            got = b.cpu().numpy()
            assert np.array_equal(got[offset:offset + rows], want), (rep, q)
            assert np.all(got[:offset] == -777.0) and np.all(got[offset + rows:] == -777.0), q
    del dev


@pytest.mark.parametrize("chunk_shift", [None, 0, 2])
def test_stream_push_matches_plain_spmv(sp, monkeypatch, chunk_shift):
    rng = np.random.default_rng(77)
    rows, cols = 3000, 2500
    csr = rand_csr(rng, rows, cols, 40, empty_frac=0.1)
    _run_case(sp, monkeypatch, csr, rows, cols, chunk_shift)


def test_stream_push_hub_rows_and_accumulate(sp, monkeypatch):
    """Rows much longer than a tile (runs of carries across chunk boundaries) and y += A x."""
    rng = np.random.default_rng(78)
    rows, cols = 400, 6000
    csr = rand_csr(rng, rows, cols, 300, skew=True)
    _run_case(sp, monkeypatch, csr, rows, cols, 1)
    _run_case(sp, monkeypatch, csr, rows, cols, 3, accumulate=True, n_targets=8)


def test_stream_push_degenerate_shapes(sp, monkeypatch):
    rng = np.random.default_rng(79)
    tiny = rand_csr(rng, 7, 9, 2)                       # a single (ragged) tile
    _run_case(sp, monkeypatch, tiny, 7, 9, None)
    empty = (np.zeros(51, np.uint32), np.zeros(0, np.uint32), np.zeros(0))
    _run_case(sp, monkeypatch, empty, 50, 10, None)     # no non-zeros: y = 0 everywhere
    one_target = rand_csr(rng, 500, 400, 30)
    _run_case(sp, monkeypatch, one_target, 500, 400, 1, n_targets=1)  # world size 1: plain SpMV
    monkeypatch.setenv("SPRS_B200_PUSH_ALWAYS", "1")  # ... or the put kernel for its carries only
    _run_case(sp, monkeypatch, one_target, 500, 400, 1, n_targets=1)
    hub = rand_csr(rng, 60, 5000, 900, skew=True)
    _run_case(sp, monkeypatch, hub, 60, 5000, 0, n_targets=1)


@pytest.mark.parametrize("n_chunks", [0, 1, 3, 8])
def test_chunked_push_matches_plain_spmv(sp, monkeypatch, n_chunks):
    """sprs_b200_spmv_chunked_push_dev: same buffers, same bits, chunk/event pipelining."""
    rng = np.random.default_rng(177)
    rows, cols = 3000, 2500
    csr = rand_csr(rng, rows, cols, 40, empty_frac=0.1)
    _run_case(sp, monkeypatch, csr, rows, cols, None, chunked=n_chunks)


def test_chunked_push_hub_rows_degenerate_shapes(sp, monkeypatch):
    rng = np.random.default_rng(178)
    hub = rand_csr(rng, 400, 6000, 300, skew=True)
    _run_case(sp, monkeypatch, hub, 400, 6000, None, chunked=4)
    _run_case(sp, monkeypatch, hub, 400, 6000, None, chunked=5, accumulate=True, n_targets=8)
    tiny = rand_csr(rng, 7, 9, 2)                       # one ragged tile: a single chunk
    _run_case(sp, monkeypatch, tiny, 7, 9, None, chunked=4)
    empty = (np.zeros(51, np.uint32), np.zeros(0, np.uint32), np.zeros(0))
    _run_case(sp, monkeypatch, empty, 50, 10, None, chunked=4)
    one_target = rand_csr(rng, 500, 400, 30)
    _run_case(sp, monkeypatch, one_target, 500, 400, None, n_targets=1, chunked=3)
    # as many chunks as tiles (5 tiles of 384, 8 chunks asked): the tapered cut points used
    # to collide at the end ("bad tile range", found by tools/fuzz_emu.py)
    lens = np.full(40, 48)                              # 1920 non-zeros = exactly 5 tiles
    ip = np.concatenate([[0], np.cumsum(lens)]).astype(np.uint32)
    ind = np.concatenate([np.sort(rng.choice(300, 48, replace=False)) for _ in lens]).astype(np.uint32)
    _run_case(sp, monkeypatch, (ip, ind, rng.standard_normal(1920)), 40, 300, None, chunked=8)


# ---------------------------------------------------------------- NVSwitch multicast exchange
class _FakeSymmHandle:
    """Stand-in for torch's symmetric-memory handle on ONE device: the "multicast" address is
    a second ordinary buffer (a store to a real multicast address lands in every rank's y;
    here it lands in that buffer), which checks everything McastAllGatherSpMV does itself --
    target order, row offsets, put-kernel arguments, barrier choice."""

    def __init__(self, mc_buf):
        self.mc_buf = mc_buf
        self.multicast_ptr = mc_buf.data_ptr()
        self.barriers = 0

    def barrier(self, channel=0):
        self.barriers += 1


class _FakeDist:
    class group:
        WORLD = "world"

    def __init__(self):
        self.all_reduces = 0

    def barrier(self):
        pass

    def all_reduce(self, t):
        self.all_reduces += 1


@pytest.mark.parametrize("mode,barrier", [("fused", "nccl"), ("push", "symm"), ("fused", "symm"),
                                          ("stream", "nccl"), ("chunked", "symm")])
def test_mcast_allgather_host_logic(sp, monkeypatch, mode, barrier):
    import torch
    import torch.distributed._symmetric_memory as symm
    from sprs_b200 import generate as G
    from sprs_b200.dist import McastAllGatherSpMV
    ctx = sp.Context.default()
    tdev = G._device(ctx)
    rng = np.random.default_rng(5)
    n = 3000
    csr = rand_csr(rng, n, n, 30, skew=True)
    bounds = [0, 1100, n]                       # this process is rank 1 of 2: rows [1100, n)
    r0, r1 = bounds[1], bounds[2]
    ip = csr[0].astype(np.int64)
    s0, s1 = int(ip[r0]), int(ip[r1])
    a = sp.CsMat((r1 - r0, n), (ip[r0:r1 + 1] - s0).astype(np.uint32), csr[1][s0:s1], csr[2][s0:s1])
    mirror = a.device()
    handles = []

    def fake_empty(size, dtype=None, device=None):
        return torch.full((size,), -777.0, dtype=dtype, device=tdev)

    def fake_rendezvous(t, group):
        assert group == "world"
        handles.append(_FakeSymmHandle(torch.full((t.numel(),), -777.0, dtype=t.dtype, device=tdev)))
        return handles[-1]

    monkeypatch.setattr(symm, "empty", fake_empty)
    monkeypatch.setattr(symm, "rendezvous", fake_rendezvous)
    fd = _FakeDist()
    op = McastAllGatherSpMV(ctx, mirror, bounds, 1, 2, n, fd, tdev, mode=mode, barrier=barrier)
    x = G.normal_vector(ctx, n, seed=3)
    ref = torch.zeros(r1 - r0, device=tdev, dtype=torch.float64)
    ctx.check(ctx.lib.sprs_b200_spmv_dev(ctx.h, mirror.h, C.c_void_p(x.data_ptr()),
                                         C.c_void_p(ref.data_ptr()), 0, G._stream_ptr()))
    for _ in range(2):
        op.y.fill_(-1.0)
        handles[0].mc_buf.fill_(-1.0)
        y = op.step(x)
        G._sync()
        want = ref.cpu().numpy()
        own, mc = y.cpu().numpy(), handles[0].mc_buf.cpu().numpy()
        assert np.array_equal(own[r0:r1], want) and np.array_equal(mc[r0:r1], want)
        assert np.all(own[:r0] == -1.0) and np.all(mc[:r0] == -1.0)   # other ranks' rows untouched
    assert (handles[0].barriers, fd.all_reduces) == ((2, 0) if barrier == "symm" else (0, 2))
    op.close()
    # no multicast object -> loud failure, never a silent unicast fallback
    def rendezvous_without_mc(t, group):
        h = _FakeSymmHandle(t)
        h.multicast_ptr = 0
        return h

    monkeypatch.setattr(symm, "rendezvous", rendezvous_without_mc)
    with pytest.raises(sp.ThirdPartyError, match="multicast"):
        McastAllGatherSpMV(ctx, mirror, bounds, 1, 2, n, fd, tdev)
