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
