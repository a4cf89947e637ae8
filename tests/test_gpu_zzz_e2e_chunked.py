"""Chunked host path (SPRS_B200_E2E_CHUNKS=n, api.cu spmv_host_chunked): the tile stream in a
few chunks, each chunk's finished rows copied to the host while the next chunk computes.  The
result must be bit-identical to the one-shot device SpMV (same kernel, same carry sums).

The switch is read once per process: the cases run in a child process with it set (and with
the test hook SPRS_B200_E2E_MIN_TILES=1, so that small matrices take the chunked path)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rand_csr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IN_CHILD = os.environ.get("SPRS_B200_E2E_MIN_TILES") == "1"

pytestmark = [pytest.mark.gpu]


@pytest.fixture(scope="module")
def sp():
    import sprs_b200
    sprs_b200.Context.default()
    return sprs_b200


def _check(sp, csr, rows, cols, accumulate):
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    a = sp.CsMat((rows, cols), *csr)
    mirror = a.device().h
    rng = np.random.default_rng(rows * 7 + cols)
    x = rng.standard_normal(cols)
    y0 = rng.standard_normal(rows) if accumulate else np.full(rows, -777.0)
    # one-shot device SpMV (not affected by the switch)
    tdev = G._device(ctx)
    dx = torch.from_numpy(x).to(tdev)
    dy = torch.from_numpy(y0.copy()).to(tdev)
    ctx.check(ctx.lib.sprs_b200_spmv_dev(ctx.h, mirror, C.c_void_p(dx.data_ptr()),
                                         C.c_void_p(dy.data_ptr()), int(accumulate), G._stream_ptr()))
    G._sync()
    want = dy.cpu().numpy()
    # host path, twice (the chunk table is built on the first call and reused)
    for _ in range(2):
        y = y0.copy()
        if accumulate:
            ctx.check(ctx.lib.sprs_b200_mul_acc_mat_vec_csr(
                ctx.h, mirror, x.ctypes.data_as(C.c_void_p), cols, y.ctypes.data_as(C.c_void_p), rows))
        else:
            ctx.check(ctx.lib.sprs_b200_mul_mat_vec(
                ctx.h, mirror, x.ctypes.data_as(C.c_void_p), cols, y.ctypes.data_as(C.c_void_p), rows))
        assert np.array_equal(y, want)


# (the cases below exist only in the child process, the driver test only in the parent: nothing
# is ever skipped)
@pytest.mark.parametrize("accumulate", [False, True])
def _chunked_host_path_matches_device_spmv(sp, accumulate):
    rng = np.random.default_rng(31)
    rows, cols = 6000, 5000
    _check(sp, rand_csr(rng, rows, cols, 30, empty_frac=0.15), rows, cols, accumulate)


def _chunked_host_path_hub_rows_across_chunks(sp):
    """Rows much longer than a chunk: their carries cross chunk boundaries; trailing and
    leading empty rows belong to the first / last chunk."""
    rng = np.random.default_rng(32)
    rows, cols = 5000, 60000
    ip, idx, dat = rand_csr(rng, rows, cols, 12, empty_frac=0.3)
    lens = np.diff(ip.astype(np.int64))
    lens[:40] = 0
    lens[-55:] = 0
    lens[100] = 20000           # ~20 tiles of 1024: spans several of the 5 chunks
    lens[2500] = 9000
    ip2 = np.zeros(rows + 1, dtype=np.int64)
    np.cumsum(lens, out=ip2[1:])
    idx2 = np.concatenate([np.sort(rng.choice(cols, size=int(n), replace=False)) for n in lens if n])
    dat2 = rng.standard_normal(int(ip2[-1]))
    _check(sp, (ip2.astype(np.uint32), idx2.astype(np.uint32), dat2), rows, cols, False)
    _check(sp, (ip2.astype(np.uint32), idx2.astype(np.uint32), dat2), rows, cols, True)


@pytest.mark.parametrize("chunks", ["5", "8", "1"])
def _e2e_chunked_child_process(chunks):
    env = dict(os.environ, SPRS_B200_E2E_MIN_TILES="1", SPRS_B200_E2E_CHUNKS=chunks)
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider",
         os.path.join(ROOT, "tests", "test_gpu_zzz_e2e_chunked.py"),
         os.path.join(ROOT, "tests", "test_gpu_spmv_spmm.py"),
         "-k", "(chunked or kat or spmv) and not child_process and not full_size and not 1m"],
        capture_output=True, text=True, timeout=1200, env=env, cwd=ROOT)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    assert r.returncode == 0, tail + r.stderr[-1500:]
    assert " passed" in tail


if IN_CHILD:
    test_chunked_host_path_matches_device_spmv = _chunked_host_path_matches_device_spmv
    test_chunked_host_path_hub_rows_across_chunks = _chunked_host_path_hub_rows_across_chunks
else:
    test_e2e_chunked_child_process = _e2e_chunked_child_process
