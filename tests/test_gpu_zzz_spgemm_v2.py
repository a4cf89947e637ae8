"""SpGEMM round-2 candidate tuning (SPRS_B200_SPGEMM_V2=1, csrc/spgemm.cu): routing of the
middle rows to the bitmap / panel kernels, warp groups for short A rows, the 1024-thread panel
kernel with two chunks in flight and skipped empty panels.  Results must be what the default
kernels produce: indptr / indices bit-exact against the oracle, values within the gate, rows of
<= 128 entries bit-exact.

The switch is read once per process, so the cases run in a child process with the variable
set.  Written after the round's last GPU session: on hardware the file is opt-in
(SPRS_B200_TEST_SPGEMM_V2=1, run by tools/r2_first_call.sh) until the variant has had its
first run there; it always runs in the CPU emulator pre-flight (tests/test_emu_preflight.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import rand_csr
from test_gpu_spgemm_csc import check_spgemm

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
IN_CHILD = os.environ.get("SPRS_B200_SPGEMM_V2") == "1"

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not IN_CHILD and os.environ.get("SPRS_B200_TEST_SPGEMM_V2") != "1"
                                 and os.environ.get("SPRS_B200_EMU") != "1",
                                 reason="opt-in until first hardware run: SPRS_B200_TEST_SPGEMM_V2=1")]


@pytest.fixture(scope="module")
def sp():
    import sprs_b200
    sprs_b200.Context.default()
    return sprs_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def csr_with_row_lengths(rng, lens, cols):
    lens = np.asarray(lens, dtype=np.int64)
    indptr = np.zeros(len(lens) + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.concatenate([np.sort(rng.choice(cols, size=int(n), replace=False))
                              for n in lens] + [np.zeros(0, dtype=np.int64)])
    data = rng.standard_normal(int(indptr[-1]))
    return indptr.astype(np.uint32), indices.astype(np.uint32), data


@pytest.mark.skipif(not IN_CHILD, reason="runs in the child process with SPRS_B200_SPGEMM_V2=1")
@pytest.mark.parametrize("p", [9000, 50000])      # one column panel / three panels (20480 each)
def test_v2_short_a_rows_long_b_rows(sp, O, p):
    """A rows of 1..33 non-zeros (every warp-group width G = 32..1 of the 1024-thread panel
    kernel and G = 8..1 of the 256-thread kernels), B rows long enough to span several chunks
    per panel and to end exactly on / next to chunk boundaries."""
    rng = np.random.default_rng(1000 + p)
    m = 400
    a = csr_with_row_lengths(rng, [1, 1, 2, 3, 4, 5, 7, 8, 9, 15, 16, 17, 31, 32, 33, 0, 2, 1], m)
    blens = rng.integers(0, 700, size=m)
    blens[:16] = [0, 1, 31, 32, 33, 63, 64, 65, 127, 128, 129, 640, 641, 1, 2, 700]
    b = csr_with_row_lengths(rng, blens, p)
    check_spgemm(sp, O, a, b, (len(a[0]) - 1, m), (m, p))


@pytest.mark.skipif(not IN_CHILD, reason="runs in the child process with SPRS_B200_SPGEMM_V2=1")
def test_v2_sparse_rows_skip_empty_panels(sp, O):
    """Rows whose entries fall into few of many panels: untouched panels are skipped, the
    cursors and the output offsets must still line up."""
    rng = np.random.default_rng(2024)
    m, p = 300, 200_000                       # ten panels
    a = csr_with_row_lengths(rng, rng.integers(0, 6, size=50), m)
    # every B row lives inside one or two random panels
    lens = rng.integers(100, 400, size=m)
    ip = np.zeros(m + 1, dtype=np.int64)
    np.cumsum(lens, out=ip[1:])
    idx = []
    for n in lens:
        lo = int(rng.integers(0, 9)) * 20480 + int(rng.integers(0, 15000))
        idx.append(np.sort(rng.choice(np.arange(lo, min(lo + 25000, p)), size=int(n), replace=False)))
    b = (ip.astype(np.uint32), np.concatenate(idx).astype(np.uint32), rng.standard_normal(int(ip[-1])))
    check_spgemm(sp, O, a, b, (50, m), (m, p))


@pytest.mark.skipif(IN_CHILD, reason="parent side")
def test_spgemm_v2_child_process():
    """Every SpGEMM test of the suite plus the cases above, with the variant switched on."""
    env = dict(os.environ, SPRS_B200_SPGEMM_V2="1")
    files = ["test_gpu_spgemm_csc.py", "test_gpu_zz_late.py", "test_gpu_zzz_spgemm_v2.py"]
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] +
        [os.path.join(ROOT, "tests", f) for f in files] +
        ["-k", "(spgemm or mul_cs or zero_rows or structural or v2) and not child_process"
               + (" and not full_size" if os.environ.get("SPRS_B200_EMU") == "1" else "")],
        capture_output=True, text=True, timeout=1500, env=env, cwd=ROOT)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    assert r.returncode == 0, tail + r.stderr[-1500:]
    assert " passed" in tail
