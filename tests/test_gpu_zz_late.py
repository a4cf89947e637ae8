"""GPU tests added after the round's last GPU session (kept in a file that sorts last so a
surprise here cannot hide the validated suites under `pytest -x`): the triplet KATs through
the device COO->CSR path, a hub-row SpGEMM case, and full-size property tests of BASELINE
configs 5 (10M x 10M R-MAT SpMV) and 3 (1M x 64 SpMM)."""
import numpy as np
import pytest

from conftest import mat_arrays, rand_csr
from test_gpu_spgemm_csc import check_spgemm, csmat

pytestmark = pytest.mark.gpu
RTOL = 1e-6


@pytest.fixture(scope="module")
def sp():
    import sprs_b200
    sprs_b200.Context.default()
    return sprs_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def test_spgemm_hub_rows(sp, O):
    """Rows of A with > 4096 non-zeros take the global dense-accumulator kernel."""
    rng = np.random.default_rng(4242)
    n, m, p = 12, 7000, 30000
    a = rand_csr(rng, n, m, 5000)
    b = rand_csr(rng, m, p, 25)
    check_spgemm(sp, O, a, b, (n, m), (m, p))


def test_triplet_kats(sp, fixtures, O):
    """triplet.rs:342-453 triplet_incremental / unordered / additions / from_vecs and
    :571-580 triplet_empty_lines, through the device COO->CSR path."""
    for name, k in fixtures["kat_triplets"].items():
        m = sp.CsMat.from_triplets(tuple(k["shape"]), k["rows"], k["cols"], k["data"])
        assert m.is_csr() and m.shape == tuple(k["shape"])
        if "expected_csc" in k:
            e = csmat(sp, k["expected_csc"])
            assert m.to_csc() == e, name          # csr_to_csc == expected (triplet.rs:392-394)
            oip, oind, od = O.triplets_to_csr(k["shape"], k["rows"], k["cols"], k["data"], np.uint64)
            assert m.indptr.tolist() == oip.tolist() and m.indices.tolist() == oind.tolist()
            assert m.data.tolist() == od.tolist(), name
        else:
            assert m.indptr.tolist() == k["expected_csr_indptr"] and m.nnz() == 0


def _sampled_rows_vs_oracle(a, x_host, y_dev, rows, O):
    hip = a.indptr.cpu().numpy()
    for r in rows:
        s, e = int(hip[r]), int(hip[r + 1])
        ci = a.indices[s:e].cpu().numpy().view(np.uint32)
        cv = a.data[s:e].cpu().numpy()
        ref = np.zeros(1)
        O.mul_acc_mat_vec_csr(np.array([0, e - s], np.uint32), ci, cv, x_host, ref)
        bound = float(np.sum(np.abs(cv * x_host[ci])))
        assert abs(float(y_dev[r]) - ref[0]) <= RTOL * bound + 1e-300, r


def check_spmv_rmat(sp, O, n, per_row, nnz_range, n_heavy, n_rnd):
    """R-MAT SpMV properties: linearity A(ax+by) = a Ax + b Ay within rounding, A 0 = 0
    exactly, the structure check, and sampled rows (the heaviest, the first ones and random
    ones) against the oracle."""
    import ctypes
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    a = G.rmat_csr(ctx, n, per_row, seed=0x5EED0005)
    assert nnz_range[0] < a.nnz < nnz_range[1]
    bad = ctypes.c_uint64(1)
    ctx.check(ctx.lib.sprs_b200_csmat_check_structure(ctx.h, a.mirror.h, ctypes.byref(bad)))
    assert bad.value == 0
    x1, x2 = G.normal_vector(ctx, n, 1), G.normal_vector(ctx, n, 2)
    y1, y2, y3 = (torch.empty(n, device=x1.device, dtype=torch.float64) for _ in range(3))
    G.spmv(ctx, a, x1, y1)
    G.spmv(ctx, a, x2, y2)
    G.spmv(ctx, a, 2.0 * x1 - 3.0 * x2, y3)
    absrow = torch.empty_like(y1)
    absa = G.DeviceCsr(ctx, n, n, a.indptr, a.indices, a.data.abs())
    G.spmv(ctx, absa, (2.0 * x1).abs() + (3.0 * x2).abs(), absrow)
    G._sync()
    assert bool(((y3 - (2.0 * y1 - 3.0 * y2)).abs() <= 1e-12 * absrow + 1e-300).all())
    G.spmv(ctx, a, torch.zeros_like(x1), y3)
    G._sync()
    assert bool((y3 == 0).all())
    lens = (a.indptr[1:] - a.indptr[:-1]).to(torch.int64)
    heavy = torch.topk(lens, n_heavy).indices.tolist()
    rnd = torch.randint(0, n, (n_rnd,), generator=torch.Generator().manual_seed(1)).tolist()
    _sampled_rows_vs_oracle(a, x1.cpu().numpy(), y1, heavy + list(range(20)) + rnd, O)


def test_spmv_rmat_10m_full_size(sp, O):
    """BASELINE config 5 at full size: 10M x 10M R-MAT, ~1e9 nnz, generated on the device."""
    check_spmv_rmat(sp, O, 10_000_000, 100, (0.98e9, 1.02e9), 40, 300)


def test_spmv_rmat_small(sp, O):
    """The same properties on a 30k x 30k R-MAT (also what the CPU emulator pre-flight runs)."""
    check_spmv_rmat(sp, O, 30_000, 20, (0.58e6, 0.62e6), 10, 60)


def check_spmm_rand(sp, O, n, k, n_rows):
    """sprs-rand A (32 nnz/row) times an n x k C-order B: sampled output rows are
    bit-identical to the oracle's csr_mulacc_dense_rowmaj."""
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    a = G.rand_csr(ctx, n, n, 32, seed=0x5EED0002)
    assert a.nnz == 32 * n
    b = torch.randn(n, k, device=a.data.device, dtype=torch.float64,
                    generator=torch.Generator(device=a.data.device).manual_seed(3))
    c = torch.empty(n, k, device=b.device, dtype=torch.float64)
    G.spmm_rowmaj(ctx, a, b, c)
    G._sync()
    hip = a.indptr.cpu().numpy()
    rows = torch.randint(0, n, (n_rows,), generator=torch.Generator().manual_seed(2)).tolist()
    for r in rows:
        s, e = int(hip[r]), int(hip[r + 1])
        ci = a.indices[s:e].to(torch.int64)
        cv = a.data[s:e].cpu().numpy()
        bsub = b[ci].cpu().numpy()                      # the B rows this A row touches
        ref = np.zeros((1, k))
        O.csr_mulacc_dense_rowmaj(np.array([0, e - s], np.uint32), np.arange(e - s, dtype=np.uint32),
                                  cv, bsub, ref)
        assert np.array_equal(c[r].cpu().numpy(), ref[0]), r


def test_spmm_1m_k64_full_size(sp, O):
    """BASELINE config 3 at full size (1M x 1M sprs-rand times 1M x 64, C-order)."""
    check_spmm_rand(sp, O, 1_000_000, 64, 200)


def test_spmm_small(sp, O):
    check_spmm_rand(sp, O, 4000, 64, 100)


def test_matrix_market_to_device(sp):
    """io.rs:682-695 read_symmetric_matrix_market: file -> TriMat (host parser) -> device
    COO->CSR -> CSC equals the reference's expected matrix; simple.mm round-trips."""
    import os
    from sprs_b200 import io as mm
    data = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matrix_market")
    csc = mm.read_matrix_market(os.path.join(data, "symmetric.mm")).to_csc()
    expected = sp.CsMat.new_csc((5, 5), [0, 1, 3, 4, 6, 8], [0, 1, 3, 2, 1, 4, 3, 4],
                                [1., 10.5, 2.505e2, 1.5e-2, 2.505e2, 3.332e1, 3.332e1, 1.2e1])
    assert csc == expected
    csr = mm.read_matrix_market(os.path.join(data, "simple.mm")).to_csr()
    assert csr.nnz() == 8 and csr.indptr.tolist() == [0, 2, 3, 4, 7, 8]
    assert csr.indices.tolist() == [0, 3, 1, 2, 1, 3, 4, 4]
    assert csr.data.tolist() == [1., 6., 10.5, 1.5e-2, 2.505e2, -2.8e2, 3.332e1, 1.2e1]


def test_indptr64_kernels():
    """The uint64-indptr instantiations (taken for nnz >= 2^32, far beyond test sizes) through
    the SPRS_B200_FORCE_INDPTR64 test hook: every SpMV / SpMM / conversion / solver test again
    in a child process (SpGEMM is 32-bit-indptr only and says so)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, SPRS_B200_FORCE_INDPTR64="1")
    files = ["test_gpu_spmv_spmm.py", "test_gpu_spgemm_csc.py", "test_gpu_zz_late.py",
             "test_gpu_zzz_solver.py"]
    r = subprocess.run(
        [sys.executable, "-m", "pytest", "-q", "-x", "-m", "gpu", "-p", "no:cacheprovider"] +
        [os.path.join(root, "tests", f) for f in files] +
        ["-k", "not spgemm and not csc_csc and not csc_csr and not issue_99 and not "
               "structural_zeros and not csvec and not full_size and not test_cpp and not "
               "indptr64"],
        capture_output=True, text=True, timeout=1500, env=env, cwd=root)
    tail = "\n".join(r.stdout.splitlines()[-15:])
    assert r.returncode == 0, tail + r.stderr[-1500:]
    assert " passed" in tail


# ---------------------------------------------------------------- CSR x sparse vector
def _csvec_case(rng, rows, cols, per_row, v_nnz, idx_dtype=np.uint32):
    ip, ind, d = rand_csr(rng, rows, cols, per_row, idx_dtype, empty_frac=0.1)
    vi = np.sort(rng.choice(cols, v_nnz, replace=False))
    vd = rng.standard_normal(v_nnz)
    return ip, ind, d, vi, vd


@pytest.mark.parametrize("shape", [(300, 400, 12, 90), (2000, 700, 40, 350), (64, 5000, 300, 4000),
                                   (50, 50, 50, 0), (1, 1, 1, 1)])
def test_csr_mul_csvec_bit_exact_vs_oracle(sp, O, shape):
    """prod.rs:162-184 / vec.rs:846-881: sequential sum over the pattern intersection in
    ascending column order -- the device result is the oracle's to the last bit, and rows
    whose dot is exactly zero (or that meet nothing) are dropped."""
    rows, cols, per_row, v_nnz = shape
    rng = np.random.default_rng(rows * 7 + cols)
    ip, ind, d, vi, vd = _csvec_case(rng, rows, cols, min(per_row, cols), v_nnz)
    for idx in (np.uint32, np.uint64):
        a = sp.CsMat.new((rows, cols), ip.astype(idx), ind.astype(idx), d)
        res = a * sp.CsVec(cols, vi, vd)
        oi, od = O.csr_mul_csvec(ip, ind, d, vi, vd)
        assert res.dim == rows
        assert np.array_equal(res.indices, oi.astype(np.int64))
        assert np.array_equal(res.data.view(np.uint64), od.view(np.uint64))
    # the free function of prod.rs, and its contract checks in the reference's order
    assert sp.prod.csr_mul_csvec(a, sp.CsVec(cols, vi, vd)) == res
    with pytest.raises(sp.SprsPanic, match="Dimension mismatch"):
        sp.prod.csr_mul_csvec(a, sp.CsVec(cols + 1, vi, vd))
    assert sp.prod.csr_mul_csvec(a, sp.CsVec(0, [], [])) == sp.CsVec.empty(0)


def test_csr_mul_csvec_structural_zeros_non_finite(sp, O):
    """An A entry opposite a STRUCTURAL zero of v takes no part in the merge dot
    (vec.rs:862-876): an Inf/NaN stored there leaves the row finite, where a dense x with
    explicit zeros would give Inf*0 = NaN.  Explicitly stored zeros of v do take part; NaN
    results are kept (`val != N::zero()`), -0.0 and cancelled sums are dropped."""
    ip = np.array([0, 3, 5, 7, 9, 9, 10], dtype=np.uint32)
    ind = np.array([0, 1, 3,  1, 2,  0, 3,  2, 4,  4], dtype=np.uint32)
    d = np.array([2.0, np.inf, 5.0,  np.nan, 1.5,  1.0, -1.0,  np.inf, 7.0,  -0.0])
    vi, vd = np.array([0, 2, 3, 4]), np.array([3.0, 0.0, 3.0, 1.0])   # v[2] is a STORED zero
    a = sp.CsMat.new((6, 5), ip, ind, d)
    res = a * sp.CsVec(5, vi, vd)
    oi, od = O.csr_mul_csvec(ip, ind, d, vi, vd)
    assert np.array_equal(res.indices, oi.astype(np.int64))
    assert np.array_equal(res.data.view(np.uint64), od.view(np.uint64))
    # row 0: 2*3 + 5*3 (Inf at column 1 skipped); row 1: 1.5*0 = 0 dropped (NaN at column 1
    # skipped); row 2: 3 - 3 = 0 dropped; row 3: Inf*0 + 7 = NaN kept; row 4 empty; row 5: -0.0
    assert res.indices.tolist() == [0, 3]
    assert res.data[0] == 21.0 and np.isnan(res.data[1])


def test_csvec_dot_by_binary_search(sp, O, fixtures):
    """prod.rs:13-72 and its test (:312-323): the device result equals the reference's
    answers and, on random vectors, the oracle's bits."""
    k = fixtures["kat_csvec_dot"]
    vec = {n: sp.CsVec(k["dim"], k[n]["indices"], k[n]["data"]) for n in ("vec1", "vec2", "vec3")}
    for n1, n2, want in k["expected"]:
        assert sp.prod.csvec_dot_by_binary_search(vec[n1], vec[n2]) == want
    rng = np.random.default_rng(8)
    for n1, n2 in ((5, 900), (900, 5), (300, 300), (0, 4), (1, 1)):
        i1, i2 = (np.sort(rng.choice(1000, n, replace=False)) for n in (n1, n2))
        d1, d2 = rng.standard_normal(n1), rng.standard_normal(n2)
        got = sp.prod.csvec_dot_by_binary_search(sp.CsVec(1000, i1, d1), sp.CsVec(1000, i2, d2))
        assert got == O.csvec_dot_by_binary_search(i1, d1, i2, d2)


def test_csvec_dot(sp, fixtures):
    """vec.rs:1649-1689 dot_product / dot_product_panics / dot_product_panics2."""
    k = fixtures["kat_csvec_dot"]
    vec = {n: sp.CsVec(k["dim"], k[n]["indices"], k[n]["data"]) for n in ("vec1", "vec2", "vec3")}
    for n1, n2, want in k["expected"]:
        assert vec[n1].dot(vec[n2]) == want
    dense = np.array(k["dense"])
    assert vec["vec1"].dot(dense) == k["vec1_dot_dense"]
    assert vec["vec1"].dot_dense(list(dense)) == k["vec1_dot_dense"]
    assert vec["vec1"].dot_dense(np.linspace(1., 8., 8)) == k["vec1_dot_dense"]
    with pytest.raises(sp.SprsPanic):
        vec["vec1"].dot(sp.CsVec(k["panic_dims"]["sparse"], k["vec2"]["indices"], k["vec2"]["data"]))
    with pytest.raises(sp.SprsPanic):
        vec["vec1"].dot(np.arange(float(k["panic_dims"]["dense"])))


def test_from_triplets_rejects_out_of_range_through_the_c_abi(sp):
    """A raw C caller (no host mirror in front) handing an out-of-range triplet gets
    ERR_STRUCTURE instead of an out-of-bounds device write (the reference asserts in
    TriMatBase::add_triplet, triplet.rs)."""
    import ctypes as C
    from sprs_b200 import _lib
    ctx = sp.Context.default()
    r = np.array([0, 1, 7], dtype=np.uint64)      # row 7 in a 4 x 4 matrix
    c = np.array([0, 2, 1], dtype=np.uint64)
    d = np.array([1.0, 2.0, 3.0])
    h = C.c_void_p()
    st = ctx.lib.sprs_b200_csmat_from_triplets(ctx.h, 4, 4, 3, r.ctypes.data_as(C.c_void_p),
                                               c.ctypes.data_as(C.c_void_p), 8,
                                               d.ctypes.data_as(C.c_void_p), C.byref(h))
    assert st == _lib.ERR_STRUCTURE and not h.value
    assert b"outside" in ctx.lib.sprs_b200_last_error(ctx.h)


def _whole_vector_vs_oracle(sp, O, a, x):
    """y = A x on the device against the CPU oracle for EVERY row (gate 1e-6 * sum|terms|)."""
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    y = torch.empty(a.rows, device=x.device, dtype=torch.float64)
    G.spmv(ctx, a, x, y)
    G._sync()
    hip, hind, hdat = a.to_host()
    hx = x.cpu().numpy()
    ref, bound = np.zeros(a.rows), np.zeros(a.rows)
    O.mul_acc_mat_vec_csr(hip, hind, hdat, hx, ref)
    np.abs(hdat, out=hdat)
    O.mul_acc_mat_vec_csr(hip, hind, hdat, np.abs(hx), bound)
    got = y.cpu().numpy()
    assert np.all(np.isfinite(got))
    excess = np.abs(got - ref) - (RTOL * bound + 1e-300)
    assert excess.max() <= 0.0, (int(np.argmax(excess)), float(excess.max()))


def test_spmv_rand_1m_whole_vector_full_size(sp, O):
    """BASELINE config 2 (1M x 1M sprs-rand, 32 nnz/row): every one of the 1e6 outputs against
    the oracle (prod.rs:274-298 restated)."""
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    a = G.rand_csr(ctx, 1_000_000, 1_000_000, 32, seed=0x5EED0002)
    _whole_vector_vs_oracle(sp, O, a, G.normal_vector(ctx, 1_000_000))


def test_spmv_rmat_10m_whole_vector_full_size(sp, O):
    """BASELINE config 5 (10M x 10M R-MAT, ~1e9 nnz): every one of the 1e7 outputs against the
    oracle -- the CPU port walks the whole matrix in a few seconds."""
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    a = G.rmat_csr(ctx, 10_000_000, 100, seed=0x5EED0005)
    _whole_vector_vs_oracle(sp, O, a, G.normal_vector(ctx, 10_000_000))


def test_spgemm_rmat_500k_row_block_bit_exact_full_size(sp, O):
    """BASELINE config 4 at full size: the rows of C = A B for a block of A holding >= 1e8
    products -- indptr and indices BIT-EXACT against the oracle (smmp.rs:81-131, 409-415),
    values within the gate (smmp.rs:151-189)."""
    import ctypes as C_
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    n = 500_000
    A = G.rmat_csr(ctx, n, 16, seed=0x5EED0004)
    B = G.rmat_csr(ctx, n, 16, seed=0x5EED1004)
    cmir, cip, cind, cdat = G.spgemm(ctx, A, B)
    # a row block of A with >= 1e8 products: n_prod_i = sum_k nnz(B_k) over A_i
    aip = (A.indptr.to(torch.int64) & 0xFFFFFFFF)
    blen = ((B.indptr[1:].to(torch.int64) & 0xFFFFFFFF) - (B.indptr[:-1].to(torch.int64) & 0xFFFFFFFF))
    per_nnz = blen[A.indices.to(torch.int64) & 0xFFFFFFFF]
    csum = torch.cumsum(per_nnz, 0)
    r0 = 2000
    base = int(csum[int(aip[r0]) - 1].item()) if int(aip[r0]) > 0 else 0
    k_end = int(torch.searchsorted(csum, torch.tensor([base + 100_000_000], device=csum.device))[0])
    r1 = min(n, int(torch.searchsorted(aip, torch.tensor([k_end], device=aip.device))[0]) + 1)
    nprod_blk = int(csum[int(aip[r1]) - 1].item()) - base
    assert nprod_blk >= 100_000_000
    blk = A.slice_rows(r0, r1)
    oip, oind, odat = O.mul_csr_csr((r1 - r0, n), blk.to_host(), (n, n), B.to_host(), threads=0)
    cip64 = cip.to(torch.int64)
    if cip.dtype == torch.int32:
        cip64 &= 0xFFFFFFFF
    s, e = int(cip64[r0]), int(cip64[r1])
    assert np.array_equal((cip64[r0:r1 + 1] - s).cpu().numpy(), np.asarray(oip, dtype=np.int64))
    assert np.array_equal(cind[s:e].cpu().numpy().view(np.uint32), np.asarray(oind, dtype=np.uint32))
    got = cdat[s:e].cpu().numpy()
    # gate per entry: 1e-6 * sum |a_ik b_kj|, from the same product on absolute values
    absA = (blk.to_host()[0], blk.to_host()[1], np.abs(blk.to_host()[2]))
    bh = B.to_host()
    _, _, obound = O.mul_csr_csr((r1 - r0, n), absA, (n, n), (bh[0], bh[1], np.abs(bh[2])), threads=0)
    assert np.all(np.abs(got - np.asarray(odat)) <= RTOL * np.asarray(obound) + 1e-300)
    del cmir
