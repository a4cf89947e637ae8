"""GPU parity tests for SpMV / SpMM through the C ABI (via the Python mirror of the
sprs operator API).  Each KAT names the reference test it replays; random cases are
checked against the CPU oracle (oracle/) with the SURVEY 8d gate
    |got - ref| <= 1e-6 * sum_j |a_ij * x_j|      (f64 values, tolerance parity)
and, where the kernel sums in storage order, bit-exactly."""
import numpy as np
import pytest

from conftest import mat_arrays, rand_csr

pytestmark = pytest.mark.gpu

RTOL = 1e-6  # BASELINE north_star: "within 1e-6 relative on f64 values"


@pytest.fixture(scope="module")
def sp():
    import sprs_b200
    sprs_b200.Context.default()  # raises without a GPU / without the .so: no fallback
    return sprs_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def csmat(sp, m, idx=np.uint64):
    ip, ind, d = mat_arrays(m, idx)
    ctor = sp.CsMat.new if m["storage"] == "CSR" else sp.CsMat.new_csc
    return ctor(tuple(m["shape"]), ip, ind, d)


def gate(got, ref, bound):
    assert np.all(np.abs(got - ref) <= RTOL * bound + 1e-300), \
        "max excess %g" % np.max(np.abs(got - ref) - RTOL * bound)


# ---------------------------------------------------------------- reference KATs
@pytest.mark.parametrize("idx", [np.uint32, np.uint64])
def test_mul_csr_vec(sp, fixtures, idx):
    """prod.rs:376-398 mul_csr_vec (direct kernel entry, res_vec accumulates)."""
    k = fixtures["kat_mul_csr_vec"]
    mat = csmat(sp, k["mat"], idx)
    res_vec = np.zeros(5)
    sp.prod.mul_acc_mat_vec_csr(mat, np.array(k["x"]), res_vec)
    assert np.all(np.abs(res_vec - np.array(k["expected"])) < k["epsilon"])
    assert res_vec[1] == 0.0
    # operator form `&A * &x` (csmat.rs:2119-2160)
    y = mat * np.array(k["x"])
    assert np.all(np.abs(y - np.array(k["expected"])) < k["epsilon"])


def test_mul_acc_accumulates(sp, fixtures):
    """prod.rs:121-125: y += A x."""
    k = fixtures["kat_mul_csr_vec"]
    mat = csmat(sp, k["mat"])
    y = np.arange(5, dtype=np.float64)
    sp.prod.mul_acc_mat_vec_csr(mat, np.array(k["x"]), y)
    assert np.all(np.abs(y - (np.arange(5) + np.array(k["expected"]))) < k["epsilon"])


def test_panics(sp, fixtures):
    """prod.rs:114-118: "Dimension mismatch" before "Storage mismatch"."""
    mat = csmat(sp, fixtures["mat1"])
    with pytest.raises(sp.SprsPanic, match="Dimension mismatch"):
        sp.prod.mul_acc_mat_vec_csr(mat, np.zeros(4), np.zeros(5))
    with pytest.raises(sp.SprsPanic, match="Dimension mismatch"):
        mat * np.zeros(6)
    with pytest.raises(sp.SprsPanic, match="Storage mismatch"):
        sp.prod.mul_acc_mat_vec_csc(mat, np.zeros(5), np.zeros(5))
    # the C ABI itself reports the same codes when called with a wrong length
    ctx = mat.context()
    x, y = np.zeros(4), np.zeros(5)
    st = ctx.lib.sprs_b200_mul_acc_mat_vec_csr(ctx.h, mat.device().h, x.ctypes.data, 4,
                                               y.ctypes.data, 5)
    assert st == sp._lib.ERR_DIMENSION
    st = ctx.lib.sprs_b200_mul_acc_mat_vec_csc(ctx.h, mat.device().h, y.ctypes.data, 5,
                                               y.ctypes.data, 5)
    assert st == sp._lib.ERR_STORAGE


def test_mul_csr_dense_rowmaj(sp, fixtures):
    """prod.rs:503-542 mul_csr_dense_rowmaj: eye, mat1*dense1 (exact), mat5*dense2 (1e-8),
    and `&a * &b` equals the direct kernel (:523)."""
    e = sp.CsMat.eye(3)
    a = np.eye(3)
    res = np.zeros((3, 3))
    sp.prod.csr_mulacc_dense_rowmaj(e, a, res)
    assert np.array_equal(res, a)

    m1 = csmat(sp, fixtures["mat1"])
    b = np.array(fixtures["mat_dense1"])
    res = np.zeros((5, 5))
    sp.prod.csr_mulacc_dense_rowmaj(m1, b, res)
    assert np.array_equal(res, np.array(fixtures["kat_mat1_x_dense1"]))
    c = m1 * b  # 5 columns < 8 -> colmaj kernel, F-order result (csmat.rs:2009)
    assert np.array_equal(c, np.array(fixtures["kat_mat1_x_dense1"]))
    assert c.flags.f_contiguous

    m5 = csmat(sp, fixtures["mat5"])
    b = np.array(fixtures["mat_dense2"])
    res = np.zeros((5, 7))
    sp.prod.csr_mulacc_dense_rowmaj(m5, b, res)
    k = fixtures["kat_mat5_x_dense2"]
    assert np.all(np.abs(res - np.array(k["expected"])) <= k["epsilon"])


def test_mul_csr_dense_colmaj(sp, fixtures):
    """prod.rs:581-595 mul_csr_dense_colmaj: F-order rhs/out, exact integers."""
    m1 = csmat(sp, fixtures["mat1"])
    b = np.asfortranarray(np.array(fixtures["mat_dense1"]))
    res = np.zeros((5, 5), order="F")
    sp.prod.csr_mulacc_dense_colmaj(m1, b, res)
    assert np.array_equal(res.flatten(order="F"),
                          np.array(fixtures["kat_mat1_x_dense1_colmaj_flat"]))


def test_wide_operator_is_c_order(sp, O):
    """csmat.rs:2009-2018: k >= 8 -> rowmaj kernel, C-order result; bit-exact vs oracle
    (the SpMM kernel sums each element sequentially and unfused, like the reference)."""
    rng = np.random.default_rng(5)
    ip, ind, d = rand_csr(rng, 300, 200, 9, empty_frac=0.1)
    a = sp.CsMat.new((300, 200), ip, ind, d)
    for k in (8, 33, 64, 100):
        b = rng.standard_normal((200, k))
        c = a * b
        assert c.flags.c_contiguous
        ref = np.zeros((300, k))
        O.csr_mulacc_dense_rowmaj(ip, ind, d, b, ref)
        assert np.array_equal(c, ref)


def test_sparse_dot_dense_layouts(sp, fixtures):
    """prod.rs:618-651 test_sparse_dot_dense restricted to CSR operands (CSC ones are in
    test_gpu_csc.py): transposed / F-order / sliced dense views, rtol 1e-7 atol 1e-12."""
    tol = fixtures["assert_close"]
    d1, d2 = np.array(fixtures["mat_dense1"]), np.array(fixtures["mat_dense2"])
    dense = [d1, np.asfortranarray(d1), d1.T, d2, d2.T]
    for name in ("mat1", "mat2", "mat5"):
        s = csmat(sp, fixtures[name])
        for dn in dense:
            if dn.shape[0] < s.cols():
                continue
            dv = dn[:s.cols(), :]
            truth = s.to_dense().dot(dv)
            test = s.dot(dv)
            assert np.all(np.abs(test - truth) <= np.abs(truth) * tol["rtol"] + tol["atol"])


def test_readme_eye_times_csvec(sp, fixtures):
    """sprs/src/lib.rs:54-60 (BASELINE config 1) and prod.rs:461-474."""
    r = fixtures["kat_readme_eye"]
    x = sp.CsVec(5, r["x"]["indices"], r["x"]["data"])
    assert sp.CsMat.eye(5) * x == x
    k = fixtures["kat_csvec"]
    v = sp.CsVec(5, k["v"]["indices"], k["v"]["data"])
    res = csmat(sp, fixtures["mat1"]) * v
    assert res == sp.CsVec(5, k["mat1_times_v"]["indices"], k["mat1_times_v"]["data"])
    zero = sp.CsVec(0, [], [])
    assert csmat(sp, fixtures["mat1"]) * zero == zero  # mul_csr_zero_csvec


# ---------------------------------------------------------------- oracle parity, random
@pytest.mark.parametrize("case", [
    dict(rows=1, cols=1, npr=1),
    dict(rows=17, cols=5, npr=2),
    dict(rows=3000, cols=2500, npr=3, empty=0.5),        # many empty rows, G=1 path
    dict(rows=2000, cols=3000, npr=32),                  # cfg2-like rows, lane groups
    dict(rows=500, cols=40000, npr=400),                 # long rows: warp queue + carries
    dict(rows=4000, cols=4000, npr=40, skew=True),       # power-law rows
    dict(rows=3, cols=100000, npr=30000),                # rows spanning many tiles
    dict(rows=50000, cols=50, npr=0.05),                 # hypersparse: tiles own 1000s of rows
])
def test_spmv_vs_oracle(sp, O, case):
    rng = np.random.default_rng(case["rows"] * 7919 + case["cols"])
    ip, ind, d = rand_csr(rng, case["rows"], case["cols"], case["npr"],
                          skew=case.get("skew", False), empty_frac=case.get("empty", 0.0))
    a = sp.CsMat.new((case["rows"], case["cols"]), ip, ind, d)
    x = rng.standard_normal(case["cols"])
    ref = np.zeros(case["rows"])
    O.mul_acc_mat_vec_csr(ip, ind, d, x, ref)
    bound = np.zeros(case["rows"])
    O.mul_acc_mat_vec_csr(ip, ind, np.abs(d), np.abs(x), bound)
    got = a * x
    gate(got, ref, bound)
    # accumulate form
    y0 = rng.standard_normal(case["rows"])
    y = y0.copy()
    sp.prod.mul_acc_mat_vec_csr(a, x, y)
    gate(y, y0 + ref, bound + np.abs(y0))
    # deterministic: same bits on a second run (no atomics in the reduction)
    assert np.array_equal(got, a * x)


def test_spmv_short_rows_bit_exact(sp, O):
    """Every row of at most 8 non-zeros is summed by ONE lane in storage order with unfused
    mul/add -> identical bits to the reference's sequential sum (mul_acc.rs:28-30), provided the
    row is not cut by a (merge-path) tile boundary; rows cut by one add their two partial
    sums, longer rows use lane groups: both are held to the tolerance gate instead."""
    rng = np.random.default_rng(11)
    ip, ind, d = rand_csr(rng, 20000, 5000, 3, empty_frac=0.2)
    a = sp.CsMat.new((20000, 5000), ip, ind, d)
    x = rng.standard_normal(5000)
    ref = np.zeros(20000)
    O.mul_acc_mat_vec_csr(ip, ind, d, x, ref)
    got = a * x
    s, e = ip[:-1].astype(np.int64), ip[1:].astype(np.int64)
    inside = ~sp.spmv_rows_cut_by_tiles(ip) & (e - s <= 8)
    assert inside.sum() > 19500
    assert np.array_equal(got[inside], ref[inside])
    bound = np.zeros(20000)
    O.mul_acc_mat_vec_csr(ip, ind, np.abs(d), np.abs(x), bound)
    gate(got, ref, bound)


def test_spmv_empty_and_zero_shapes(sp):
    """Edge cases: nnz == 0, rows == 0, all-empty rows (SURVEY 8a edge semantics)."""
    z = sp.CsMat.zero((7, 3))
    assert np.array_equal(z * np.ones(3), np.zeros(7))
    y = np.arange(7.0)
    sp.prod.mul_acc_mat_vec_csr(z, np.ones(3), y)
    assert np.array_equal(y, np.arange(7.0))
    e = sp.CsMat.zero((0, 4))
    assert (e * np.ones(4)).shape == (0,)


def test_sliced_indptr_upload(sp, O, fixtures):
    """slice_outer view with a non-zero-based indptr (indptr.rs:122-124) is rebased on
    upload, like proper_indptr() (csmat.rs:919-921)."""
    rng = np.random.default_rng(3)
    ip, ind, d = rand_csr(rng, 1000, 800, 12)
    a = sp.CsMat.new((1000, 800), ip, ind, d)
    x = rng.standard_normal(800)
    full = a * x
    part = a.slice_outer(300, 700)
    assert part.indptr[0] != 0
    ref, bound = np.zeros(1000), np.zeros(1000)
    O.mul_acc_mat_vec_csr(ip, ind, d, x, ref)
    O.mul_acc_mat_vec_csr(ip, ind, np.abs(d), np.abs(x), bound)
    # the slice is tiled from its own first non-zero, so sums may round differently
    gate(part * x, ref[300:700], bound[300:700])
    gate(full, ref, bound)


def test_spmv_linearity_and_scaling_full_size(sp):
    """Size-independent properties at a BASELINE-scale shape (1M x 1M, 32 nnz/row
    generated on the device): A(ax + by) ~= a Ax + b Ay, and row sums via x = 1."""
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    n = 1_000_000
    a = G.rand_csr(ctx, n, n, 32, seed=0x5EED0002)
    assert a.nnz == 32_000_000
    ip = a.indptr.to(torch.int64)
    assert bool((ip[1:] >= ip[:-1]).all()) and int(ip[-1]) == a.nnz
    # strictly ascending columns inside every row (sprs invariant)
    idx = a.indices.to(torch.int64)
    same_row = torch.ones(a.nnz - 1, dtype=torch.bool, device=idx.device)
    starts = ip[1:-1]
    starts = starts[(starts > 0) & (starts < a.nnz)]
    same_row[starts - 1] = False
    assert bool(((idx[1:] > idx[:-1]) | ~same_row).all())
    x1, x2 = G.normal_vector(ctx, n, 1), G.normal_vector(ctx, n, 2)
    y1, y2, y3 = (torch.empty(n, device=x1.device, dtype=torch.float64) for _ in range(3))
    G.spmv(ctx, a, x1, y1)
    G.spmv(ctx, a, x2, y2)
    G.spmv(ctx, a, 2.0 * x1 - 3.0 * x2, y3)
    absrow = torch.empty_like(y1)
    absa = G.DeviceCsr(ctx, n, n, a.indptr, a.indices, a.data.abs())
    G.spmv(ctx, absa, (2.0 * x1).abs() + (3.0 * x2).abs(), absrow)
    torch.cuda.synchronize()
    assert bool(((y3 - (2.0 * y1 - 3.0 * y2)).abs() <= 1e-12 * absrow + 1e-300).all())
    # sample 2000 rows against the oracle on the host
    from oracle import oracle as O
    rows = torch.randint(0, n, (2000,), generator=torch.Generator().manual_seed(0)).tolist()
    hip, hx = a.indptr.cpu().numpy(), x1.cpu().numpy()
    for r in rows[:2000]:
        s, e = int(hip[r]), int(hip[r + 1])
        ci = a.indices[s:e].cpu().numpy().view(np.uint32)
        cv = a.data[s:e].cpu().numpy()
        ref = np.zeros(1)
        O.mul_acc_mat_vec_csr(np.array([0, e - s], np.uint32), ci, cv, hx, ref)
        bound = float(np.sum(np.abs(cv * hx[ci])))
        assert abs(float(y1[r]) - ref[0]) <= RTOL * bound + 1e-300


@pytest.mark.parametrize("offsets", [(0, 0, 0), (5, 5, 5), (0, 4, 2), (1, 2, 1), (0, 3), (2, 2)])
def test_spmv_allgather_targets_identical_bits_full_size(sp, offsets, monkeypatch):
    """(device-generated matrix: hardware only, like the other *_full_size tests; the emulator
    covers the same entry point through tools/fuzz_emu.py.)
    sprs_b200_spmv_allgather_dev (the fused all-gather of the multi-GPU path with every target
    on this device): targets 1.. receive the rows of a tile as ONE TMA bulk store from shared
    memory (odd first / last rows as plain stores); every target must hold the plain SpMV's bits,
    and nothing outside the row block may be touched.  Targets whose addresses differ in
    16-byte parity (fourth case) take the store-per-row path; a single remote target (the
    multicast form) defaults to plain stores, so the two-target cases force the staged form."""
    import ctypes as C
    if len(offsets) == 2:
        monkeypatch.setenv("SPRS_B200_SPMV_PEER_STORES", "tma")
    import torch
    from sprs_b200 import generate as G
    ctx = sp.Context.default()
    n = 300_000
    a = G.rmat_csr(ctx, n, 24, seed=21)
    x = G.normal_vector(ctx, n, 5)
    ref = torch.empty(n, device="cuda", dtype=torch.float64)
    G.spmv(ctx, a, x, ref)
    pad = 8
    bufs = [torch.full((n + 2 * pad,), -7.0, device="cuda", dtype=torch.float64) for _ in offsets]
    ptrs = (C.c_void_p * len(bufs))(*[b.data_ptr() + 8 * o for b, o in zip(bufs, offsets)])
    torch.cuda.synchronize()
    ctx.check(ctx.lib.sprs_b200_spmv_allgather_dev(ctx.h, a.mirror.h, C.c_void_p(x.data_ptr()), 0,
                                                  len(bufs), ptrs, 0, None))
    torch.cuda.synchronize()
    for b, o in zip(bufs, offsets):
        assert torch.equal(b[o:o + n].view(torch.int64), ref.view(torch.int64)), o
        assert bool((b[:o] == -7.0).all()) and bool((b[o + n:] == -7.0).all()), o
