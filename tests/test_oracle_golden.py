"""Pins the CPU oracle (oracle/sprs_oracle.cpp) against the reference's own
known-answer tests.  Each test names the reference test it replays.
CPU only (-m "not gpu")."""
import numpy as np
import pytest
import scipy.sparse as sp

from conftest import mat_arrays, rand_csr
from oracle import oracle as O

WIDTHS = [(np.uint32, np.uint32), (np.uint64, np.uint64), (np.uint32, np.uint64)]


@pytest.mark.parametrize("I,P", WIDTHS)
def test_mul_csr_vec(fixtures, I, P):
    """prod.rs:376-398 mul_csr_vec / :401-423 mul_csr_vec_ndarray."""
    k = fixtures["kat_mul_csr_vec"]
    ip, ind, d = mat_arrays(k["mat"], I, P)
    y = np.zeros(5)
    O.mul_acc_mat_vec_csr(ip, ind, d, np.array(k["x"]), y)
    assert np.all(np.abs(y - np.array(k["expected"])) < k["epsilon"])
    assert y[1] == 0.0  # empty row keeps its incoming value


@pytest.mark.parametrize("I,P", WIDTHS)
def test_mul_csc_vec(fixtures, I, P):
    """prod.rs:326-349 mul_csc_vec."""
    k = fixtures["kat_mul_csc_vec"]
    ip, ind, d = mat_arrays(k["mat"], I, P)
    y = np.zeros(5)
    O.mul_acc_mat_vec_csc(ip, ind, d, np.array(k["x"]), y)
    assert np.all(np.abs(y - np.array(k["expected"])) < k["epsilon"])


def test_mul_acc_accumulates(fixtures):
    """prod.rs:121-125: the kernels accumulate into res_vec (F13)."""
    k = fixtures["kat_mul_csr_vec"]
    ip, ind, d = mat_arrays(k["mat"])
    y = np.arange(5, dtype=np.float64)
    O.mul_acc_mat_vec_csr(ip, ind, d, np.array(k["x"]), y)
    assert np.all(np.abs(y - (np.arange(5) + np.array(k["expected"]))) < k["epsilon"])


@pytest.mark.parametrize("I,P", WIDTHS)
def test_mul_csr_dense_rowmaj(fixtures, I, P):
    """prod.rs:503-542 mul_csr_dense_rowmaj (eye, mat1*dense1 exact, mat5*dense2 1e-8)."""
    eye = (np.arange(4, dtype=P), np.arange(3, dtype=I), np.ones(3))
    a = np.eye(3)
    res = np.zeros((3, 3))
    O.csr_mulacc_dense_rowmaj(*eye, a, res)
    assert np.array_equal(res, a)

    ip, ind, d = mat_arrays(fixtures["mat1"], I, P)
    b = np.array(fixtures["mat_dense1"])
    res = np.zeros((5, 5))
    O.csr_mulacc_dense_rowmaj(ip, ind, d, b, res)
    assert np.array_equal(res, np.array(fixtures["kat_mat1_x_dense1"]))

    ip, ind, d = mat_arrays(fixtures["mat5"], I, P)
    b = np.array(fixtures["mat_dense2"])
    res = np.zeros((5, 7))
    O.csr_mulacc_dense_rowmaj(ip, ind, d, b, res)
    k = fixtures["kat_mat5_x_dense2"]
    assert np.all(np.abs(res - np.array(k["expected"])) <= k["epsilon"])


def test_mul_csr_dense_colmaj(fixtures):
    """prod.rs:581-595 mul_csr_dense_colmaj: F-order rhs and out, exact."""
    ip, ind, d = mat_arrays(fixtures["mat1"])
    b = np.asfortranarray(np.array(fixtures["mat_dense1"]))
    res = np.zeros((5, 5), order="F")
    O.csr_mulacc_dense_colmaj(ip, ind, d, b, res)
    assert np.array_equal(res.flatten(order="F"),
                          np.array(fixtures["kat_mat1_x_dense1_colmaj_flat"]))


def test_mul_csc_dense(fixtures):
    """prod.rs:545-578 mul_csc_dense_rowmaj / mul_csc_dense_colmaj."""
    ip, ind, d = mat_arrays(fixtures["mat1_csc"])
    b = np.array(fixtures["mat_dense1"])
    res = np.zeros((5, 5))
    O.csc_mulacc_dense_rowmaj(ip, ind, d, b, res)
    assert np.array_equal(res, np.array(fixtures["kat_mat1_x_dense1"]))
    bf = np.asfortranarray(b)
    res = np.zeros((5, 5), order="F")
    O.csc_mulacc_dense_colmaj(ip, ind, d, bf, res)
    assert np.array_equal(res, np.array(fixtures["kat_mat1_x_dense1"]))


def test_sparse_dot_dense_layout_sweep(fixtures):
    """prod.rs:618-651 test_sparse_dot_dense: 6 sparse x 5 dense layouts (incl.
    transposed / F-order views) vs dense dot, rtol 1e-7 atol 1e-12 (:604-605)."""
    tol = fixtures["assert_close"]

    def t_csr(m):  # transpose_into of a CSR == same arrays read as CSC
        return {"storage": "CSC", "shape": m["shape"][::-1], "indptr": m["indptr"],
                "indices": m["indices"], "data": m["data"]}
    sparse = [fixtures[k] for k in ("mat1", "mat1_csc", "mat2")] + [t_csr(fixtures["mat2"])] + \
             [fixtures["mat4"], fixtures["mat5"]]
    d1, d2 = np.array(fixtures["mat_dense1"]), np.array(fixtures["mat_dense2"])
    dense = [d1, np.asfortranarray(d1), d1.T, d2, d2.T]
    n = 0
    for m in sparse:
        ip, ind, dat = mat_arrays(m)
        rows, cols = m["shape"]
        cls = sp.csr_matrix if m["storage"] == "CSR" else sp.csc_matrix
        truth_mat = cls((dat, ind, ip), shape=(rows, cols)).toarray()
        for dn in dense:
            if dn.shape[0] < cols:
                continue
            dv = dn[:cols, :]
            k = dv.shape[1]
            truth = truth_mat @ dv
            # operator dispatch, csmat.rs:2009: k>=8 rowmaj/C-order else colmaj/F-order
            if k >= 8:
                out = np.zeros((rows, k))
                f = O.csr_mulacc_dense_rowmaj if m["storage"] == "CSR" else O.csc_mulacc_dense_rowmaj
            else:
                out = np.zeros((rows, k), order="F")
                f = O.csr_mulacc_dense_colmaj if m["storage"] == "CSR" else O.csc_mulacc_dense_colmaj
            f(ip, ind, dat, dv, out)
            assert np.all(np.abs(out - truth) <= np.abs(truth) * tol["rtol"] + tol["atol"])
            n += 1
    assert n >= 20


@pytest.mark.parametrize("I,P", WIDTHS)
@pytest.mark.parametrize("threads", [1, 4])
def test_mul_csr_csr(fixtures, I, P, threads):
    """prod.rs:426-437 mul_csr_csr, smmp.rs:468-473, :492-501 (Fixed(4) on 5 rows):
    whole-matrix equality (indptr, indices, data)."""
    a = mat_arrays(fixtures["mat1"], I, P)
    b = mat_arrays(fixtures["mat2"], I, P)
    for rhs, exp in ((a, "mat1_self_matprod"), (b, "mat1_matprod_mat2")):
        cip, cind, cd = O.mul_csr_csr((5, 5), a, (5, 5), rhs, threads=threads)
        e = fixtures[exp]
        assert cip.tolist() == e["indptr"]
        assert cind.tolist() == e["indices"]
        assert cd.tolist() == e["data"]


def test_symbolic_and_numeric(fixtures):
    """smmp.rs:423-465 symbolic_and_numeric: indptr, indices, data separately."""
    a = mat_arrays(fixtures["mat1"])
    b = mat_arrays(fixtures["mat2"])
    cip, cind = O.symbolic(5, 5, a[0], a[1], b[0], b[1])
    cd = O.numeric(5, 5, a, b, cip, cind)
    e = fixtures["mat1_matprod_mat2"]
    assert cip.tolist() == e["indptr"]
    assert cind.tolist() == e["indices"]
    assert cd.tolist() == e["data"]


def test_mul_csc_csc_via_transposes(fixtures):
    """prod.rs:439-446 mul_csc_csc: (CSC,CSC) = mul_csr_csr(B^T, A^T) read back as
    CSC (csmat.rs:1944-1947)."""
    a = mat_arrays(fixtures["mat1_csc"])
    b = mat_arrays(fixtures["mat4"])
    cip, cind, cd = O.mul_csr_csr((5, 5), b, (5, 5), a)
    e = fixtures["mat1_csc_matprod_mat4"]
    assert cip.tolist() == e["indptr"] and cind.tolist() == e["indices"]
    assert cd.tolist() == e["data"]


def test_mul_csc_csr_mixed(fixtures):
    """prod.rs:448-458 mul_csc_csr: CSR*CSC converts rhs (csmat.rs:1935-1938)."""
    a = mat_arrays(fixtures["mat1"])
    ac = mat_arrays(fixtures["mat1_csc"])
    rhs_csr = O.convert_mat_storage(5, 5, *ac)
    assert rhs_csr[0].tolist() == fixtures["mat1"]["indptr"]
    assert rhs_csr[1].tolist() == fixtures["mat1"]["indices"]
    assert rhs_csr[2].tolist() == fixtures["mat1"]["data"]
    cip, cind, cd = O.mul_csr_csr((5, 5), a, (5, 5), rhs_csr)
    e = fixtures["mat1_self_matprod"]
    assert (cip.tolist(), cind.tolist(), cd.tolist()) == (e["indptr"], e["indices"], e["data"])


def test_csr_to_csc(fixtures):
    """csmat.rs:2571 csr_to_csc via raw::convert_mat_storage (csmat.rs:1782-1829)."""
    ip, ind, d = mat_arrays(fixtures["mat1"])
    oip, oind, od = O.convert_mat_storage(5, 5, ip, ind, d)
    e = fixtures["mat1_csc"]
    assert (oip.tolist(), oind.tolist(), od.tolist()) == (e["indptr"], e["indices"], e["data"])


def test_mul_zero_rows_and_issue_99():
    """smmp.rs:476-489 mul_zero_rows (gh#239); csmat.rs:3047-3052 issue_99."""
    e32 = np.zeros(0, dtype=np.uint32)
    a = (np.zeros(1, dtype=np.uint32), e32, np.zeros(0))
    b = (np.zeros(12, dtype=np.uint32), e32, np.zeros(0))
    cip, cind, cd = O.mul_csr_csr((0, 11), a, (11, 11), b)
    assert cip.tolist() == [0] and len(cind) == 0 and len(cd) == 0
    a = (np.zeros(11, dtype=np.uint32), e32, np.zeros(0))
    b = (np.zeros(2, dtype=np.uint32), e32, np.zeros(0))
    cip, cind, cd = O.mul_csr_csr((10, 1), a, (1, 9), b, threads=4)
    assert cip.tolist() == [0] * 11 and len(cind) == 0


def test_csvec_dot_by_binary_search(fixtures):
    """prod.rs:312-323."""
    k = fixtures["kat_csvec_dot"]
    for n1, n2, want in k["expected"]:
        a, b = k[n1], k[n2]
        assert O.csvec_dot_by_binary_search(a["indices"], a["data"], b["indices"], b["data"]) == want
    # random: equals the merge dot (same terms, ascending order), whichever vector is shorter
    rng = np.random.default_rng(3)
    for n1, n2 in ((5, 200), (200, 5), (60, 60), (0, 9)):
        i1, i2 = (np.sort(rng.choice(300, n, replace=False)) for n in (n1, n2))
        d1, d2 = rng.standard_normal(n1), rng.standard_normal(n2)
        oi, od = O.csr_mul_csvec(np.array([0, n1], np.uint64), i1.astype(np.uint64), d1, i2, d2)
        want = od[0] if len(od) else 0.0
        assert O.csvec_dot_by_binary_search(i1, d1, i2, d2) == want


def test_csvec_products(fixtures):
    """prod.rs:461-468 mul_csr_csvec; lib.rs:54-60 README eye(5)*CsVec;
    prod.rs:470-474 zero-dim CsVec handled by the host wrapper."""
    k = fixtures["kat_csvec"]
    ip, ind, d = mat_arrays(fixtures["mat1"])
    oi, od = O.csr_mul_csvec(ip, ind, d, np.array(k["v"]["indices"]), np.array(k["v"]["data"]))
    assert oi.tolist() == k["mat1_times_v"]["indices"]
    assert od.tolist() == k["mat1_times_v"]["data"]
    r = fixtures["kat_readme_eye"]
    eye = (np.arange(6, dtype=np.uint32), np.arange(5, dtype=np.uint32), np.ones(5))
    oi, od = O.csr_mul_csvec(*eye, np.array(r["x"]["indices"]), np.array(r["x"]["data"]))
    assert oi.tolist() == r["x"]["indices"] and od.tolist() == r["x"]["data"]


def test_structural_zeros_kept_in_spgemm():
    """SURVEY F12 / smmp.rs:109-129: numeric cancellation keeps the pattern."""
    # A = [1 1], B = [[1],[-1]] -> C = [[0]] with one STRUCTURAL entry.
    a = (np.array([0, 2], np.uint32), np.array([0, 1], np.uint32), np.array([1., 1.]))
    b = (np.array([0, 1, 2], np.uint32), np.array([0, 0], np.uint32), np.array([1., -1.]))
    cip, cind, cd = O.mul_csr_csr((1, 2), a, (2, 1), b)
    assert cip.tolist() == [0, 1] and cind.tolist() == [0] and cd.tolist() == [0.0]


def test_sliced_indptr(fixtures):
    """indptr.rs:122-124 / slicing.rs:65-89: non-zero-based indptr of a row
    slice works without rebasing (oracle subtracts indptr[0] like sprs)."""
    ip, ind, d = mat_arrays(fixtures["mat1"])
    x = np.arange(1., 6.)
    full = np.zeros(5)
    O.mul_acc_mat_vec_csr(ip, ind, d, x, full)
    part = np.zeros(3)
    s = ip[2]
    O.mul_acc_mat_vec_csr(ip[2:], ind[s:], d[s:], x, part)
    assert np.array_equal(part, full[2:])


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_random_vs_scipy(seed):
    """Independent cross-check: scipy csr @ on random inputs --
    SpGEMM pattern equality, values to 1e-12 relative of sum|terms|."""
    rng = np.random.default_rng(seed)
    n, m, p = 300, 200, 250
    a = rand_csr(rng, n, m, 6, empty_frac=0.1)
    b = rand_csr(rng, m, p, 5, skew=True)
    A = sp.csr_matrix((a[2], a[1], a[0]), shape=(n, m))
    B = sp.csr_matrix((b[2], b[1], b[0]), shape=(m, p))
    for th in (1, 3, 0):
        cip, cind, cd = O.mul_csr_csr((n, m), a, (m, p), b, threads=th)
        Cs = (A @ B)
        Cs.sort_indices()
        assert np.array_equal(cip, Cs.indptr) and np.array_equal(cind, Cs.indices)
        bound = (abs(A) @ abs(B))
        bound.sort_indices()
        assert np.all(np.abs(cd - Cs.data) <= 1e-12 * bound.data + 1e-300)
    x = rng.standard_normal(m)
    y = np.zeros(n)
    O.mul_acc_mat_vec_csr(*a, x, y)
    assert np.allclose(y, A @ x, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("I", [np.uint32, np.uint64])
def test_triplets_to_csr(fixtures, I):
    """triplet.rs:342-453 triplet_incremental / unordered / additions / from_vecs and
    :571-580 triplet_empty_lines: TriMat::to_csr == expected CSC converted to CSR."""
    for name, k in fixtures["kat_triplets"].items():
        ip, ind, d = O.triplets_to_csr(k["shape"], k["rows"], k["cols"], k["data"], I)
        if "expected_csc" in k:
            e = k["expected_csc"]
            eip, eind, ed = mat_arrays(e, I)
            rip, rind, rd = O.convert_mat_storage(e["shape"][1], e["shape"][0], eip, eind, ed)
            assert ip.tolist() == rip.tolist(), name
            assert ind.tolist() == rind.tolist() and d.tolist() == rd.tolist(), name
        else:
            assert ip.tolist() == k["expected_csr_indptr"] and len(ind) == 0 and len(d) == 0


# ---------------------------------------------------------------------------------------
# BiCGSTAB (sprs/src/sparse/linalg/bicgstab.rs), the iterative caller of SpMV
def _kat_bicgstab_csr(fixtures, I=np.uint64):
    k = fixtures["kat_bicgstab"]
    ip, ind, d = mat_arrays(k["a"], I)  # CSC
    return k, O.convert_mat_storage(4, 4, ip, ind, d)


def test_bicgstab_reference_kat(fixtures):
    """bicgstab.rs:356-390 test_bicgstab_f64: Ok within max_iter at tol 1e-60 and
    |1 - b/b_recovered| < tol.  Reaching an exactly zero residual depends on every rounding
    of the iteration, so this pins the restatement's operation order."""
    k, csr = _kat_bicgstab_csr(fixtures)
    ok, s = O.BiCGSTAB.solve(csr, k["x0"], k["b"], k["tol"], k["max_iter"])
    assert ok, "the reference's test unwraps an Ok"
    assert s.iteration_count() <= k["max_iter"]
    assert s.soft_restart_threshold() == k["soft_restart_threshold"]
    b_rec = np.zeros(4)
    O.mul_acc_mat_vec_csr(*csr, s.x(), b_rec)
    assert np.all(np.abs(1.0 - np.array(k["b"]) / b_rec) < k["tol"])
    assert np.allclose(s.x(), k["x_exact"], rtol=1e-15, atol=0)
    assert s.err() < k["tol"] and s.hard_restart_count() >= 1


def test_bicgstab_new_and_restarts(fixtures):
    """bicgstab.rs:120-146 (new), :177-196 (restarts): state after construction and the
    restart counters."""
    k, csr = _kat_bicgstab_csr(fixtures, np.uint32)
    A = np.zeros((4, 4))
    ip, ind, d = csr
    for r in range(4):
        A[r, ind[ip[r]:ip[r + 1]]] = d[ip[r]:ip[r + 1]]
    x0, b = np.array([0.5, -1.0, 2.0, 0.25]), np.array(k["b"])
    s = O.BiCGSTAB(csr, x0, b)
    r0 = b - A @ x0
    assert np.allclose(s.r(), r0, rtol=1e-15) and np.array_equal(s.r(), s.rhat())
    assert np.array_equal(s.r(), s.p()) and np.array_equal(s.x(), x0)
    assert np.array_equal(s.b(), b)
    assert np.isclose(s.err(), np.linalg.norm(r0), rtol=1e-15)
    assert s.rho() == s.err() * s.err()
    assert (s.iteration_count(), s.soft_restart_count(), s.hard_restart_count()) == (0, 0, 0)
    e1 = s.step()
    assert e1 == s.err() and s.iteration_count() == 1
    assert np.isclose(np.linalg.norm(s.r()), e1, rtol=1e-14)
    soft = s.soft_restart_count()
    s.soft_restart()
    assert s.soft_restart_count() == soft + 1 and np.array_equal(s.rhat(), s.r())
    assert np.array_equal(s.p(), s.r()) and s.rho() == s.err() * s.err()
    s.hard_restart()  # counts as a hard restart only (bicgstab.rs:195)
    assert s.hard_restart_count() == 1 and s.soft_restart_count() == soft + 1
    assert np.allclose(s.r(), b - A @ s.x(), rtol=1e-13, atol=1e-15)
    # a threshold of 0 never soft-restarts, a huge one always does
    s0 = O.BiCGSTAB(csr, x0, b).with_restart_threshold(0.0)
    s1 = O.BiCGSTAB(csr, x0, b).with_restart_threshold(1e300)
    for _ in range(3):
        s0.step()
        s1.step()
    assert s0.soft_restart_count() == 0 and s1.soft_restart_count() == 3


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_bicgstab_random_dominant_systems(seed):
    """Diagonally dominant non-symmetric systems: converges to the direct solution, the
    accepted error is the true residual, and an unreachable tolerance returns Err with the
    state (bicgstab.rs:151-175)."""
    rng = np.random.default_rng(1000 + seed)
    n = 300
    A = sp.random(n, n, density=0.03, random_state=rng, format="csr")
    A = (A + sp.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)).tocsr()
    A.sort_indices()
    csr = (A.indptr.astype(np.uint32), A.indices.astype(np.uint32), A.data.copy())
    b = rng.standard_normal(n)
    ok, s = O.BiCGSTAB.solve(csr, np.zeros(n), b, 1e-10, 200)
    assert ok
    x_direct = np.linalg.solve(A.toarray(), b)
    assert np.allclose(s.x(), x_direct, rtol=1e-8, atol=1e-10)
    assert np.isclose(np.linalg.norm(b - A @ s.x()), s.err(), rtol=1e-6, atol=1e-16)
    ok, s = O.BiCGSTAB.solve(csr, np.zeros(n), b, 0.0, 4)
    assert not ok and s.iteration_count() == 4
