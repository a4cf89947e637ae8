"""BiCGSTAB with device-resident vectors (csrc/solver.cu) against the oracle's restatement of
sprs/src/sparse/linalg/bicgstab.rs and the reference's own test.  Written after the
round's last GPU session: the file sorts last so that a surprise here cannot hide the
validated suites under `pytest -x`."""
import os
import subprocess

import numpy as np
import pytest
import scipy.sparse as sparse

from conftest import mat_arrays

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def sp():
    import sprs_b200
    sprs_b200.Context.default()
    return sprs_b200


@pytest.fixture(scope="module")
def O():
    from oracle import oracle
    return oracle


def dominant_system(n, per_row, seed):
    """Non-symmetric, strictly diagonally dominant CSR (u32) and a right-hand side."""
    rng = np.random.default_rng(seed)
    A = sparse.random(n, n, density=per_row / n, random_state=rng, format="csr")
    A = (A + sparse.diags(np.asarray(abs(A).sum(axis=1)).ravel() + 1.0)).tocsr()
    A.sort_indices()
    csr = (A.indptr.astype(np.uint32), A.indices.astype(np.uint32), A.data.copy())
    return A, csr, rng.standard_normal(n)


def test_bicgstab_reference_kat(sp, fixtures, O):
    """bicgstab.rs:356-390 test_bicgstab_f64, as written there: CSC matrix, tol 1e-60."""
    k = fixtures["kat_bicgstab"]
    a = sp.CsMat.new_csc((4, 4), *mat_arrays(k["a"], np.uint64))
    b = sp.CsVec(4, [0, 1, 2, 3], k["b"])
    x0 = sp.CsVec(4, [0, 1, 2, 3], k["x0"])
    res = sp.linalg.bicgstab.BiCGSTAB.solve(a, x0, b, k["tol"], k["max_iter"])  # .unwrap()
    b_recovered = a * res.x()
    print("Iteration count", res.iteration_count())
    print("Soft restart count", res.soft_restart_count())
    print("Hard restart count", res.hard_restart_count())
    for inp, out in zip(b.to_dense(), b_recovered):
        assert abs(1.0 - inp / out) < k["tol"], "Solved output did not match input"
    # with n <= 4 every device sum is the reference's sequential sum: same trajectory
    ip, ind, d = mat_arrays(k["a"], np.uint64)
    ok, ref = O.BiCGSTAB.solve(O.convert_mat_storage(4, 4, ip, ind, d), k["x0"], k["b"],
                               k["tol"], k["max_iter"])
    assert ok
    assert res.iteration_count() == ref.iteration_count()
    assert res.hard_restart_count() == ref.hard_restart_count()
    assert res.soft_restart_count() == ref.soft_restart_count()
    assert res.x().tolist() == ref.x().tolist()


def test_bicgstab_state_after_new_and_steps(sp, O):
    """new / step / soft_restart / hard_restart against the oracle, a few steps of a
    well-conditioned system (sums differ only in their order: tolerance 1e-9)."""
    n = 5000
    A, csr, b = dominant_system(n, 8, 11)
    x0 = np.linspace(-1.0, 1.0, n)
    a = sp.CsMat((n, n), *csr)
    dev = sp.linalg.BiCGSTAB.new(a, x0, b)
    ref = O.BiCGSTAB(csr, x0, b)
    close = dict(rtol=1e-9, atol=1e-12)
    assert np.array_equal(dev.x(), x0) and np.array_equal(dev.b(), b)
    assert np.allclose(dev.r(), ref.r(), **close)
    assert np.array_equal(dev.r(), dev.rhat()) and np.array_equal(dev.r(), dev.p())
    assert np.isclose(dev.err(), ref.err(), rtol=1e-12) and dev.rho() == dev.err() * dev.err()
    assert (dev.iteration_count(), dev.soft_restart_count(), dev.hard_restart_count()) == (0, 0, 0)
    assert dev.soft_restart_threshold() == 0.1
    for it in range(1, 5):
        e_dev, e_ref = dev.step(), ref.step()
        assert np.isclose(e_dev, e_ref, rtol=1e-8), it
        assert e_dev == dev.err() and dev.iteration_count() == it
        assert np.allclose(dev.x(), ref.x(), **close), it
        assert np.allclose(dev.r(), ref.r(), rtol=1e-7, atol=1e-12), it
        assert np.allclose(dev.p(), ref.p(), rtol=1e-7, atol=1e-12), it
        assert np.isclose(dev.rho(), ref.rho(), rtol=1e-7, atol=1e-20), it
        assert dev.soft_restart_count() == ref.soft_restart_count()
        # the running estimate is the norm of the stored residual
        assert np.isclose(np.linalg.norm(dev.r()), e_dev, rtol=1e-12)
    soft = dev.soft_restart_count()
    dev.soft_restart()
    assert dev.soft_restart_count() == soft + 1 and dev.rho() == dev.err() * dev.err()
    assert np.array_equal(dev.rhat(), dev.r()) and np.array_equal(dev.p(), dev.r())
    dev.hard_restart()
    assert dev.hard_restart_count() == 1 and dev.soft_restart_count() == soft + 1
    true_r = b - A @ dev.x()
    assert np.allclose(dev.r(), true_r, rtol=1e-9, atol=1e-13)
    assert np.isclose(dev.err(), np.linalg.norm(true_r), rtol=1e-9)
    assert np.array_equal(dev.rhat(), dev.r()) and np.array_equal(dev.p(), dev.r())


@pytest.mark.parametrize("storage", ["CSR", "CSC"])
def test_bicgstab_solve_dominant(sp, O, storage):
    """solve(): Ok, x equals the oracle's and the direct solution, and the accepted error
    is the true residual norm (hard restart before returning, bicgstab.rs:162-169)."""
    n = 40000
    A, csr, b = dominant_system(n, 10, 23)
    if storage == "CSR":
        a = sp.CsMat((n, n), *csr)
    else:
        Ac = A.tocsc()
        Ac.sort_indices()
        a = sp.CsMat.new_csc((n, n), Ac.indptr.astype(np.uint32), Ac.indices.astype(np.uint32),
                             Ac.data.copy())
    tol = 1e-9
    res = sp.linalg.BiCGSTAB.solve(a, np.zeros(n), b, tol, 200)
    ok, ref = O.BiCGSTAB.solve(csr, np.zeros(n), b, tol, 200)
    assert ok
    x = res.x()
    assert np.allclose(x, ref.x(), rtol=1e-7, atol=1e-10)
    true_err = np.linalg.norm(b - A @ x)
    assert true_err < tol * 1.001
    assert np.isclose(res.err(), true_err, rtol=1e-5, atol=1e-16)
    assert res.hard_restart_count() >= 1
    assert abs(res.iteration_count() - ref.iteration_count()) <= 2
    # the solution vector can be consumed on the device without a copy
    assert res.device_vector("x") != 0


def test_bicgstab_err_and_contracts(sp):
    n = 2000
    A, csr, b = dominant_system(n, 6, 5)
    a = sp.CsMat((n, n), *csr)
    with pytest.raises(sp.linalg.NotConverged) as ei:  # Err(solver), bicgstab.rs:173-174
        sp.linalg.BiCGSTAB.solve(a, np.zeros(n), b, 0.0, 3)
    assert ei.value.solver.iteration_count() == 3
    assert np.isfinite(ei.value.solver.err())
    with pytest.raises(sp.SprsPanic, match="Dimension mismatch"):
        sp.linalg.BiCGSTAB(a, np.zeros(n - 1), b)
    with pytest.raises(sp.SprsPanic, match="Dimension mismatch"):
        sp.linalg.BiCGSTAB(a, np.zeros(n), b[:-1])
    rect = sp.CsMat((3, 4), np.array([0, 1, 2, 3], np.uint32), np.array([0, 1, 2], np.uint32),
                    np.ones(3))
    with pytest.raises(sp.SprsPanic, match="Dimension mismatch"):
        sp.linalg.BiCGSTAB(rect, np.zeros(4), np.zeros(3))
    # restart threshold: 0 never soft-restarts, a huge one restarts every step
    s0 = sp.linalg.BiCGSTAB(a, np.zeros(n), b).with_restart_threshold(0.0)
    s1 = sp.linalg.BiCGSTAB(a, np.zeros(n), b).with_restart_threshold(1e300)
    for _ in range(3):
        s0.step()
        s1.step()
    assert s0.soft_restart_count() == 0 and s1.soft_restart_count() == 3
    assert s1.soft_restart_threshold() == 1e300
    # a DeviceCsMat (matrix already resident) is accepted as well
    s2 = sp.linalg.BiCGSTAB(a.device(), np.zeros(n), b)
    assert s2.step() == s0.__class__(a, np.zeros(n), b).step()  # repeatable to the last bit


def test_bicgstab_operator_form_matches_matrix_form(sp):
    """sprs_b200_bicgstab_new_op: y = A x through a caller-supplied operator (here the same
    device SpMV, enqueued on the solver's stream) -- the same bits as the matrix form, step by
    step.  This is the hook the row-partitioned solver hangs its SpMV + all-gather on."""
    import ctypes as C
    n = 3000
    A, csr, b = dominant_system(n, 9, 31)
    a = sp.CsMat((n, n), *csr)
    ctx, mirror = a.context(), a.device()
    calls = []

    def matvec(d_x, d_y, stream):
        calls.append(stream)
        ctx.check(ctx.lib.sprs_b200_spmv_dev(ctx.h, mirror.h, C.c_void_p(d_x), C.c_void_p(d_y), 0,
                                             C.c_void_p(stream)))

    s_op = sp.linalg.BiCGSTAB.with_operator(ctx, n, matvec, np.zeros(n), b)
    s_mat = sp.linalg.BiCGSTAB(a, np.zeros(n), b)
    assert len(calls) == 1 and s_op.err() == s_mat.err()
    for _ in range(5):
        assert s_op.step() == s_mat.step()
    assert np.array_equal(s_op.x(), s_mat.x()) and np.array_equal(s_op.p(), s_mat.p())
    assert len(calls) == 11
    s_op.run(1e-9, 100)
    s_mat.run(1e-9, 100)
    assert s_op.iteration_count() == s_mat.iteration_count()
    assert np.array_equal(s_op.x(), s_mat.x())
    with pytest.raises(ZeroDivisionError):       # the operator's exception, not a status code
        sp.linalg.BiCGSTAB.with_operator(ctx, n, lambda *a: 1 // 0, np.zeros(n), b)


def test_cpp_bicgstab():
    exe = os.path.join(ROOT, "tests", "cpp", "test_bicgstab")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.rstrip().splitlines()[-1].startswith("OK ")
