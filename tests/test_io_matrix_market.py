"""CPU tests of the MatrixMarket reader / writer (sprs_b200/io.py), replaying the reference's
io.rs tests on the reference's own data files (tests/golden/matrix_market/)."""
import os

import numpy as np
import pytest

from sprs_b200 import io as mm

DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "matrix_market")


def test_simple_matrix_market_read():
    """io.rs:476-490 simple_matrix_market_read / :605-622 ..._from_bufread."""
    for mat in (mm.read_matrix_market(os.path.join(DATA, "simple.mm")),
                mm.read_matrix_market_from_bufread(open(os.path.join(DATA, "simple.mm")))):
        assert mat.rows() == 5 and mat.cols() == 5 and mat.nnz() == 8
        assert mat.row_inds == [0, 1, 2, 0, 3, 3, 3, 4]
        assert mat.col_inds == [0, 1, 2, 3, 1, 3, 4, 4]
        assert mat.data == [1., 10.5, 1.5e-02, 6., 2.505e2, -2.8e2, 3.332e1, 1.2e+1]


def test_int_matrix_market_read():
    """io.rs:625-636 int_matrix_market_read."""
    mat = mm.read_matrix_market(os.path.join(DATA, "simple_int.mm"), dtype=np.int64)
    assert (mat.rows(), mat.cols(), mat.nnz()) == (5, 5, 8)
    assert mat.row_inds == [0, 1, 2, 0, 3, 3, 3, 4] and mat.col_inds == [0, 1, 2, 3, 1, 3, 4, 4]
    assert mat.data == [1, 1, 1, 6, 2, -2, 3, 1]


def test_failing_matrix_market_reads():
    """io.rs:492-533 failing_matrix_market_reads (the f64 / i64 arms)."""
    cplx, flt, integer = (os.path.join(DATA, p) for p in
                          ("complex/simple.mtx", "simple.mm", "simple_int.mm"))
    mm.read_matrix_market(integer, dtype=np.int64)
    mm.read_matrix_market(flt)
    for path, dt in ((cplx, np.float64), (cplx, np.int64), (flt, np.int64), (integer, np.float64)):
        with pytest.raises(mm.IoError):
            mm.read_matrix_market(path, dtype=dt)
    with pytest.raises(mm.IoError) as e:
        mm.read_matrix_market(cplx)
    assert str(e.value) == "Tried to load complex file into real matrix."
    with pytest.raises(mm.IoError) as e:
        mm.read_matrix_market(os.path.join(DATA, "complex/hermitian-int.mtx"), dtype=np.int64)
    assert e.value == mm.IoError(mm.IoError.UNSUPPORTED)
    with pytest.raises(mm.IoError):  # pattern file into a real matrix
        mm.read_matrix_market(os.path.join(DATA, "pattern.mm"))


def test_bad_files():
    """io.rs:638-652 matrix_market_read_fail_too_many_in_entry / ..._not_enough_entries."""
    for f in ("bad_files/too_many_elems_in_entry.mm", "bad_files/not_enough_entries.mm"):
        with pytest.raises(mm.IoError) as e:
            mm.read_matrix_market(os.path.join(DATA, f))
        assert e.value == mm.IoError(mm.IoError.BAD_FILE)


def test_read_write_read(tmp_path):
    """io.rs:654-667 read_write_read_matrix_market."""
    mat = mm.read_matrix_market(os.path.join(DATA, "simple.mm"))
    p = tmp_path / "simple.mm"
    mm.write_matrix_market(p, mat)
    mat2 = mm.read_matrix_market(p)
    assert mat == mat2
    mm.write_matrix_market(p, mat2)
    assert mm.read_matrix_market(p) == mat
    assert open(p).read().startswith("%%MatrixMarket matrix coordinate real general\n% written by sprs\n5 5 8\n")


def test_symmetric_expansion(tmp_path):
    """io.rs:682-701 read_symmetric_matrix_market: off-diagonal entries are mirrored; the
    symmetric writer keeps one triangle and patches the entry count."""
    mat = mm.read_matrix_market(os.path.join(DATA, "symmetric.mm"))
    assert mat.nnz() == 8
    assert mat.row_inds == [0, 1, 2, 3, 1, 4, 3, 4] and mat.col_inds == [0, 1, 2, 1, 3, 3, 4, 4]
    assert mat.data == [1., 10.5, 1.5e-2, 2.505e2, 2.505e2, 3.332e1, 3.332e1, 1.2e1]
    p = tmp_path / "symmetric.mm"
    mm.write_matrix_market_sym(p, mat, mm.SYMMETRIC)
    txt = open(p).read().splitlines()
    assert txt[0] == "%%MatrixMarket matrix coordinate real symmetric" and txt[2].split() == ["5", "5", "6"]
    assert len(txt[2]) == len("5 5 8")
    mat2 = mm.read_matrix_market(p)
    key = lambda m: sorted(zip(m.row_inds, m.col_inds, m.data))
    assert key(mat2) == key(mat)


def test_skew_symmetric(tmp_path):
    """io.rs:785-800 skew_symmetric_matrix_market: mirrored entries are negated, diagonal
    entries are an error."""
    p = tmp_path / "skew.mm"
    p.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 2\n2 1 4.5\n3 2 -1\n")
    mat = mm.read_matrix_market(p)
    assert list(zip(mat.row_inds, mat.col_inds, mat.data)) == [(1, 0, 4.5), (0, 1, -4.5),
                                                               (2, 1, -1.0), (1, 2, 1.0)]
    p.write_text("%%MatrixMarket matrix coordinate real skew-symmetric\n3 3 1\n2 2 4.5\n")
    with pytest.raises(mm.IoError):
        mm.read_matrix_market(p)


def test_real_hermitian_is_unsupported(tmp_path):
    """num_matrixmarket.rs:158-177: mm_conj() is None for f64 too, so a real hermitian file
    with an off-diagonal entry is UnsupportedMatrixMarketFormat (io.rs:236-241)."""
    p = tmp_path / "h.mm"
    p.write_text("%%MatrixMarket matrix coordinate real hermitian\n2 2 1\n2 1 3.0\n")
    with pytest.raises(mm.IoError) as e:
        mm.read_matrix_market(p)
    assert e.value == mm.IoError(mm.IoError.UNSUPPORTED)


def test_header_and_index_errors(tmp_path):
    p = tmp_path / "x.mm"
    for body in ("%%MatrixMarket matrix array real general\n1 1\n1.0\n",      # not coordinate
                 "%%MatrixMarket matrix coordinate real general\n2 2\n",      # size line short
                 "%%MatrixMarket matrix coordinate real general\n2 2 1\n0 1 1.0\n",   # 0-based
                 "%%MatrixMarket matrix coordinate real general\n2 2 1\n1 x 1.0\n",
                 "%%MatrixMarket matrix coordinate real general\n\n2 2 0\n",  # blank line = size line
                 "%%MatrixMarket matrix coordinate real weird\n2 2 0\n"):
        p.write_text(body)
        with pytest.raises(mm.IoError) as e:
            mm.read_matrix_market(p)
        assert e.value == mm.IoError(mm.IoError.BAD_FILE)
    p.write_text("%%MATRIXMARKET MATRIX COORDINATE REAL GENERAL\n% c\n 2   2  1 \n\n  +2 1  1e1\n")
    mat = mm.read_matrix_market(p)  # tags are case-insensitive, blank lines between entries ok
    assert (mat.shape, mat.row_inds, mat.col_inds, mat.data) == ((2, 2), [1], [0], [10.0])
