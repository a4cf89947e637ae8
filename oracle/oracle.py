"""ctypes front-end of the CPU oracle (oracle/sprs_oracle.cpp).

TEST INFRASTRUCTURE ONLY.  Import this from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs -- never from sprs_b200/.
Function names follow the reference (sprs/src/sparse/prod.rs, smmp.rs).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "sprs_oracle.cpp")
    if force or not os.path.exists(so) or (
            os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(so)):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"],
                              stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            so = build()
        _LIB = C.CDLL(so)
        _LIB.oracle_num_procs.restype = C.c_int
    return _LIB


def _suffix(indptr, indices):
    ib, pb = indices.dtype.itemsize, indptr.dtype.itemsize
    suf = {(4, 4): "44", (8, 8): "88", (4, 8): "48"}.get((ib, pb))
    if suf is None:
        raise TypeError("oracle supports (I,Iptr) byte widths (4,4), (8,8), (4,8)")
    return suf


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _csx(indptr, indices, data):
    indptr = np.ascontiguousarray(indptr)
    indices = np.ascontiguousarray(indices)
    data = np.ascontiguousarray(data, dtype=np.float64)
    assert indptr.dtype.kind in "ui" and indices.dtype.kind in "ui"
    return indptr, indices, data


def num_procs():
    return int(lib().oracle_num_procs())


def mul_acc_mat_vec_csr(indptr, indices, data, x, y):
    """prod.rs:103-127: y += A x (A CSR), in place on y."""
    indptr, indices, data = _csx(indptr, indices, data)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert y.dtype == np.float64 and y.flags.c_contiguous
    f = getattr(lib(), "oracle_mul_acc_mat_vec_csr_" + _suffix(indptr, indices))
    f(C.c_size_t(len(indptr) - 1), _p(indptr), _p(indices), _p(data), _p(x), _p(y))
    return y


def mul_acc_mat_vec_csc(indptr, indices, data, x, y):
    """prod.rs:74-99: y += A x (A CSC), in place on y."""
    indptr, indices, data = _csx(indptr, indices, data)
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert y.dtype == np.float64 and y.flags.c_contiguous
    f = getattr(lib(), "oracle_mul_acc_mat_vec_csc_" + _suffix(indptr, indices))
    f(C.c_size_t(len(indptr) - 1), _p(indptr), _p(indices), _p(data), _p(x), _p(y))
    return y


def _dense(name, indptr, indices, data, rhs, out):
    indptr, indices, data = _csx(indptr, indices, data)
    assert rhs.dtype == np.float64 and out.dtype == np.float64 and rhs.ndim == 2 and out.ndim == 2
    es = 8
    f = getattr(lib(), "oracle_%s_%s" % (name, _suffix(indptr, indices)))
    f(C.c_size_t(len(indptr) - 1), C.c_size_t(rhs.shape[1]), _p(indptr), _p(indices), _p(data),
      _p(rhs), C.c_ssize_t(rhs.strides[0] // es), C.c_ssize_t(rhs.strides[1] // es),
      _p(out), C.c_ssize_t(out.strides[0] // es), C.c_ssize_t(out.strides[1] // es))
    return out


def csr_mulacc_dense_colmaj(indptr, indices, data, rhs, out):
    """prod.rs:274-298 (any-stride views, as ndarray allows)."""
    return _dense("csr_mulacc_dense_colmaj", indptr, indices, data, rhs, out)


def csr_mulacc_dense_rowmaj(indptr, indices, data, rhs, out):
    """prod.rs:189-214."""
    return _dense("csr_mulacc_dense_rowmaj", indptr, indices, data, rhs, out)


def csc_mulacc_dense_colmaj(indptr, indices, data, rhs, out):
    """prod.rs:246-269."""
    return _dense("csc_mulacc_dense_colmaj", indptr, indices, data, rhs, out)


def csc_mulacc_dense_rowmaj(indptr, indices, data, rhs, out):
    """prod.rs:219-241."""
    return _dense("csc_mulacc_dense_rowmaj", indptr, indices, data, rhs, out)


def csr_mul_csvec(indptr, indices, data, v_indices, v_data):
    """prod.rs:162-184: returns (indices, data) of the sparse result."""
    indptr, indices, data = _csx(indptr, indices, data)
    v_indices = np.ascontiguousarray(v_indices, dtype=indices.dtype)
    v_data = np.ascontiguousarray(v_data, dtype=np.float64)
    rows = len(indptr) - 1
    oi = np.empty(rows, dtype=indices.dtype)
    od = np.empty(rows, dtype=np.float64)
    f = getattr(lib(), "oracle_csr_mul_csvec_" + _suffix(indptr, indices))
    f.restype = C.c_size_t
    n = f(C.c_size_t(rows), _p(indptr), _p(indices), _p(data), C.c_size_t(len(v_indices)),
          _p(v_indices), _p(v_data), _p(oi), _p(od))
    return oi[:n].copy(), od[:n].copy()


def csvec_dot_by_binary_search(i1, d1, i2, d2):
    """prod.rs:13-72: dot product of two sparse vectors (indices ascending)."""
    i1 = np.ascontiguousarray(i1, dtype=np.uint64)
    i2 = np.ascontiguousarray(i2, dtype=np.uint64)
    d1 = np.ascontiguousarray(d1, dtype=np.float64)
    d2 = np.ascontiguousarray(d2, dtype=np.float64)
    f = getattr(lib(), "oracle_csvec_dot_by_binary_search_" + _suffix(i1, i1))
    f.restype = C.c_double
    return float(f(C.c_size_t(len(i1)), _p(i1), _p(d1), C.c_size_t(len(i2)), _p(i2), _p(d2)))


def convert_mat_storage(outer, inner, indptr, indices, data):
    """csmat.rs:1782-1829: CSR<->CSC; returns (indptr, indices, data)."""
    indptr, indices, data = _csx(indptr, indices, data)
    oip = np.zeros(inner + 1, dtype=indptr.dtype)
    oind = np.empty_like(indices)
    od = np.empty_like(data)
    f = getattr(lib(), "oracle_convert_mat_storage_" + _suffix(indptr, indices))
    f(C.c_size_t(outer), C.c_size_t(inner), _p(indptr), _p(indices), _p(data), _p(oip),
      _p(oind), _p(od))
    return oip, oind, od


def triplets_to_csr(shape, row_inds, col_inds, data, index_dtype=np.uint32):
    """triplet_iter.rs:127-224 TriMat::to_csr: returns (indptr, indices, data)."""
    ri = np.ascontiguousarray(row_inds, dtype=index_dtype)
    ci = np.ascontiguousarray(col_inds, dtype=index_dtype)
    v = np.ascontiguousarray(data, dtype=np.float64)
    n = len(ri)
    ip = np.zeros(shape[0] + 1, dtype=index_dtype)
    ind = np.empty(max(n, 1), dtype=index_dtype)
    d = np.empty(max(n, 1), dtype=np.float64)
    f = getattr(lib(), "oracle_triplets_to_csr_" + _suffix(ip, ind))
    f.restype = C.c_size_t
    k = f(C.c_size_t(shape[0]), C.c_size_t(n), _p(ri), _p(ci), _p(v), _p(ip), _p(ind), _p(d))
    return ip, ind[:k].copy(), d[:k].copy()


def mul_csr_csr(a_shape, a, b_shape, b, threads=0):
    """smmp.rs:196-416.  a, b = (indptr, indices, data); threads=0 -> the
    reference's Automatic rule, n -> Fixed(n).  Returns (indptr, indices, data)."""
    aip, aind, ad = _csx(*a)
    bip, bind, bd = _csx(*b)
    assert a_shape[1] == b_shape[0]
    assert aip.dtype == bip.dtype and aind.dtype == bind.dtype
    suf = _suffix(aip, aind)
    L = lib()
    f = getattr(L, "oracle_mul_csr_csr_" + suf)
    f.restype = C.c_void_p
    h = f(C.c_size_t(a_shape[0]), C.c_size_t(a_shape[1]), C.c_size_t(b_shape[1]), _p(aip),
          _p(aind), _p(ad), _p(bip), _p(bind), _p(bd), C.c_size_t(threads))
    h = C.c_void_p(h)
    nnzf = getattr(L, "oracle_spgemm_nnz_" + suf)
    nnzf.restype = C.c_size_t
    nnz = nnzf(h)
    cip = np.empty(a_shape[0] + 1, dtype=aip.dtype)
    cind = np.empty(nnz, dtype=aind.dtype)
    cd = np.empty(nnz, dtype=np.float64)
    getattr(L, "oracle_spgemm_fetch_" + suf)(h, _p(cip), _p(cind), _p(cd))
    getattr(L, "oracle_spgemm_free_" + suf)(h)
    return cip, cind, cd


def symbolic(a_rows, b_cols, a_indptr, a_indices, b_indptr, b_indices):
    """smmp.rs:81-131.  Returns (c_indptr, c_indices)."""
    a_indptr = np.ascontiguousarray(a_indptr)
    a_indices = np.ascontiguousarray(a_indices)
    b_indptr = np.ascontiguousarray(b_indptr, dtype=a_indptr.dtype)
    b_indices = np.ascontiguousarray(b_indices, dtype=a_indices.dtype)
    suf = _suffix(a_indptr, a_indices)
    f = getattr(lib(), "oracle_symbolic_" + suf)
    f.restype = C.c_size_t
    cip = np.zeros(a_rows + 1, dtype=a_indptr.dtype)
    cap = 0
    cind = np.empty(0, dtype=a_indices.dtype)
    n = f(C.c_size_t(a_rows), C.c_size_t(b_cols), _p(a_indptr), _p(a_indices), _p(b_indptr),
          _p(b_indices), _p(cip), _p(cind), C.c_size_t(cap))
    cind = np.empty(n, dtype=a_indices.dtype)
    f(C.c_size_t(a_rows), C.c_size_t(b_cols), _p(a_indptr), _p(a_indices), _p(b_indptr),
      _p(b_indices), _p(cip), _p(cind), C.c_size_t(n))
    return cip, cind


def numeric(a_rows, b_cols, a, b, c_indptr, c_indices):
    """smmp.rs:151-189.  Returns c_data."""
    aip, aind, ad = _csx(*a)
    bip, bind, bd = _csx(*b)
    c_indptr = np.ascontiguousarray(c_indptr, dtype=aip.dtype)
    c_indices = np.ascontiguousarray(c_indices, dtype=aind.dtype)
    cd = np.zeros(len(c_indices), dtype=np.float64)
    f = getattr(lib(), "oracle_numeric_" + _suffix(aip, aind))
    f(C.c_size_t(a_rows), C.c_size_t(b_cols), _p(aip), _p(aind), _p(ad), _p(bip), _p(bind),
      _p(bd), _p(c_indptr), _p(c_indices), _p(cd))
    return cd


class BiCGSTAB:
    """linalg/bicgstab.rs:95-300 on dense f64 vectors (see the restatement's header comment).
    The matrix is CSR (indptr, indices, data); a CSC matrix goes through
    convert_mat_storage first -- A*v sums in ascending column order either way."""

    def __init__(self, csr, x0, b):
        self._csr = _csx(*csr)  # borrowed by the C side: keep alive
        ip, ind, d = self._csr
        self.n = len(ip) - 1
        x0 = np.ascontiguousarray(x0, dtype=np.float64)
        b = np.ascontiguousarray(b, dtype=np.float64)
        assert x0.shape == (self.n,) and b.shape == (self.n,)
        L = lib()
        f = getattr(L, "oracle_bicgstab_new_" + _suffix(ip, ind))
        f.restype = C.c_void_p
        self._h = C.c_void_p(f(C.c_size_t(self.n), _p(ip), _p(ind), _p(d), _p(x0), _p(b)))
        L.oracle_bicgstab_step.restype = C.c_double
        L.oracle_bicgstab_solve.restype = C.c_int

    def __del__(self):
        if getattr(self, "_h", None):
            lib().oracle_bicgstab_free(self._h)
            self._h = None

    @classmethod
    def solve(cls, csr, x0, b, tol, max_iter):
        """Returns (ok, solver): Ok(solver) / Err(solver) of bicgstab.rs:151-175."""
        s = cls(csr, x0, b)
        ok = lib().oracle_bicgstab_solve(s._h, C.c_double(tol), C.c_size_t(max_iter))
        return bool(ok), s

    def step(self):
        return float(lib().oracle_bicgstab_step(self._h))

    def soft_restart(self):
        lib().oracle_bicgstab_soft_restart(self._h)

    def hard_restart(self):
        lib().oracle_bicgstab_hard_restart(self._h)

    def with_restart_threshold(self, thresh):
        lib().oracle_bicgstab_set_threshold(self._h, C.c_double(thresh))
        return self

    def _vec(self, which):
        out = np.empty(self.n, dtype=np.float64)
        lib().oracle_bicgstab_get(self._h, C.c_int(which), _p(out))
        return out

    def _stats(self):
        counts = (C.c_size_t * 3)()
        scal = (C.c_double * 3)()
        lib().oracle_bicgstab_stats(self._h, counts, scal)
        return list(counts), list(scal)

    def x(self):
        return self._vec(0)

    def r(self):
        return self._vec(1)

    def rhat(self):
        return self._vec(2)

    def p(self):
        return self._vec(3)

    def b(self):
        return self._vec(4)

    def iteration_count(self):
        return self._stats()[0][0]

    def soft_restart_count(self):
        return self._stats()[0][1]

    def hard_restart_count(self):
        return self._stats()[0][2]

    def err(self):
        return self._stats()[1][0]

    def rho(self):
        return self._stats()[1][1]

    def soft_restart_threshold(self):
        return self._stats()[1][2]


def ext_spmv_csr_omp(indptr, indices, data, x, y, threads):
    """EXTENSION (not in the reference): all-cores row-chunked SpMV, u32 only."""
    indptr, indices, data = _csx(indptr, indices, data)
    assert indptr.dtype.itemsize == 4 and indices.dtype.itemsize == 4
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().oracle_ext_spmv_csr_omp_44(C.c_size_t(len(indptr) - 1), _p(indptr), _p(indices),
                                     _p(data), _p(x), _p(y), C.c_int(threads))
    return y
