"""Writes tests/golden/sprs_fixtures.json.

The numbers below are the reference's own known-answer data, transcribed as
DATA (not code) from the reference's tests; each entry cites where it lives:

  sprs/src/test_data.rs:6-123        mat1..mat5, expected products, dense mats
  sprs/src/sparse/prod.rs:326-423    SpMV KATs (CSC :326-373, CSR :376-423)
  sprs/src/sparse/prod.rs:461-500    CsMat * CsVec KATs
  sprs/src/sparse/prod.rs:503-595    SpMM KATs
  sprs/src/sparse/smmp.rs:476-513    zero-row / empty edge cases
  sprs/src/sparse/csmat.rs:3047-3052 issue_99 (10x1 * 1x9)
  sprs/src/lib.rs:54-73              README example eye(5) * CsVec
  sprs/src/sparse/triplet.rs:342-453, 571-580   TriMat -> CSR/CSC KATs

The reference is Rust and cannot run in this image (no cargo), so these
fixtures -- not a live run of the reference -- are what pins the oracle.
scipy is used below only as an independent cross-check of the transcription
(every expected product is recomputed and compared before the file is written).

Run:  python tests/golden/make_fixtures.py
"""
import json
import os

import numpy as np
import scipy.sparse as sp

F = {}


def csmat(storage, shape, indptr, indices, data):
    return {"storage": storage, "shape": list(shape), "indptr": indptr,
            "indices": indices, "data": data}


# ---- sprs/src/test_data.rs:6-61
F["mat1"] = csmat("CSR", (5, 5), [0, 2, 4, 5, 6, 7], [2, 3, 3, 4, 2, 1, 3],
                  [3., 4., 2., 5., 5., 8., 7.])
F["mat1_csc"] = csmat("CSC", (5, 5), [0, 0, 1, 3, 6, 7], [3, 0, 2, 0, 1, 4, 1],
                      [8., 3., 5., 4., 2., 7., 5.])
F["mat2"] = csmat("CSR", (5, 5), [0, 4, 6, 6, 8, 10], [0, 1, 2, 4, 0, 3, 2, 3, 1, 2],
                  [6., 7., 3., 3., 8., 9., 2., 4., 4., 4.])
F["mat3"] = csmat("CSR", (5, 4), [0, 2, 4, 5, 6, 7], [2, 3, 2, 3, 2, 1, 3],
                  [3., 4., 2., 5., 5., 8., 7.])
F["mat4"] = csmat("CSC", (5, 5), [0, 4, 6, 6, 8, 10], [0, 1, 2, 4, 0, 3, 2, 3, 1, 2],
                  [6., 7., 3., 3., 8., 9., 2., 4., 4., 4.])
F["mat5"] = csmat("CSR", (5, 15), [0, 5, 11, 14, 20, 22],
                  [1, 2, 6, 7, 13, 3, 4, 6, 8, 13, 14, 7, 11, 13, 3, 8, 9, 10, 11, 14, 4, 12],
                  [4.8, 2., 3.7, 5.9, 6., 1.6, 0.3, 9.2, 9.9, 4.8, 6.1, 4.4, 6., 0.1, 7.2,
                   1., 1.4, 6.4, 2.8, 3.4, 5.5, 3.5])
# ---- sprs/src/test_data.rs:63-84 expected sparse products
F["mat1_self_matprod"] = csmat("CSR", (5, 5), [0, 2, 4, 5, 7, 8], [1, 2, 1, 3, 2, 3, 4, 1],
                               [32., 15., 16., 35., 25., 16., 40., 56.])
F["mat1_matprod_mat2"] = csmat("CSR", (5, 5), [0, 2, 5, 5, 7, 9], [2, 3, 1, 2, 3, 0, 3, 2, 3],
                               [8., 16., 20., 24., 8., 64., 72., 14., 28.])
F["mat1_csc_matprod_mat4"] = csmat(
    "CSC", (5, 5), [0, 4, 7, 7, 11, 14], [0, 1, 2, 3, 0, 1, 4, 0, 1, 2, 4, 0, 2, 3],
    [9., 15., 15., 56., 36., 18., 63., 22., 8., 10., 28., 12., 20., 32.])
# ---- sprs/src/test_data.rs:86-123 dense matrices (row-major listing)
F["mat_dense1"] = [[0., 1., 2., 3., 4.], [5., 6., 5., 4., 3.], [4., 5., 4., 3., 2.],
                   [3., 4., 3., 2., 1.], [1., 2., 1., 1., 0.]]
F["mat_dense2"] = [
    [8.2, 1.8, 0.9, 2.6, 6.7, 7.6, 8.3], [8.7, 9.4, 2.6, 6.4, 3.5, 1.2, 4.7],
    [5.3, 9., 8.7, 9.8, 4.6, 2.5, 4.6], [4.7, 6.2, 3.7, 5.6, 4.7, 8.3, 3.],
    [3.5, 6.4, 2.3, 7.3, 4.2, 3.3, 8.9], [3.6, 6.2, 7.3, 3.1, 1.5, 4.1, 0.8],
    [8.8, 8.7, 1.6, 6.1, 5.6, 0.1, 8.5], [4.8, 4.1, 8.1, 0., 0.4, 3., 5.1],
    [6.6, 3.4, 1.7, 3.9, 2.2, 5.5, 6.8], [4.8, 3.7, 9.2, 7.4, 3.5, 1.5, 5.8],
    [4.3, 6.9, 6.5, 5.7, 7.6, 9.5, 5.8], [5.7, 6.9, 8.5, 0.1, 5.8, 9.6, 4.9],
    [6.9, 5.4, 0., 1.2, 4.8, 1.5, 7.9], [2.8, 5.1, 0.6, 3., 8.4, 8.6, 1.],
    [8.1, 1.9, 6.3, 0.2, 0.3, 5.9, 0.]]

# ---- sprs/src/sparse/prod.rs:376-398 mul_csr_vec (tol 1e-7 in the reference)
F["kat_mul_csr_vec"] = {
    "mat": csmat("CSR", (5, 5), [0, 3, 3, 5, 6, 7], [1, 2, 3, 2, 3, 4, 4],
                 [0.75672424, 0.1649078, 0.30140296, 0.10358244, 0.6283315, 0.39244208,
                  0.57202407]),
    "x": [0.1, 0.2, -0.1, 0.3, 0.9],
    "expected": [0.22527496, 0., 0.17814121, 0.35319787, 0.51482166],
    "epsilon": 1e-7}
# ---- sprs/src/sparse/prod.rs:326-349 mul_csc_vec
F["kat_mul_csc_vec"] = {
    "mat": csmat("CSC", (5, 5), [0, 2, 4, 5, 6, 7], [2, 3, 3, 4, 2, 1, 3],
                 [0.35310881, 0.42380633, 0.28035896, 0.58082095, 0.53350123, 0.88132896,
                  0.72527863]),
    "x": [0.1, 0.2, -0.1, 0.3, 0.9],
    "expected": [0., 0.26439869, -0.01803924, 0.75120319, 0.11616419],
    "epsilon": 1e-7}
# ---- sprs/src/sparse/prod.rs:503-542 mul_csr_dense_rowmaj
F["kat_mat1_x_dense1"] = [[24., 31., 24., 17., 10.], [11., 18., 11., 9., 2.],
                          [20., 25., 20., 15., 10.], [40., 48., 40., 32., 24.],
                          [21., 28., 21., 14., 7.]]
F["kat_mat5_x_dense2"] = {
    "expected": [[130.04, 150.1, 87.19, 90.89, 99.48, 80.43, 99.3],
                 [217.72, 161.61, 79.47, 121.5, 124.23, 146.91, 157.79],
                 [55.6, 59.95, 86.7, 0.9, 37.4, 71.66, 51.94],
                 [118.18, 123.16, 128.04, 92.02, 106.84, 175.1, 87.36],
                 [43.4, 54.1, 12.65, 44.35, 39.9, 23.4, 76.6]],
    "epsilon": 1e-8}
# ---- sprs/src/sparse/prod.rs:581-595 mul_csr_dense_colmaj: F-order listing of
# the same 5x5 product (column by column)
F["kat_mat1_x_dense1_colmaj_flat"] = [24., 11., 20., 40., 21., 31., 18., 25., 48., 28., 24.,
                                      11., 20., 40., 21., 17., 9., 15., 32., 14., 10., 2.,
                                      10., 24., 7.]
# ---- sprs/src/sparse/prod.rs:461-500 CsMat x CsVec
F["kat_csvec"] = {
    "v": {"dim": 5, "indices": [0, 2, 4], "data": [1., 1., 1.]},
    "mat1_times_v": {"dim": 5, "indices": [0, 1, 2], "data": [3., 5., 5.]},
    "v_times_mat1": {"dim": 5, "indices": [2, 3], "data": [8., 11.]}}
# ---- sprs/src/sparse/prod.rs:312-323 test_csvec_dot_by_binary_search
F["kat_csvec_dot"] = {
    "dim": 8,
    "vec1": {"indices": [0, 2, 4, 6], "data": [1., 1., 1., 1.]},
    "vec2": {"indices": [1, 3, 5, 7], "data": [2., 2., 2., 2.]},
    "vec3": {"indices": [1, 2, 5, 6], "data": [3., 3., 3., 3.]},
    "expected": [["vec1", "vec2", 0.], ["vec1", "vec1", 4.], ["vec2", "vec2", 16.],
                 ["vec1", "vec3", 6.], ["vec2", "vec3", 12.]],
    # sprs/src/sparse/vec.rs:1649-1689 dot_product (same vectors and answers through
    # CsVec::dot) + the dense right-hand side and the two dimension panics
    "dense": [1., 2., 3., 4., 5., 6., 7., 8.], "vec1_dot_dense": 16.,
    "panic_dims": {"sparse": 9, "dense": 9}}
# ---- sprs/src/lib.rs:54-60 README: eye(5) * CsVec == x
F["kat_readme_eye"] = {"n": 5, "x": {"dim": 5, "indices": [0, 2, 4], "data": [1., 2., 3.]}}
# ---- sprs/src/sparse/smmp.rs:476-489 mul_zero_rows ; csmat.rs:3047-3052 issue_99
F["kat_edge"] = {
    "zero_rows": {"a_shape": [0, 11], "b_shape": [11, 11], "c_shape": [0, 11], "c_nnz": 0},
    "issue_99": {"a_shape": [10, 1], "b_shape": [1, 9], "c_shape": [10, 9], "c_nnz": 0}}
# ---- sprs/src/sparse/triplet.rs:342-453 TriMat KATs (expected matrices are given as CSC)
_tri_expected_csc = csmat("CSC", (4, 4), [0, 2, 3, 4, 6], [0, 1, 0, 3, 2, 3],
                          [1., 3., 2., 5., 4., 6.])
F["kat_triplets"] = {
    "incremental": {"shape": [4, 4], "rows": [0, 0, 1, 2, 3, 3], "cols": [0, 1, 0, 3, 2, 3],
                    "data": [1., 2., 3., 4., 5., 6.], "expected_csc": _tri_expected_csc},
    "unordered": {"shape": [4, 4], "rows": [0, 0, 1, 2, 3, 3], "cols": [1, 0, 0, 3, 3, 2],
                  "data": [2., 1., 3., 4., 6., 5.], "expected_csc": _tri_expected_csc},
    "additions": {"shape": [4, 4], "rows": [0, 0, 3, 1, 2, 3, 3], "cols": [1, 0, 2, 0, 3, 3, 2],
                  "data": [2., 1., 3., 3., 4., 6., 2.], "expected_csc": _tri_expected_csc},
    "from_vecs": {"shape": [5, 4], "rows": [0, 0, 1, 2, 3, 3, 4, 4],
                  "cols": [0, 1, 0, 3, 2, 3, 1, 3], "data": [1., 2., 3., 4., 5., 6., 7., 8.],
                  "expected_csc": csmat("CSC", (5, 4), [0, 2, 4, 5, 8],
                                        [0, 1, 0, 4, 3, 2, 3, 4],
                                        [1., 3., 2., 7., 5., 4., 6., 8.])},
    # triplet.rs:571-580 triplet_empty_lines (gh#170), first part
    "empty": {"shape": [2, 4], "rows": [], "cols": [], "data": [],
              "expected_csr_indptr": [0, 0, 0]}}
# ---- sprs/src/sparse/linalg/bicgstab.rs:356-390 test_bicgstab_f64 (and the doc example
# :29-68): solve() must return Ok within 50 iterations at tol 1e-60, i.e. reach an exactly
# zero true residual, and |1 - b/b_recovered| < tol for b_recovered = A x.  `x_exact` is the
# rational solution of the system (checked below), not something the reference states.
F["kat_bicgstab"] = {
    "a": csmat("CSC", (4, 4), [0, 2, 4, 6, 8], [0, 3, 1, 2, 1, 2, 0, 3],
               [1.0, 2., 21., 6., 6., 2., 2., 8.]),
    "b": [1.0, 1.0, 1.0, 1.0], "x0": [1.0, 1.0, 1.0, 1.0], "tol": 1e-60, "max_iter": 50,
    "soft_restart_threshold": 0.1,  # bicgstab.rs:134
    "x_exact": [1.5, -2.0 / 3.0, 2.5, -0.25]}
# ---- sprs/src/sparse/prod.rs:604-605 layout-sweep tolerances
F["assert_close"] = {"rtol": 1e-7, "atol": 1e-12}


def to_scipy(m):
    cls = sp.csr_matrix if m["storage"] == "CSR" else sp.csc_matrix
    return cls((np.array(m["data"]), np.array(m["indices"]), np.array(m["indptr"])),
               shape=tuple(m["shape"]))


def same(a, m):
    b = to_scipy(m)
    return (a != b).nnz == 0 and a.shape == b.shape


def main():
    m1, m1c, m2, m4, m5 = (to_scipy(F[k]) for k in ("mat1", "mat1_csc", "mat2", "mat4", "mat5"))
    assert (m1 != m1c).nnz == 0
    assert same(m1 @ m1, F["mat1_self_matprod"])
    assert same(m1 @ m2, F["mat1_matprod_mat2"])
    assert same(m1c @ m4, F["mat1_csc_matprod_mat4"])
    d1, d2 = np.array(F["mat_dense1"]), np.array(F["mat_dense2"])
    assert np.array_equal(m1 @ d1, np.array(F["kat_mat1_x_dense1"]))
    assert np.allclose(m5 @ d2, np.array(F["kat_mat5_x_dense2"]["expected"]), atol=1e-8, rtol=0)
    assert np.array_equal((m1 @ d1).flatten(order="F"),
                          np.array(F["kat_mat1_x_dense1_colmaj_flat"]))
    for k in ("kat_mul_csr_vec", "kat_mul_csc_vec"):
        y = to_scipy(F[k]["mat"]) @ np.array(F[k]["x"])
        assert np.abs(y - np.array(F[k]["expected"])).max() < F[k]["epsilon"]
    for name, k in F["kat_triplets"].items():
        if "expected_csc" in k:
            got = sp.coo_matrix((k["data"], (k["rows"], k["cols"])), shape=tuple(k["shape"])).tocsc()
            got.sum_duplicates()
            assert same(got, k["expected_csc"]), name
    kb = F["kat_bicgstab"]
    assert np.allclose(to_scipy(kb["a"]) @ np.array(kb["x_exact"]), kb["b"], rtol=0, atol=1e-15)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sprs_fixtures.json")
    with open(out, "w") as f:
        json.dump(F, f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
