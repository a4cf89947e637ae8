"""CPU-only: the C-ABI library loads and exports every symbol include/sprs_b200.h declares,
the Python binding covers all of them, the product never links the oracle, and it fails
loudly (no fallback) when there is no GPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "sprs_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sprs_b200_\w+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    import sprs_b200
    names = declared_symbols()
    assert len(names) >= 35
    lib = ctypes.CDLL(sprs_b200._lib.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), "libsprs_b200.so does not export " + n
    assert set(names) == set(sprs_b200._lib.PROTOTYPES), \
        set(names) ^ set(sprs_b200._lib.PROTOTYPES)
    sprs_b200._lib.load()
    assert lib.sprs_b200_version() >= 100


def test_product_does_not_link_oracle():
    import sprs_b200
    out = subprocess.run(["ldd", sprs_b200._lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    for root, _, files in os.walk(os.path.join(ROOT, "sprs_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "liboracle" not in txt, f


def test_no_cpu_fallback_without_gpu():
    import sprs_b200
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("GPU present")
    with pytest.raises(sprs_b200.ThirdPartyError):
        sprs_b200.Context(0)
    import numpy as np
    a = sprs_b200.CsMat.eye(3)
    with pytest.raises(sprs_b200.ThirdPartyError):
        a * np.ones(3)
    with pytest.raises(sprs_b200.ThirdPartyError):  # the solver has no host path either
        sprs_b200.linalg.BiCGSTAB.solve(a, np.ones(3), np.ones(3), 1e-9, 10)


def test_host_structure_checks():
    """check_compressed_structure (sparse.rs:300-369) runs on the host before any upload."""
    import numpy as np
    import sprs_b200 as sp
    with pytest.raises(sp.SprsPanic):
        sp.CsMat.new((2, 2), [0, 2, 1], [0, 1], [1., 2.])          # unsorted indptr
    with pytest.raises(sp.SprsPanic):
        sp.CsMat.new((2, 2), [0, 1, 2], [0, 5], [1., 2.])          # out of bounds
    with pytest.raises(sp.SprsPanic):
        sp.CsMat.new((1, 3), [0, 2], [2, 1], [1., 2.])             # unsorted indices
    m = sp.CsMat.new((2, 3), [0, 1, 2], [2, 0], [1., 2.])
    assert m.transpose_view().shape == (3, 2) and m.transpose_view().is_csc()
    s = sp.CsMat.new((4, 4), [0, 1, 2, 3, 4], [0, 1, 2, 3], np.ones(4)).slice_outer(1, 3)
    assert s.shape == (2, 4) and s.indptr[0] == 1 and s.nnz() == 2
    # sprs/tests/slicing.rs:4-17, 50-75: slice_outer on eye(11), closed and open ranges
    eye = sp.CsMat.eye(11)
    assert list(eye.slice_outer(2, 7)) == [(1.0, (i, i + 2)) for i in range(5)]
    assert list(eye.slice_outer(None, 5)) == [(1.0, (i, i)) for i in range(5)]
    assert list(eye.slice_outer(9, None)) == [(1.0, (0, 9)), (1.0, (1, 10))]
    assert list(eye.slice_outer(slice(9, None))) == [(1.0, (0, 9)), (1.0, (1, 10))]
    assert eye.slice_outer() == eye
    with pytest.raises(sp.SprsPanic):
        eye.slice_outer(5, 12)
    # sprs/tests/gh374.rs: a transposition whose row count does not fit the index type panics
    # before any work (csmat.rs:1794-1797).  2^31 rows with i32 indices; the panic precedes
    # every use of the arrays, so a header-only object is enough (no 8 GB indptr)
    big = object.__new__(sp.CsMat)
    big.storage, big.shape = sp.CSR, (1 << 31, 16)
    big.indptr, big.indices, big.data = np.zeros(2, np.int64), np.zeros(0, np.int32), np.zeros(0)
    big._ctx = big._dev = None
    with pytest.raises(sp.SprsPanic, match="Index type is not large enough to hold"):
        big.to_other_storage()


def test_bicgstab_contract_checks_precede_device_work():
    """`&a * &x0`, `&b - ..` and `&a * &p` panic on mismatched dimensions before anything is
    computed (prod.rs:170, binop.rs:455): no device needed to see the panic."""
    import numpy as np
    import sprs_b200
    from sprs_b200.linalg import BiCGSTAB, NotConverged, bicgstab
    assert bicgstab.BiCGSTAB is BiCGSTAB and issubclass(NotConverged, Exception)
    a = sprs_b200.CsMat.eye(4)
    for x0, b in ((np.ones(3), np.ones(4)), (np.ones(4), np.ones(5)),
                  (sprs_b200.CsVec(5, [0], [1.0]), np.ones(4))):
        with pytest.raises(sprs_b200.SprsPanic, match="Dimension mismatch"):
            BiCGSTAB(a, x0, b)
    rect = sprs_b200.CsMat((3, 4), np.array([0, 1, 2, 3]), np.array([0, 1, 2]), np.ones(3))
    with pytest.raises(sprs_b200.SprsPanic, match="Dimension mismatch"):
        BiCGSTAB(rect, np.ones(4), np.ones(3))
    with pytest.raises(TypeError):
        BiCGSTAB(np.eye(4), np.ones(4), np.ones(4))


def test_cpp_host_mirror_logic_without_gpu():
    """C++ host mirror (include/sprs_b200.hpp): structure checks, views and the panics that
    fire before any device work; the product itself must fail loudly without a GPU."""
    exe = os.path.join(ROOT, "tests", "cpp", "test_host_logic")
    if not os.path.exists(exe):
        subprocess.check_call(["make", "-C", os.path.dirname(exe)])
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    r = subprocess.run([exe] + (["gpu"] if has_gpu else []), capture_output=True, text=True,
                       timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


def test_rust_sys_crate_declares_every_symbol():
    """rust/sprs-b200-sys (source only: no Rust toolchain in the image) declares exactly the
    functions of include/sprs_b200.h -- the crate is what INTEGRATION.md hands a maintainer."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "sprs_b200.h")).read(), flags=re.S)
    in_header = set(re.findall(r"\b(sprs_b200_[a-z0-9_]+)\s*\(", header))
    rust = open(os.path.join(root, "rust", "sprs-b200-sys", "src", "lib.rs")).read()
    in_rust = set(re.findall(r"pub fn (sprs_b200_[a-z0-9_]+)", rust))
    assert in_header == in_rust, (sorted(in_header - in_rust), sorted(in_rust - in_header))


def test_ctypes_prototypes_match_header_arity_and_widths():
    """Every ctypes signature in sprs_b200/_lib.py has the parameter count of its C prototype,
    64-bit integers where the header says uint64_t / int64_t and pointers where it has '*' (a
    mismatch would only show as a corrupted argument at run time)."""
    import ctypes as C
    from sprs_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = re.sub(r"/\*.*?\*/", "", open(os.path.join(root, "include", "sprs_b200.h")).read(), flags=re.S)
    protos = re.findall(r"\b(?:int|uint64_t|const char\*)\s+(sprs_b200_[a-z0-9_]+)\s*\(([^;]*?)\)\s*;",
                        header, flags=re.S)
    assert len(protos) >= 60
    for name, params in protos:
        ps = [" ".join(p.split()) for p in params.split(",")]
        ps = [] if ps == ["void"] else ps
        _, args = _lib.PROTOTYPES[name]
        assert len(ps) == len(args), name
        for p, a in zip(ps, args):
            is_ptr = "*" in p or "[" in p or p.startswith("sprs_b200_matvec_fn")
            if is_ptr:
                assert a in (C.c_void_p, C.c_char_p, _lib.MATVEC_FN) or hasattr(a, "contents"), (name, p)
            elif p.startswith(("uint64_t", "int64_t")):
                assert C.sizeof(a) == 8, (name, p)
            elif p.startswith("double"):
                assert a is C.c_double, (name, p)
            else:
                assert p.startswith("int ") and a is C.c_int, (name, p)
