// test_host_logic.cpp -- CPU-only checks of the C++ host mirror (include/sprs_b200.hpp): the
// structure checks of check_compressed_structure (sprs/src/sparse.rs:300-369), views, the
// panics that fire BEFORE any device work (prod.rs:114-118, smmp.rs:207), and the loud
// failure (ThirdPartyError, never a CPU fallback) when no GPU is present.
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/sprs_b200.hpp"

using namespace sprs;
static int g_checks = 0;
#define CHECK(cond)                                                             \
    do {                                                                        \
        ++g_checks;                                                             \
        if (!(cond)) {                                                          \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            exit(1);                                                            \
        }                                                                       \
    } while (0)

template <class F>
static std::string panic_of(F f) {
    try { f(); } catch (const Panic& p) { return p.what(); } catch (const ThirdPartyError&) { return "THIRD_PARTY"; }
    return "";
}

int main(int argc, char** argv) {
    const bool expect_gpu = argc > 1 && std::string(argv[1]) == "gpu";
    // structure checks
    CHECK(panic_of([] { CsMat::new_({2, 2}, {0, 2, 1}, {0, 1}, {1., 2.}); }) == "Unsorted indptr");
    CHECK(panic_of([] { CsMat::new_({2, 2}, {0, 1, 2}, {0, 5}, {1., 2.}); }) == "Out of bounds index");
    CHECK(panic_of([] { CsMat::new_({1, 3}, {0, 2}, {2, 1}, {1., 2.}); }) == "Unsorted indices");
    CHECK(panic_of([] { CsMat::new_({2, 2}, {0, 1}, {0}, {1.}); }) == "Indptr length does not match dimension");
    CHECK(panic_of([] { CsVec(3, {2, 1}, {1., 2.}); }) == "Unsorted or out-of-bounds indices");
    // views
    auto m = CsMat::new_({2, 3}, {0, 1, 2}, {2, 0}, {1., 2.});
    auto t = m.transpose_view();
    CHECK(t.is_csc() && t.rows() == 3 && t.cols() == 2 && t.nnz() == 2);
    auto e = CsMat::eye(4).slice_outer(1, 3);
    CHECK(e.rows() == 2 && e.cols() == 4 && e.nnz() == 2 && e.indptr()[0] == 1);
    CHECK(m.to_dense_at(0, 2) == 1. && m.to_dense_at(1, 0) == 2. && m.to_dense_at(1, 1) == 0.);
    Array2 a = Array2::from_rows({{1, 2, 3}, {4, 5, 6}});
    CHECK(a.reversed_axes()(2, 1) == 6. && a.to_f_order().rs == 1 && a.to_f_order() == a);
    // panics fire before the device is touched
    Array1 x4(4, 0.0), y2(2, 0.0), x3(3, 0.0);
    CHECK(panic_of([&] { (void)(m * x4); }) == "Dimension mismatch");
    CHECK(panic_of([&] { prod::mul_acc_mat_vec_csr(m, x4, y2); }) == "Dimension mismatch");
    CHECK(panic_of([&] { prod::mul_acc_mat_vec_csc(m, x3, y2); }) == "Storage mismatch");
    CHECK(panic_of([&] { (void)smmp::mul_csr_csr(m, m); }) == "Dimension mismatch");
    Array2 b = Array2::zeros(4, 2), out = Array2::zeros(2, 2);
    CHECK(panic_of([&] { prod::csr_mulacc_dense_rowmaj(m, b, out); }) == "Dimension mismatch");
    // the product itself needs the GPU: without one it fails loudly
    const std::string r = panic_of([&] { (void)(m * x3); });
    if (expect_gpu) CHECK(r == "");
    else CHECK(r == "THIRD_PARTY");
    printf("OK %d checks\n", g_checks);
    return 0;
}
