// test_bicgstab.cpp -- the reference's BiCGSTAB test (sprs/src/sparse/linalg/bicgstab.rs:
// 356-390 test_bicgstab_f64, also the module's doc example :29-68) replayed through the C++
// host mirror on the GPU, plus a larger diagonally dominant system.  Run by
// tests/test_gpu_zz_late.py; exits non-zero on the first failure.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>

#include "../../include/sprs_b200.hpp"

using namespace sprs;
using sprs::linalg::bicgstab::BiCGSTAB;
static int g_checks = 0;
#define CHECK(cond)                                                             \
    do {                                                                        \
        ++g_checks;                                                             \
        if (!(cond)) {                                                          \
            fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond);    \
            exit(1);                                                            \
        }                                                                       \
    } while (0)

// bicgstab.rs:356-390
static void test_bicgstab_f64() {
    auto a = CsMat::new_csc({4, 4}, {0, 2, 4, 6, 8}, {0, 3, 1, 2, 1, 2, 0, 3},
                            {1.0, 2., 21., 6., 6., 2., 2., 8.});
    const double tol = 1e-60;
    const size_t max_iter = 50;
    Array1 b(4, 1.0), x0(4, 1.0);
    auto res = BiCGSTAB<size_t, size_t>::solve(a, x0, b, tol, max_iter);
    CHECK(res.first);  // .unwrap()
    const Array1 x = res.second->x();
    const Array1 b_recovered = a * x;
    printf("Iteration count %zu\n", res.second->iteration_count());
    printf("Soft restart count %zu\n", res.second->soft_restart_count());
    printf("Hard restart count %zu\n", res.second->hard_restart_count());
    for (size_t i = 0; i < 4; ++i) CHECK(std::fabs(1.0 - b[i] / b_recovered[i]) < tol);
    CHECK(res.second->iteration_count() <= max_iter);
    CHECK(res.second->hard_restart_count() >= 1);
    CHECK(res.second->soft_restart_threshold() == 0.1);  // bicgstab.rs:134
}

// tridiagonal-plus-far-diagonal, strictly diagonally dominant and non-symmetric
static void dominant_system() {
    using M = CsMatI<uint32_t, uint32_t>;
    const size_t n = 20000;
    std::vector<uint32_t> ip(1, 0), ind;
    std::vector<double> d;
    for (size_t i = 0; i < n; ++i) {
        const size_t far = (i * 7919 + 13) % n;
        std::vector<std::pair<size_t, double>> row;
        if (i > 0) row.push_back({i - 1, -1.0});
        row.push_back({i, 6.0 + (double)(i % 5)});
        if (i + 1 < n) row.push_back({i + 1, -2.0});
        if (far + 1 < i || far > i + 1) row.push_back({far, 0.5});
        std::sort(row.begin(), row.end());
        for (auto& e : row) {
            ind.push_back((uint32_t)e.first);
            d.push_back(e.second);
        }
        ip.push_back((uint32_t)ind.size());
    }
    auto a = M::new_({n, n}, ip, ind, d);
    Array1 x_true(n), x0(n, 0.0);
    for (size_t i = 0; i < n; ++i) x_true[i] = std::sin(0.01 * (double)i) + 2.0;
    const Array1 b = a * x_true;
    auto res = BiCGSTAB<uint32_t, uint32_t>::solve(a, x0, b, 1e-8, 200);
    CHECK(res.first);
    const Array1 x = res.second->x();
    double worst = 0.0;
    for (size_t i = 0; i < n; ++i) worst = std::fmax(worst, std::fabs(x[i] - x_true[i]));
    CHECK(worst < 1e-7);
    // the accepted error is the TRUE residual norm (hard restart before returning)
    const Array1 ax = a * x;
    double rr = 0.0;
    for (size_t i = 0; i < n; ++i) rr += (b[i] - ax[i]) * (b[i] - ax[i]);
    CHECK(std::sqrt(rr) < 1e-8 * 1.001);
    CHECK(std::fabs(std::sqrt(rr) - res.second->err()) <= 1e-12);
    // operator form: the same SpMV behind a caller-supplied operator -> the same bits per step
    {
        Context& ctx = Context::thread_default();
        BiCGSTAB<uint32_t, uint32_t> s_mat(a, x0, b);
        BiCGSTAB<uint32_t, uint32_t> s_op(n, [&](const double* d_x, double* d_y, void* stream) {
            ctx.check(sprs_b200_spmv_dev(ctx.handle(), a.device(), d_x, d_y, 0, stream));
        }, x0, b);
        CHECK(s_op.err() == s_mat.err());
        for (int it = 0; it < 4; ++it) CHECK(s_op.step() == s_mat.step());
        const Array1 xo = s_op.x(), xm = s_mat.x();
        bool same = true;
        for (size_t i = 0; i < n; ++i) same = same && xo[i] == xm[i];
        CHECK(same);
    }
    // an impossible tolerance runs into the iteration limit: Err, state still returned
    auto res2 = BiCGSTAB<uint32_t, uint32_t>::solve(a, x0, b, 0.0, 3);
    CHECK(!res2.first);
    CHECK(res2.second->iteration_count() == 3);
}

static void contract_violations() {
    auto a = CsMat::eye(4);
    bool panicked = false;
    try {
        BiCGSTAB<size_t, size_t> s(a, Array1(3, 1.0), Array1(4, 1.0));
    } catch (const Panic& e) {
        panicked = std::string(e.what()) == "Dimension mismatch";
    }
    CHECK(panicked);
}

int main() {
    try {
        test_bicgstab_f64();
        dominant_system();
        contract_violations();
    } catch (const std::exception& e) {
        fprintf(stderr, "EXCEPTION: %s\n", e.what());
        return 2;
    }
    printf("OK %d checks\n", g_checks);
    return 0;
}
