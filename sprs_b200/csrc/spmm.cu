// spmm.cu -- CSR x dense row-major matrix for sm_100a (B200).
//
// Replaces prod::csr_mulacc_dense_rowmaj (sprs/src/sparse/prod.rs:189-214), the
// kernel `&A * &B` picks when B has >= 8 columns (sprs/src/sparse/csmat.rs:2009-2018):
//     out[i,:] += sum_j A[i,j] * B[j,:]        (k-wide axpy per non-zero)
// B and C are C-order with leading dimensions ldb / ldc.
//
// One warp owns one row of A and a panel of up to 32*KV output columns; the row's
// (index, value) pairs are read coalesced 32 at a time and broadcast with shuffles,
// every lane gathers its slice of the B row (a 512-byte row at k = 64 is two fully
// coalesced 256-byte requests per non-zero) and keeps KV accumulators in registers;
// C is written exactly once.  Each output element is the sequential, unfused sum in
// storage order (mul_acc.rs:28-30), so results are bit-identical to the reference.
//
// Compulsory bytes: 12*nnz + 8*k*(cols + rows); the B-row gathers (8*k per nnz) are
// served by L2 / HBM depending on B's size (DESIGN.md "SpMM").

#include <cstdlib>

#include "common.cuh"

namespace {

constexpr int SPMM_NT = 256;

template <typename P, int KV>
__global__ void __launch_bounds__(SPMM_NT)
    spmm_rowmaj_kernel(const P* __restrict__ indptr, const uint32_t* __restrict__ indices,
                       const double* __restrict__ data, const double* __restrict__ B,
                       uint64_t ldb, uint32_t k, double* __restrict__ C, uint64_t ldc,
                       uint32_t rows, int accumulate) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)SPMM_NT + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * SPMM_NT) >> 5;
    for (uint64_t row = warp0; row < rows; row += nwarps) {
        const uint64_t s = (uint64_t)indptr[row], e = (uint64_t)indptr[row + 1];
        double* crow = C + row * ldc;
        for (uint32_t c0 = 0; c0 < k; c0 += 32 * KV) {
            double acc[KV];
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const uint32_t c = c0 + lane + 32 * q;
                acc[q] = (accumulate && c < k) ? crow[c] : 0.0;
            }
            for (uint64_t kk = s; kk < e; kk += 32) {
                const bool in = kk + lane < e;
                const uint32_t my_idx = in ? indices[kk + lane] : 0u;
                const double my_val = in ? data[kk + lane] : 0.0;
                const int n = (e - kk) < 32 ? (int)(e - kk) : 32;
                for (int j = 0; j < n; ++j) {
                    const uint32_t col = __shfl_sync(0xffffffffu, my_idx, j);
                    const double v = __shfl_sync(0xffffffffu, my_val, j);
                    const double* brow = B + (uint64_t)col * ldb;
#pragma unroll
                    for (int q = 0; q < KV; ++q) {
                        const uint32_t c = c0 + lane + 32 * q;
                        if (c < k) acc[q] = __dadd_rn(acc[q], __dmul_rn(v, __ldg(brow + c)));
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const uint32_t c = c0 + lane + 32 * q;
                if (c < k) crow[c] = acc[q];
            }
        }
    }
}

// U = non-zeros whose B-row loads are issued together before their products are added (in
// storage order, so the sums are the same bits).  The kernel above, measured in round 1,
// stalls on every B load before the next one is issued: 2 loads in flight per warp -- 64 warps
// x 2 x 256 B per ~1 us of loaded DRAM latency is the 5.0 TB/s it reached
// (profiles/r1_ncu_spmm_v1.csv).  This variant (opt-in, SPRS_B200_SPMM_UNROLL=4, until timed)
// puts 8 loads per warp in flight.
template <typename P, int KV, int U>
__global__ void __launch_bounds__(SPMM_NT)
    spmm_rowmaj_unrolled_kernel(const P* __restrict__ indptr, const uint32_t* __restrict__ indices,
                       const double* __restrict__ data, const double* __restrict__ B,
                       uint64_t ldb, uint32_t k, double* __restrict__ C, uint64_t ldc,
                       uint32_t rows, int accumulate) {
    const int lane = threadIdx.x & 31;
    const uint64_t warp0 = (blockIdx.x * (uint64_t)SPMM_NT + threadIdx.x) >> 5;
    const uint64_t nwarps = ((uint64_t)gridDim.x * SPMM_NT) >> 5;
    for (uint64_t row = warp0; row < rows; row += nwarps) {
        const uint64_t s = (uint64_t)indptr[row], e = (uint64_t)indptr[row + 1];
        double* crow = C + row * ldc;
        for (uint32_t c0 = 0; c0 < k; c0 += 32 * KV) {
            double acc[KV];
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const uint32_t c = c0 + lane + 32 * q;
                acc[q] = (accumulate && c < k) ? crow[c] : 0.0;
            }
            for (uint64_t kk = s; kk < e; kk += 32) {
                const bool in = kk + lane < e;
                const uint32_t my_idx = in ? indices[kk + lane] : 0u;
                const double my_val = in ? data[kk + lane] : 0.0;
                const int n = (e - kk) < 32 ? (int)(e - kk) : 32;
                for (int j = 0; j < n; j += U) {
                    double v[U], b[U][KV];
#pragma unroll
                    for (int u = 0; u < U; ++u) {  // all loads of U non-zeros first
                        const uint32_t col = __shfl_sync(0xffffffffu, my_idx, (j + u) & 31);
                        v[u] = __shfl_sync(0xffffffffu, my_val, (j + u) & 31);
                        const double* brow = B + (uint64_t)col * ldb;
#pragma unroll
                        for (int q = 0; q < KV; ++q) {
                            const uint32_t c = c0 + lane + 32 * q;
                            b[u][q] = (j + u < n && c < k) ? __ldg(brow + c) : 0.0;
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)  // then the sums, in storage order
                        if (j + u < n) {
#pragma unroll
                            for (int q = 0; q < KV; ++q) {
                                const uint32_t c = c0 + lane + 32 * q;
                                if (c < k) acc[q] = __dadd_rn(acc[q], __dmul_rn(v[u], b[u][q]));
                            }
                        }
                }
            }
#pragma unroll
            for (int q = 0; q < KV; ++q) {
                const uint32_t c = c0 + lane + 32 * q;
                if (c < k) crow[c] = acc[q];
            }
        }
    }
}

// L2-blocked variant (opt-in, SPRS_B200_SPMM_PANEL=4|8|16|32; default off until it has been
// measured): B is consumed in column panels of G columns so that the panel (B.rows x G
// doubles, 64 MB for 1M rows at G = 8) stays resident in the 126 MB L2 while A is re-streamed
// once per panel -- at k = 64 the one-pass kernel above gathers a 512-byte B row per
// non-zero and 89 % of those gathers miss L2 (14.35 GB of DRAM traffic for 1.4 GB of
// compulsory bytes).  A group of G lanes owns one row: lane g keeps the
// accumulator of column c0 + g, the row's (index, value) pairs are loaded G at a time and
// broadcast inside the group.  Same sequential unfused sums: bit-identical results.
template <typename P, int G>
__global__ void __launch_bounds__(SPMM_NT)
    spmm_panel_kernel(const P* __restrict__ indptr, const uint32_t* __restrict__ indices,
                      const double* __restrict__ data, const double* __restrict__ B,
                      uint64_t ldb, uint32_t c0, uint32_t k, double* __restrict__ C,
                      uint64_t ldc, uint32_t rows, int accumulate) {
    constexpr int GROUPS = 32 / G;
    const int lane = threadIdx.x & 31, gl = lane % G, gid = lane / G;
    const unsigned gmask = (G == 32 ? 0xffffffffu : ((1u << G) - 1u)) << (gid * G);
    const uint64_t group0 = ((blockIdx.x * (uint64_t)SPMM_NT + threadIdx.x) >> 5) * GROUPS + gid;
    const uint64_t ngroups = (((uint64_t)gridDim.x * SPMM_NT) >> 5) * GROUPS;
    const uint32_t c = c0 + gl;
    const bool live = c < k;
    for (uint64_t row = group0; row < rows; row += ngroups) {
        const uint64_t s = (uint64_t)indptr[row], e = (uint64_t)indptr[row + 1];
        double* crow = C + row * ldc;
        double acc = (accumulate && live) ? crow[c] : 0.0;
        for (uint64_t kk = s; kk < e; kk += G) {
            const bool in = kk + gl < e;
            const uint32_t my_idx = in ? indices[kk + gl] : 0u;
            const double my_val = in ? data[kk + gl] : 0.0;
            const int n = (e - kk) < (uint64_t)G ? (int)(e - kk) : G;
            for (int j = 0; j < n; ++j) {
                const uint32_t col = __shfl_sync(gmask, my_idx, j, G);
                const double v = __shfl_sync(gmask, my_val, j, G);
                if (live) acc = __dadd_rn(acc, __dmul_rn(v, __ldg(B + (uint64_t)col * ldb + c)));
            }
        }
        if (live) crow[c] = acc;
    }
}

int spmm_unroll() {
    static const int u = [] {
        const char* e = getenv("SPRS_B200_SPMM_UNROLL");
        return (e && atoi(e) == 4) ? 4 : 1;
    }();
    return u;
}

int spmm_panel_width() {
    static const int w = [] {
        const char* e = getenv("SPRS_B200_SPMM_PANEL");
        const int v = e ? atoi(e) : 0;
        return (v == 4 || v == 8 || v == 16 || v == 32) ? v : 0;
    }();
    return w;
}

}  // namespace

int spmm_rowmaj_launch(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_b,
                       uint64_t ldb, uint64_t k, double* d_c, uint64_t ldc, int accumulate,
                       cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR)
        SPRS_FAIL(ctx, SPRS_B200_ERR_STORAGE, "Storage mismatch: spmm needs a CSR mirror");
    if (m->rows == 0 || k == 0) return SPRS_B200_OK;
    if (k > 0xffffffffull) SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "k too large");
    if (const int pw = spmm_panel_width()) {
        const uint64_t groups_per_block = (uint64_t)SPMM_NT / pw;
        uint64_t pblocks = (m->rows + groups_per_block - 1) / groups_per_block;
        const uint64_t pcap = (uint64_t)ctx->sm_count * 64;
        if (pblocks > pcap) pblocks = pcap;
        for (uint64_t c0 = 0; c0 < k; c0 += pw) {
#define SPMM_PANEL(P, G)                                                                        \
    spmm_panel_kernel<P, G><<<(unsigned)pblocks, SPMM_NT, 0, s>>>(                              \
        (const P*)m->d_indptr, m->d_indices, m->d_data, d_b, ldb, (uint32_t)c0, (uint32_t)k,    \
        d_c, ldc, (uint32_t)m->rows, accumulate)
            if (m->indptr_bytes == 4) {
                if (pw == 4) SPMM_PANEL(uint32_t, 4);
                else if (pw == 8) SPMM_PANEL(uint32_t, 8);
                else if (pw == 16) SPMM_PANEL(uint32_t, 16);
                else SPMM_PANEL(uint32_t, 32);
            } else {
                if (pw == 4) SPMM_PANEL(uint64_t, 4);
                else if (pw == 8) SPMM_PANEL(uint64_t, 8);
                else if (pw == 16) SPMM_PANEL(uint64_t, 16);
                else SPMM_PANEL(uint64_t, 32);
            }
#undef SPMM_PANEL
            ctx->launches += 1;
        }
        SPRS_CUDA(ctx, cudaGetLastError());
        return SPRS_B200_OK;
    }
    const uint64_t warps_needed = m->rows;
    uint64_t blocks = (warps_needed * 32 + SPMM_NT - 1) / SPMM_NT;
    const uint64_t cap = (uint64_t)ctx->sm_count * 64;
    if (blocks > cap) blocks = cap;
    const unsigned grid = (unsigned)blocks;
#define SPMM_LAUNCH(P, KV)                                                                    \
    do {                                                                                      \
        if (spmm_unroll() == 4)                                                               \
            spmm_rowmaj_unrolled_kernel<P, KV, 4><<<grid, SPMM_NT, 0, s>>>(                            \
                (const P*)m->d_indptr, m->d_indices, m->d_data, d_b, ldb, (uint32_t)k, d_c,   \
                ldc, (uint32_t)m->rows, accumulate);                                          \
        else                                                                                  \
            spmm_rowmaj_kernel<P, KV><<<grid, SPMM_NT, 0, s>>>(                            \
                (const P*)m->d_indptr, m->d_indices, m->d_data, d_b, ldb, (uint32_t)k, d_c,   \
                ldc, (uint32_t)m->rows, accumulate);                                          \
    } while (0)
    if (m->indptr_bytes == 4) {
        if (k <= 32) SPMM_LAUNCH(uint32_t, 1);
        else if (k <= 64) SPMM_LAUNCH(uint32_t, 2);
        else SPMM_LAUNCH(uint32_t, 4);
    } else {
        if (k <= 32) SPMM_LAUNCH(uint64_t, 1);
        else if (k <= 64) SPMM_LAUNCH(uint64_t, 2);
        else SPMM_LAUNCH(uint64_t, 4);
    }
#undef SPMM_LAUNCH
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}
