// spmv.cu -- CSR x dense-vector product for sm_100a (B200).
//
// Replaces prod::mul_acc_mat_vec_csr (sprs/src/sparse/prod.rs:103-127) and the
// one-column case of prod::csr_mulacc_dense_colmaj (prod.rs:274-298), which is what
// `&A * &x` runs (sprs/src/sparse/csmat.rs:2142-2148).
//
// Design (DESIGN.md "SpMV"): the nnz stream is cut into fixed tiles of SPMV_TILE
// non-zeros (not rows), so every CTA streams the same number of bytes whatever the
// row-length distribution (R-MAT rows are heavily skewed).  Per tile:
//   1. one elected thread issues two 1-D TMA bulk copies (cp.async.bulk ->
//      SASS UBLKCP) that land the tile's `data` (16 KB) and `indices` (8 KB) in
//      shared memory, completion on an mbarrier; L2 policy evict_first because the
//      matrix is streamed exactly once;
//   2. phase A: every thread gathers x[col] for 8 non-zeros (L2 policy evict_last:
//      x is the only re-used operand), multiplies (unfused, like MulAcc::mul_acc,
//      mul_acc.rs:28-30) and writes the products back to shared memory.  The
//      L1TEX/LSU pipe carries only the gathers -- the matrix stream bypasses it;
//   3. phase B: the rows that END in this tile are reduced from shared memory by
//      lane groups of G = 1..32 lanes (G picked per tile from its mean row length,
//      very long rows go through a per-tile warp queue), and y is written once.
//      The row that continues into the next tile leaves its partial in carry[t].
//   4. a second tiny kernel adds the carries in tile order (deterministic, no atomics).
// Rows of <= 6 nnz-per-row tiles are summed by one thread in storage order, i.e.
// bit-identical to the reference's sequential sum; longer rows use a tree and agree
// to rounding (parity gate: |d| <= 1e-6 * sum|terms|).
//
// Algorithmic bytes per nnz: 12 (8 data + 4 index) + 8 per row (y) -- the
// BASELINE roofline 12*nnz + 8*n; indptr (4 B/row) and x gathers are overhead.

#include "common.cuh"

namespace {

constexpr int SPMV_TILE = 2048;  // nnz per tile: 24 KB of shared memory
constexpr int SPMV_NT = 256;     // threads per CTA: 8 gathers in flight per thread
constexpr int SPMV_EPT = SPMV_TILE / SPMV_NT;
constexpr int SPMV_QCAP = SPMV_TILE / 32 + 1;

// ---- PTX wrappers: mbarrier + 1-D TMA bulk copy + L2 cache policies -------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes,
                                         uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
        "[%0], [%1], %2, [%3], %4;" ::"r"(smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ double ldg_f64_hint(const double* p, uint64_t policy) {
    double v;
    asm("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
    return v;
}

// ---- partition: tile_row[t] = first row whose end lies beyond nnz position t*TILE
template <typename P>
__global__ void tile_row_kernel(const P* __restrict__ indptr, uint32_t rows, uint64_t n_tiles,
                                uint32_t* __restrict__ tile_row) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    if (t == 0) {
        tile_row[0] = 0;  // leading empty rows belong to tile 0
        return;
    }
    if (t == n_tiles) {
        tile_row[t] = rows;  // trailing empty rows belong to the last tile
        return;
    }
    const uint64_t k0 = t * (uint64_t)SPMV_TILE;
    uint32_t lo = 0, hi = rows;  // first r with indptr[r+1] > k0
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint64_t)indptr[(size_t)mid + 1] > k0)
            hi = mid;
        else
            lo = mid + 1;
    }
    tile_row[t] = lo;
}

struct TileCtx {
    uint64_t k0, k1;
    uint32_t r1;  // first row NOT owned (== carry row when < rows)
    double* y;
    double* carry_slot;
    int accumulate;
};

__device__ __forceinline__ void emit_row(const TileCtx& tc, uint64_t r, double sum) {
    if (r < tc.r1) {
        tc.y[r] = tc.accumulate ? __dadd_rn(tc.y[r], sum) : sum;
    } else {
        *tc.carry_slot = sum;  // row continues in a later tile: spmv_fixup_kernel adds it
    }
}

template <typename P, int G>
__device__ __forceinline__ void reduce_rows(const TileCtx& tc, const P* __restrict__ indptr,
                                            const double* sprod, uint32_t r0, uint64_t r_last,
                                            int* qcount, uint32_t* qrow, int* qs, int* qe) {
    constexpr int NG = SPMV_NT / G;
    const int gid = threadIdx.x / G, gl = threadIdx.x % G;
    for (uint64_t base = r0; base <= r_last; base += NG) {
        const uint64_t r = base + gid;
        const bool valid = r <= r_last;
        int ls = 0, le = 0;
        if (valid) {
            uint64_t s = (uint64_t)indptr[r], e = (uint64_t)indptr[r + 1];
            s = s > tc.k0 ? s : tc.k0;
            e = e < tc.k1 ? e : tc.k1;
            if (e > s) {
                ls = (int)(s - tc.k0);
                le = (int)(e - tc.k0);
            }
        }
        const bool is_long = (le - ls) > 32 * G;
        double acc = 0.0;
        if (!is_long) {
            for (int j = ls + gl; j < le; j += G) acc = __dadd_rn(acc, sprod[j]);
        }
#pragma unroll
        for (int o = G / 2; o > 0; o >>= 1)
            acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (gl == 0 && valid) {
            if (is_long) {
                const int q = atomicAdd(qcount, 1);
                qrow[q] = (uint32_t)r;
                qs[q] = ls;
                qe[q] = le;
            } else {
                emit_row(tc, r, acc);
            }
        }
    }
}

template <typename P>
__global__ void __launch_bounds__(SPMV_NT)
    spmv_tile_kernel(const P* __restrict__ indptr, const uint32_t* __restrict__ indices,
                     const double* __restrict__ data, const uint32_t* __restrict__ tile_row,
                     const double* __restrict__ x, double* __restrict__ y,
                     double* __restrict__ carry, uint64_t nnz, uint32_t rows, int accumulate) {
    __shared__ __align__(128) double sprod[SPMV_TILE];
    __shared__ __align__(128) uint32_t sidx[SPMV_TILE];
    __shared__ __align__(8) uint64_t bar;
    __shared__ int qcount;
    __shared__ uint32_t qrow[SPMV_QCAP];
    __shared__ int qs[SPMV_QCAP], qe[SPMV_QCAP];

    const int tid = threadIdx.x;
    const uint64_t t = blockIdx.x;
    const uint64_t k0 = t * (uint64_t)SPMV_TILE;
    const uint64_t k1 = (k0 + SPMV_TILE < nnz) ? k0 + SPMV_TILE : nnz;
    const int cnt = (int)(k1 - k0);
    const bool full = cnt == SPMV_TILE;

    if (tid == 0) {
        qcount = 0;
        if (full) mbar_init(&bar, 1);
    }
    __syncthreads();
    if (full) {
        if (tid == 0) {
            const uint64_t pol = policy_evict_first();
            mbar_expect_tx(&bar, SPMV_TILE * 12);
            bulk_g2s(sprod, data + k0, SPMV_TILE * 8, &bar, pol);
            bulk_g2s(sidx, indices + k0, SPMV_TILE * 4, &bar, pol);
        }
    } else {  // ragged last tile: guarded loads (no out-of-bounds bulk copy)
        for (int e = tid; e < cnt; e += SPMV_NT) {
            sprod[e] = data[k0 + e];
            sidx[e] = indices[k0 + e];
        }
    }
    const uint32_t r0 = tile_row[t], r1 = tile_row[t + 1];  // overlaps the TMA flight
    const uint64_t polx = policy_evict_last();
    if (full) {
        mbar_wait(&bar, 0);
        // ---- phase A: gather, multiply, products back to shared memory
        uint32_t c[SPMV_EPT];
        double xv[SPMV_EPT];
#pragma unroll
        for (int i = 0; i < SPMV_EPT; ++i) c[i] = sidx[tid + i * SPMV_NT];
#pragma unroll
        for (int i = 0; i < SPMV_EPT; ++i) xv[i] = ldg_f64_hint(x + c[i], polx);
#pragma unroll
        for (int i = 0; i < SPMV_EPT; ++i)
            sprod[tid + i * SPMV_NT] = __dmul_rn(sprod[tid + i * SPMV_NT], xv[i]);
    } else {
        __syncthreads();
        for (int e = tid; e < cnt; e += SPMV_NT)
            sprod[e] = __dmul_rn(sprod[e], ldg_f64_hint(x + sidx[e], polx));
    }
    __syncthreads();

    // ---- phase B: segmented reduction of the rows that end (or start) in this tile
    TileCtx tc;
    tc.k0 = k0;
    tc.k1 = k1;
    tc.r1 = r1;
    tc.y = y;
    tc.carry_slot = carry + t;
    tc.accumulate = accumulate;
    const uint64_t r_last = (r1 < rows) ? (uint64_t)r1 : (uint64_t)r1 - 1;  // carry row incl.
    const uint64_t nrows_t = r_last - r0 + 1;
    const uint32_t avg = (uint32_t)((uint64_t)cnt / nrows_t);
    if (avg <= 6)
        reduce_rows<P, 1>(tc, indptr, sprod, r0, r_last, &qcount, qrow, qs, qe);
    else if (avg <= 12)
        reduce_rows<P, 2>(tc, indptr, sprod, r0, r_last, &qcount, qrow, qs, qe);
    else if (avg <= 24)
        reduce_rows<P, 4>(tc, indptr, sprod, r0, r_last, &qcount, qrow, qs, qe);
    else if (avg <= 48)
        reduce_rows<P, 8>(tc, indptr, sprod, r0, r_last, &qcount, qrow, qs, qe);
    else if (avg <= 96)
        reduce_rows<P, 16>(tc, indptr, sprod, r0, r_last, &qcount, qrow, qs, qe);
    else
        reduce_rows<P, 32>(tc, indptr, sprod, r0, r_last, &qcount, qrow, qs, qe);
    __syncthreads();
    const int nq = qcount;  // rows too long for their lane group: one warp each
    const int warp = tid >> 5, lane = tid & 31;
    for (int q = warp; q < nq; q += SPMV_NT / 32) {
        double acc = 0.0;
        for (int j = qs[q] + lane; j < qe[q]; j += 32) acc = __dadd_rn(acc, sprod[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
            acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
        if (lane == 0) emit_row(tc, qrow[q], acc);
    }
}

// carries: tile t left the partial sum of row tile_row[t+1] in carry[t]; consecutive
// tiles with the same carry row form a run that is summed in tile order by its head.
__global__ void spmv_fixup_kernel(const uint32_t* __restrict__ tile_row,
                                  const double* __restrict__ carry, double* __restrict__ y,
                                  uint64_t n_tiles) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t + 1 >= n_tiles) return;
    const uint32_t row = tile_row[t + 1];
    if (t > 0 && tile_row[t] == row) return;  // not the head of its run
    double sum = carry[t];
    for (uint64_t u = t + 1; u + 1 < n_tiles && tile_row[u + 1] == row; ++u)
        sum = __dadd_rn(sum, carry[u]);
    y[row] = __dadd_rn(y[row], sum);
}

}  // namespace

int spmv_prepare(sprs_b200_ctx* ctx, sprs_b200_csmat* m, cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR) return SPRS_B200_OK;  // CSC mirrors are converted first
    m->n_tiles = m->nnz == 0 ? 1 : (m->nnz + SPMV_TILE - 1) / SPMV_TILE;
    SPRS_CUDA(ctx, cudaMalloc((void**)&m->d_tile_row, (m->n_tiles + 1) * sizeof(uint32_t)));
    SPRS_CUDA(ctx, cudaMalloc((void**)&m->d_carry, m->n_tiles * sizeof(double)));
    const uint64_t n = m->n_tiles + 1;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (m->indptr_bytes == 4)
        tile_row_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)m->d_indptr,
                                                       (uint32_t)m->rows, m->n_tiles,
                                                       m->d_tile_row);
    else
        tile_row_kernel<uint64_t><<<grid, 256, 0, s>>>((const uint64_t*)m->d_indptr,
                                                       (uint32_t)m->rows, m->n_tiles,
                                                       m->d_tile_row);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

int spmv_launch(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x, double* d_y,
                int accumulate, cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR)
        SPRS_FAIL(ctx, SPRS_B200_ERR_STORAGE, "Storage mismatch: spmv needs a CSR mirror");
    if (m->rows == 0) return SPRS_B200_OK;
    if (!m->d_tile_row) SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "csmat has no SpMV partition");
    if (m->n_tiles > 0x7fffffffull)
        SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "too many tiles for one launch");
    const unsigned grid = (unsigned)m->n_tiles;
    if (m->indptr_bytes == 4)
        spmv_tile_kernel<uint32_t><<<grid, SPMV_NT, 0, s>>>(
            (const uint32_t*)m->d_indptr, m->d_indices, m->d_data, m->d_tile_row, d_x, d_y,
            m->d_carry, m->nnz, (uint32_t)m->rows, accumulate);
    else
        spmv_tile_kernel<uint64_t><<<grid, SPMV_NT, 0, s>>>(
            (const uint64_t*)m->d_indptr, m->d_indices, m->d_data, m->d_tile_row, d_x, d_y,
            m->d_carry, m->nnz, (uint32_t)m->rows, accumulate);
    ctx->launches += 1;
    if (m->n_tiles > 1) {
        const unsigned fgrid = (unsigned)((m->n_tiles - 1 + 255) / 256);
        spmv_fixup_kernel<<<fgrid, 256, 0, s>>>(m->d_tile_row, m->d_carry, d_y, m->n_tiles);
        ctx->launches += 1;
    }
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}
