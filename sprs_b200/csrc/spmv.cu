// spmv.cu -- CSR x dense-vector product for sm_100a (B200).
//
// Replaces prod::mul_acc_mat_vec_csr (sprs/src/sparse/prod.rs:103-127) and the
// one-column case of prod::csr_mulacc_dense_colmaj (prod.rs:274-298), which is what
// `&A * &x` runs (sprs/src/sparse/csmat.rs:2142-2148).
//
// Design (DESIGN.md "SpMV"): the nnz stream is cut into fixed tiles of WT non-zeros
// (not rows), so every tile streams the same number of bytes whatever the row-length
// distribution (R-MAT rows are heavily skewed).  A tile belongs to ONE WARP; warps are
// persistent and autonomous (no CTA-wide barrier anywhere), each running a private
// STAGES-deep TMA pipeline:
//   1. lane 0 issues two 1-D TMA bulk copies per tile (cp.async.bulk -> SASS UBLKCP)
//      that land the tile's `data` and `indices` in the warp's shared-memory stage,
//      completion on a per-stage mbarrier; L2 policy evict_first (the matrix is
//      streamed exactly once), issued STAGES tiles ahead;
//   2. phase A: every lane gathers x[col] for WT/32 non-zeros (all loads in flight
//      before the first use; L2 policy evict_last: x is the only re-used operand),
//      multiplies (unfused, like MulAcc::mul_acc, mul_acc.rs:28-30) and writes the
//      products back to the stage;
//   3. phase B: the rows that END in this tile are reduced from shared memory by lane
//      groups of G = 1..32 lanes (G picked per tile from its mean row length; row
//      boundaries are prefetched into registers before the TMA wait), and y is written
//      once.  The row that continues into the next tile leaves its partial in carry[t];
//   4. a second tiny kernel adds the carries in tile order (deterministic, no atomics).
// Because the warps drift apart, gathers (L1TEX-bound), shared-memory reductions and
// TMA waits of different warps overlap instead of alternating in CTA-wide phases
// (profiles/r1_spmv_notes.md: the first CTA-tile version sat at 44 % L1TEX utilisation
// with barrier + MIO-throttle stalls).
// Rows inside tiles whose mean row length is <= 6 are summed by one lane in storage
// order, i.e. bit-identical to the reference's sequential sum; longer rows use a tree
// and agree to rounding (parity gate: |d| <= 1e-6 * sum|terms|).
//
// Algorithmic bytes per nnz: 12 (8 data + 4 index) + 8 per row (y) -- the
// BASELINE roofline 12*nnz + 8*n; indptr (4 B/row) and x gathers are overhead.

#include "common.cuh"

#include <cstdlib>

namespace {

// ---- PTX wrappers: mbarrier + 1-D TMA bulk copy + L2 cache policies -------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count)
                 : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
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

__device__ __forceinline__ uint32_t ldg_stream_u32(const uint32_t* p, uint64_t policy) {
    uint32_t v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;"
        : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ double ldg_stream_f64(const double* p, uint64_t policy) {
    double v;
    asm("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;"
        : "=d"(v) : "l"(p), "l"(policy));
    return v;
}

// ---- partition: tile_row[t] = first row whose end lies beyond nnz position t*wt
template <typename P>
__global__ void tile_row_kernel(const P* __restrict__ indptr, uint32_t rows, uint64_t n_tiles,
                                uint32_t wt, uint32_t* __restrict__ tile_row) {
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
    const uint64_t k0 = t * (uint64_t)wt;
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

constexpr int SPMV_REG_ROWS = 9;  // tiles touching <= this many rows reduce in registers
// (8 rows through the multi-value butterfly + one more -- typically the row cut by the tile
// end -- through a plain one; a two-pass 16-row version measured slower, sweep v6)

struct TileCtx {
    uint64_t k0, k1;
    uint32_t r1;  // first row NOT owned (== carry row when < rows)
    double* y;    // this GPU's y (target 0)
    const SpmvTargets* yt;  // MULTI only: the kernel parameter itself (constant bank)
    double* carry_slot;
    int accumulate;
};

// y[r] is written to every target buffer: target 0 is this GPU's own y; targets 1.. are the
// peer GPUs' y buffers mapped through CUDA IPC (fused SpMV + all-gather over NVLink: the
// result of a row leaves for the peers the moment it is reduced, overlapped with the rest
// of the kernel, instead of a separate collective afterwards).
template <bool MULTI>
__device__ __forceinline__ void emit_row(const TileCtx& tc, uint64_t r, double sum) {
    if (r < tc.r1) {
        const double v = tc.accumulate ? __dadd_rn(tc.y[r], sum) : sum;
        tc.y[r] = v;
        if (MULTI) {
#pragma unroll
            for (int q = 1; q < SPMV_MAX_TARGETS; ++q)
                if (q < tc.yt->n) tc.yt->p[q][r] = v;
        }
    } else {
        *tc.carry_slot = sum;  // row continues in a later tile: spmv_fixup_kernel adds it
    }
}

// Reduce rows [r0, r_last] of one warp tile with groups of G lanes per row.  Row
// boundaries come 31 rows at a time: lane L holds indptr[rbase + L].
template <typename P, int G, bool MULTI>
__device__ __forceinline__ void reduce_rows_warp(const TileCtx& tc, const P* __restrict__ indptr,
                                                 const double* sprod, uint32_t r0,
                                                 uint64_t r_last, uint64_t b_first, int lane) {
    constexpr int NG = 32 / G;
    const int gid = lane / G, gl = lane % G;
    uint64_t b = b_first;  // boundaries of the first chunk were prefetched by the caller
    for (uint64_t rbase = r0; rbase <= r_last; rbase += 31) {
        if (rbase != r0) {
            const uint64_t rr = rbase + lane;
            b = rr <= r_last + 1 ? (uint64_t)indptr[rr] : 0;
        }
        const int nrows = (r_last - rbase + 1) < 31 ? (int)(r_last - rbase + 1) : 31;
        for (int j0 = 0; j0 < nrows; j0 += NG) {
            const int j = j0 + gid;
            const bool valid = j < nrows;
            const int js = valid ? j : 0;
            uint64_t s = __shfl_sync(0xffffffffu, b, js);
            uint64_t e = __shfl_sync(0xffffffffu, b, js + 1);
            int ls = 0, le = 0;
            s = s > tc.k0 ? s : tc.k0;
            e = e < tc.k1 ? e : tc.k1;
            if (valid && e > s) {
                ls = (int)(s - tc.k0);
                le = (int)(e - tc.k0);
            }
            const bool is_long = (G < 32) && (le - ls) > 16 * G;
            double acc = 0.0;
            if (!is_long)
                for (int q = ls + gl; q < le; q += G) acc = __dadd_rn(acc, sprod[q]);
#pragma unroll
            for (int o = G / 2; o > 0; o >>= 1)
                acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
            if (gl == 0 && valid && !is_long) emit_row<MULTI>(tc, rbase + j, acc);
            if (G < 32) {  // rows too long for their group: the whole warp takes them
                unsigned pending = __ballot_sync(0xffffffffu, gl == 0 && valid && is_long);
                while (pending) {
                    const int src = __ffs(pending) - 1;
                    pending &= pending - 1;
                    const int qs = __shfl_sync(0xffffffffu, ls, src);
                    const int qe = __shfl_sync(0xffffffffu, le, src);
                    const int jj = __shfl_sync(0xffffffffu, j, src);
                    double a2 = 0.0;
                    for (int q = qs + lane; q < qe; q += 32) a2 = __dadd_rn(a2, sprod[q]);
#pragma unroll
                    for (int o = 16; o > 0; o >>= 1)
                        a2 = __dadd_rn(a2, __shfl_xor_sync(0xffffffffu, a2, o));
                    if (lane == 0) emit_row<MULTI>(tc, rbase + jj, a2);
                }
            }
        }
    }
}

// Register-path reduction for tiles that touch at most 8 rows (the bulk of the non-zeros of
// long-row matrices): the products never go to shared memory.  Each lane adds its EPL
// products into up to 8 per-row partials (its element e = lane + 32*i belongs to row j iff
// bl_j <= e < bl_{j+1}), then ONE multi-value butterfly reduces the 8 partials across the
// warp in 18 shuffles (xor 16 / 8 / 4 halve the number of live values, xor 2 / 1 finish);
// lane 4*j ends up with the sum of row j.  Versus the shared-memory path this removes the
// product store + reload and ~3/4 of the shuffles -- all of them L1TEX/MIO wavefronts, the
// pipe the gathers saturate (profiles/r1_spmv_notes.md section 4).  Deterministic order.
template <int EPL, bool MULTI>
__device__ __forceinline__ void reduce_rows_regs(const TileCtx& tc, const double (&p)[EPL],
                                                 int bl, uint32_t r0, int jbase, int nrows,
                                                 int lane) {
    // rows jbase .. jbase+nrows-1 of the tile (nrows <= 8); bl: lane L = boundary of row L
    constexpr unsigned FULL = 0xffffffffu;
    double part[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) part[j] = 0.0;
    int lo = __shfl_sync(FULL, bl, jbase);
    if (nrows > 8) {  // ninth row: everything from boundary 8 to boundary 9, plain butterfly
        const int lo8 = __shfl_sync(FULL, bl, jbase + 8), hi8 = __shfl_sync(FULL, bl, jbase + 9);
        double ex = 0.0;
#pragma unroll
        for (int i = 0; i < EPL; ++i) {
            const int e = lane + 32 * i;
            ex = (e >= lo8 && e < hi8) ? __dadd_rn(ex, p[i]) : ex;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) ex = __dadd_rn(ex, __shfl_xor_sync(FULL, ex, o));
        if (lane == 0) emit_row<MULTI>(tc, (uint64_t)r0 + jbase + 8, ex);
        nrows = 8;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (j < nrows) {  // warp-uniform
            const int hi = __shfl_sync(FULL, bl, jbase + j + 1);
#pragma unroll
            for (int i = 0; i < EPL; ++i) {
                const int e = lane + 32 * i;
                part[j] = (e >= lo && e < hi) ? __dadd_rn(part[j], p[i]) : part[j];
            }
            lo = hi;
        }
    }
    double v4[4], v2[2], v;
    {
        const bool up = lane & 16;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const double send = up ? part[j] : part[j + 4];
            const double keep = up ? part[j + 4] : part[j];
            v4[j] = __dadd_rn(keep, __shfl_xor_sync(FULL, send, 16));
        }
    }
    {
        const bool up = lane & 8;
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const double send = up ? v4[j] : v4[j + 2];
            const double keep = up ? v4[j + 2] : v4[j];
            v2[j] = __dadd_rn(keep, __shfl_xor_sync(FULL, send, 8));
        }
    }
    {
        const bool up = lane & 4;
        const double send = up ? v2[0] : v2[1];
        const double keep = up ? v2[1] : v2[0];
        v = __dadd_rn(keep, __shfl_xor_sync(FULL, send, 4));
    }
    v = __dadd_rn(v, __shfl_xor_sync(FULL, v, 2));
    v = __dadd_rn(v, __shfl_xor_sync(FULL, v, 1));
    const int row = lane >> 2;  // 4*bit4 + 2*bit3 + bit2
    if ((lane & 3) == 0 && row < nrows) emit_row<MULTI>(tc, (uint64_t)r0 + jbase + row, v);
}

// SIGNAL (single-target launches only): the kernel also publishes its progress for the
// pipelined all-gather (spmv_launch_stream_push below).  Tiles are grouped into chunks of
// 2^chunk_shift consecutive tiles; a warp visits its tiles in increasing order, counts the
// ones it finishes inside the current chunk and adds that count to progress[chunk] when it
// moves on to another chunk (release: its y stores are fenced first).
template <typename P, int WT, int STAGES, int NWARPS, int MINB, bool MULTI, bool SIGNAL>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
    spmv_warp_kernel(const P* __restrict__ indptr, const uint32_t* __restrict__ indices,
                     const double* __restrict__ data, const uint32_t* __restrict__ tile_row,
                     const double* __restrict__ x, SpmvTargets yt,
                     double* __restrict__ carry, uint64_t nnz, uint32_t rows, uint64_t t_begin,
                     uint64_t n_tiles /* end of this launch's tile range */, int accumulate,
                     unsigned long long* progress, int chunk_shift,
                     unsigned long long* tile_counter) {
    constexpr int EPL = WT / 32;               // non-zeros per lane per tile
    // SIGNAL flavour with a one-stage ring and a tile counter: tiles are HANDED OUT (every warp
    // claims the next unclaimed tile of the launch) instead of dealt round-robin, so CTAs that
    // share their SM with the put kernel, or start late because it took their slot, simply
    // claim fewer tiles; a static deal would leave their share for a second wave.  Claims are
    // made two tiles ahead, so the atomic's round trip never sits in front of a gather.
    // Which warp reduces a tile does not change any sum.
    constexpr bool DYN = SIGNAL && (STAGES <= 1);
    constexpr bool DIRECT = STAGES == 0;       // no TMA ring: stream through registers
    constexpr int STAGE_BYTES = DIRECT ? WT * 8 : WT * 12;
    constexpr int NST = DIRECT ? 1 : STAGES;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t bars[NWARPS][NST];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned char* wsm = smem_raw + (size_t)warp * NST * STAGE_BYTES;
    // this launch covers tiles [t_begin, n_tiles): the whole matrix, or one chunk of it when the
    // caller pipelines something behind finished row ranges (spmv_launch_tile_range)
    const bool dyn = DYN && tile_counter != nullptr;
    uint64_t gw = t_begin + (uint64_t)blockIdx.x * NWARPS + warp;
    const uint64_t GW = (uint64_t)gridDim.x * NWARPS;
    uint64_t t_claimed = 0;  // dyn: this warp's next tile, claimed one iteration ago
    if (dyn) {
        unsigned long long c0 = 0, c1 = 0;
        if (lane == 0) {
            c0 = atomicAdd(tile_counter, 1ull);
            c1 = atomicAdd(tile_counter, 1ull);
        }
        gw = t_begin + __shfl_sync(0xffffffffu, c0, 0);
        t_claimed = t_begin + __shfl_sync(0xffffffffu, c1, 0);
    }
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t polx = policy_evict_last();

    if (!DIRECT && lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) mbar_init(&bars[warp][s], 1);
        fence_mbar_init();
    }
    __syncwarp();
#define SPMV_ISSUE(T, S)                                                                    \
    do { /* lane 0 only; full tiles only (the ragged last tile is loaded by hand) */        \
        const uint64_t k0_ = (T) * (uint64_t)WT;                                            \
        if (k0_ + WT <= nnz) {                                                              \
            unsigned char* st_ = wsm + (size_t)(S) * STAGE_BYTES;                           \
            mbar_expect_tx(&bars[warp][(S)], STAGE_BYTES);                                  \
            bulk_g2s(st_, data + k0_, WT * 8, &bars[warp][(S)], pol_stream);                \
            bulk_g2s(st_ + WT * 8, indices + k0_, WT * 4, &bars[warp][(S)], pol_stream);    \
        }                                                                                   \
    } while (0)
    if (!DIRECT && lane == 0) {
#pragma unroll
        for (int s = 0; s < NST; ++s) {
            const uint64_t t = gw + (uint64_t)s * GW;
            if (t < n_tiles) SPMV_ISSUE(t, s);
        }
    }
    uint32_t phases = 0;
    int s = 0;
    unsigned long long sig_count = 0;  // SIGNAL: tiles finished in the current chunk
    // Row range and the first 32 row boundaries of a tile are fetched ONE TILE AHEAD, so the
    // two dependent global loads (tile_row -> indptr) never stall the in-order issue in
    // front of the gathers.
    uint32_t r0 = 0, r1 = 0;
    uint64_t b_first = 0;
    if (gw < n_tiles) {
        r0 = tile_row[gw];
        r1 = tile_row[gw + 1];
        const uint64_t rl = (r1 < rows) ? (uint64_t)r1 : (uint64_t)r1 - 1;
        const uint64_t rr = (uint64_t)r0 + lane;
        b_first = rr <= rl + 1 ? (uint64_t)indptr[rr] : 0;
    }
    for (uint64_t t = gw; t < n_tiles;) {
        unsigned long long claim = 0;  // dyn: the tile after next, in flight during this tile
        if (dyn && lane == 0) claim = atomicAdd(tile_counter, 1ull);
        const uint64_t k0 = t * (uint64_t)WT;
        const uint64_t k1 = (k0 + WT < nnz) ? k0 + WT : nnz;
        const int cnt = (int)(k1 - k0);
        const bool full = cnt == WT;
        double* sprod = (double*)(wsm + (size_t)s * STAGE_BYTES);
        uint32_t* sidx = (uint32_t*)(wsm + (size_t)s * STAGE_BYTES + WT * 8);
        const uint64_t r_last = (r1 < rows) ? (uint64_t)r1 : (uint64_t)r1 - 1;
        const uint64_t tnext = dyn ? t_claimed : t + GW;
        uint32_t r0n = 0, r1n = 0;
        const int nrows_t = (r_last - r0 + 1) > 64 ? 64 : (int)(r_last - r0 + 1);
        const bool regpath = full && nrows_t <= SPMV_REG_ROWS;  // warp-uniform
        double preg[EPL];
        if (full && DIRECT) {
            // register path: coalesced streaming loads (no L1 allocation, L2 evict_first),
            // products to the warp's 8*WT-byte shared buffer for the reduction
            uint32_t c[EPL];
            double v[EPL], xv[EPL];
#pragma unroll
            for (int i = 0; i < EPL; ++i)
                c[i] = ldg_stream_u32(indices + k0 + lane + i * 32, pol_stream);
#pragma unroll
            for (int i = 0; i < EPL; ++i)
                v[i] = ldg_stream_f64(data + k0 + lane + i * 32, pol_stream);
#pragma unroll
            for (int i = 0; i < EPL; ++i) xv[i] = ldg_f64_hint(x + c[i], polx);
            if (tnext < n_tiles) {
                r0n = tile_row[tnext];
                r1n = tile_row[tnext + 1];
            }
            if (regpath) {
#pragma unroll
                for (int i = 0; i < EPL; ++i) preg[i] = __dmul_rn(v[i], xv[i]);
            } else {
#pragma unroll
                for (int i = 0; i < EPL; ++i) sprod[lane + i * 32] = __dmul_rn(v[i], xv[i]);
            }
        } else if (full) {
            mbar_wait(&bars[warp][s], (phases >> s) & 1u);
            phases ^= 1u << s;
            uint32_t c[EPL];
            double xv[EPL];
#pragma unroll
            for (int i = 0; i < EPL; ++i) c[i] = sidx[lane + i * 32];
#pragma unroll
            for (int i = 0; i < EPL; ++i) xv[i] = ldg_f64_hint(x + c[i], polx);
            if (tnext < n_tiles) {  // next tile's row range: in flight behind the gathers
                r0n = tile_row[tnext];
                r1n = tile_row[tnext + 1];
            }
            if (regpath) {
#pragma unroll
                for (int i = 0; i < EPL; ++i) preg[i] = __dmul_rn(sprod[lane + i * 32], xv[i]);
            } else {
#pragma unroll
                for (int i = 0; i < EPL; ++i)
                    sprod[lane + i * 32] = __dmul_rn(sprod[lane + i * 32], xv[i]);
            }
        } else {  // ragged last tile: guarded loads, no bulk copy past the arrays
            for (int e = lane; e < cnt; e += 32)
                sprod[e] = __dmul_rn(data[k0 + e], ldg_f64_hint(x + indices[k0 + e], polx));
            if (tnext < n_tiles) {
                r0n = tile_row[tnext];
                r1n = tile_row[tnext + 1];
            }
        }
        __syncwarp();
        uint64_t b_next = 0;
        if (tnext < n_tiles) {  // boundaries of the next tile: in flight behind the reduction
            const uint64_t rln = (r1n < rows) ? (uint64_t)r1n : (uint64_t)r1n - 1;
            const uint64_t rr = (uint64_t)r0n + lane;
            b_next = rr <= rln + 1 ? (uint64_t)indptr[rr] : 0;
        }

        TileCtx tc;
        tc.k0 = k0;
        tc.k1 = k1;
        tc.r1 = r1;
        tc.y = yt.p[0];
        tc.yt = &yt;
        tc.carry_slot = carry + t;
        tc.accumulate = accumulate;
        const uint32_t avg = (uint32_t)((uint64_t)cnt / (r_last - r0 + 1));
        if (regpath) {
            // tile-local row boundaries: lane L holds clamp(indptr[r0+L] - k0, 0, WT), WT beyond
            int bl = WT;
            if (lane <= nrows_t) {
                const uint64_t bb = b_first > k0 ? b_first - k0 : 0;
                bl = bb < (uint64_t)WT ? (int)bb : WT;
            }
            reduce_rows_regs<EPL, MULTI>(tc, preg, bl, r0, 0, nrows_t, lane);
        } else if (avg <= 6)
            reduce_rows_warp<P, 1, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
        else if (avg <= 12)
            reduce_rows_warp<P, 2, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
        else if (avg <= 24)
            reduce_rows_warp<P, 4, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
        else if (avg <= 48)
            reduce_rows_warp<P, 8, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
        else if (avg <= 96)
            reduce_rows_warp<P, 16, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
        else
            reduce_rows_warp<P, 32, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
        __syncwarp();
        // refill this stage (generic-proxy accesses above must be ordered before the
        // async-proxy writes of the next bulk copy)
        const uint64_t tn = dyn ? tnext : t + (uint64_t)NST * GW;
        if (!DIRECT && lane == 0 && tn < n_tiles) {
            fence_proxy_async();
            SPMV_ISSUE(tn, s);
        }
        if (SIGNAL && progress) {
            ++sig_count;
            if (tnext >= n_tiles || (tnext >> chunk_shift) != (t >> chunk_shift)) {
                __threadfence();  // every lane: its y stores are visible device-wide ...
                __syncwarp();     // ... before lane 0 publishes the count
                if (lane == 0) {
                    __threadfence();  // release by the publishing thread itself (cumulative
                                      // over what the barrier made visible to it)
                    atomicAdd(&progress[t >> chunk_shift], sig_count);
                }
                sig_count = 0;
            }
        }
        s = (s + 1 == NST) ? 0 : s + 1;
        r0 = r0n;
        r1 = r1n;
        b_first = b_next;
        t = tnext;
        if (dyn) t_claimed = t_begin + __shfl_sync(0xffffffffu, claim, 0);
    }
}

#undef SPMV_ISSUE

// carries: tile t left the partial sum of row tile_row[t+1] in carry[t]; consecutive
// tiles with the same carry row form a run that is summed in tile order by its head.
__global__ void spmv_fixup_kernel(const uint32_t* __restrict__ tile_row,
                                  const double* __restrict__ carry, SpmvTargets yt,
                                  uint64_t n_tiles) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t + 1 >= n_tiles) return;
    const uint32_t row = tile_row[t + 1];
    if (t > 0 && tile_row[t] == row) return;  // not the head of its run
    // tile_row is sorted: the run [t, end) of tiles whose carry row is `row` ends at the
    // first u with tile_row[u + 1] > row.  Binary search, then a load-independent sum in
    // tile order (a hub row of 1e6 non-zeros is a run of ~4000 tiles: the old
    // load-compare-branch loop made this kernel 7 % of the step).
    uint64_t lo = t + 1, hi = n_tiles - 1;  // candidates for `end` (carry rows exist for u < n_tiles-1)
    for (uint64_t step = 1; lo + step < hi; step <<= 1) {  // gallop: most runs are 1-2 tiles
        if (tile_row[lo + step + 1] > row) {
            hi = lo + step;
            break;
        }
        lo += step;  // tiles lo .. lo+step still carry `row`... see invariant below
    }
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (tile_row[mid + 1] > row)
            hi = mid;
        else
            lo = mid + 1;
    }
    double sum = carry[t];
#pragma unroll 8
    for (uint64_t u = t + 1; u < lo; ++u) sum = __dadd_rn(sum, carry[u]);
    const double v = __dadd_rn(yt.p[0][row], sum);
#pragma unroll
    for (int q = 0; q < SPMV_MAX_TARGETS; ++q)
        if (q < yt.n) yt.p[q][row] = v;
}

// ---- launch configuration ---------------------------------------------------------
struct SpmvVariant {
    int wt, stages, nwarps, ctas_per_sm;
};
// default picked from the round-1 sweep (profiles/r1_spmv_variants.md);
// SPRS_B200_SPMV_VARIANT="wt,stages,nwarps,ctas" overrides it for tuning runs.
SpmvVariant spmv_variant() {
    static SpmvVariant v = [] {
        SpmvVariant d{384, 1, 8, 3};
        if (const char* e = getenv("SPRS_B200_SPMV_VARIANT")) {
            int a, b, c, g;
            if (sscanf(e, "%d,%d,%d,%d", &a, &b, &c, &g) == 4) d = SpmvVariant{a, b, c, g};
        }
        return d;
    }();
    return v;
}

// SPRS_B200_SPMV_DYNAMIC: unset = tiles are handed out dynamically in the pipelined launches
// (stream push, chunked push, chunked host path: kernels that share SMs with a put kernel);
// 1 = in every single-target SpMV; 0 = nowhere (the static round-robin deal).
int spmv_dynamic_mode() {
    static const int mode = [] {
        const char* e = getenv("SPRS_B200_SPMV_DYNAMIC");
        return e ? atoi(e) : -1;
    }();
    return mode;
}

struct SpmvSignal {  // progress counters of the pipelined all-gather; null = no signalling
    unsigned long long* progress = nullptr;
    int chunk_shift = 0;
    uint64_t t0 = 0, t1 = 0;  // tile range of this launch; t1 == 0: the whole matrix
    // dynamic tile hand-out (needs the SIGNAL flavour; null = the static round-robin deal):
    // an 8-byte counter, zeroed on the launch's stream right before the kernel
    unsigned long long* tile_counter = nullptr;
};

template <typename P, int WT, int STAGES, int NWARPS, int CTAS>
int launch_variant(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x,
                   const SpmvTargets& yt, int accumulate, const SpmvSignal& sig,
                   cudaStream_t s) {
    // CTAS resident CTAs per SM is also the kernel's __launch_bounds__ minBlocks: it sets the
    // register budget (ptxas otherwise picks ~40 registers and spills the gather buffers).
    // the hand-out lives in the single-target SIGNAL flavour only
    unsigned long long* const tile_counter = yt.n <= 1 ? sig.tile_counter : nullptr;
    const int flavour = (sig.progress || tile_counter) ? 2 : (yt.n > 1 ? 1 : 0);
    auto kern = flavour == 2   ? spmv_warp_kernel<P, WT, STAGES, NWARPS, CTAS, false, true>
                : flavour == 1 ? spmv_warp_kernel<P, WT, STAGES, NWARPS, CTAS, true, false>
                               : spmv_warp_kernel<P, WT, STAGES, NWARPS, CTAS, false, false>;
    const size_t smem = STAGES == 0 ? (size_t)NWARPS * WT * 8 : (size_t)NWARPS * STAGES * WT * 12;
    // function attributes are per device: remember them per (device, kernel flavour)
    static bool configured_flags[64][3] = {};
    bool& configured = configured_flags[ctx->device & 63][flavour];
    if (!configured) {
        SPRS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        // Shared-memory carve-out = exactly what CTAS resident CTAs need, everything else
        // stays L1: every in-flight gather holds an L1 line, so the gather rate is bounded
        // by L1 lines / L2 latency (profiles/r1_spmv_notes.md).  A max-shared carve-out
        // cost 40 % of the throughput; too small a carve-out would drop resident CTAs.
        int carve = (int)(((smem + 1024) * CTAS * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (const char* e = getenv("SPRS_B200_SPMV_CARVEOUT")) carve = atoi(e);
        if (carve > 100) carve = 100;
        SPRS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            carve));
        configured = true;
    }
    const uint64_t t0 = sig.t1 ? sig.t0 : 0, t1 = sig.t1 ? sig.t1 : m->n_tiles;
    uint64_t grid = (uint64_t)ctx->sm_count * CTAS;
    const uint64_t need = (t1 - t0 + NWARPS - 1) / NWARPS;
    if (grid > need) grid = need;
    if (tile_counter) SPRS_CUDA(ctx, cudaMemsetAsync(tile_counter, 0, 8, s));
    kern<<<(unsigned)grid, NWARPS * 32, smem, s>>>((const P*)m->d_indptr, m->d_indices, m->d_data,
                                                   m->d_tile_row, d_x, yt, m->d_carry, m->nnz,
                                                   (uint32_t)m->rows, t0, t1, accumulate,
                                                   sig.progress, sig.chunk_shift,
                                                   tile_counter);
    return SPRS_B200_OK;
}

template <typename P>
int launch_dispatch(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x,
                    const SpmvTargets& yt, int accumulate, const SpmvSignal& sig,
                    cudaStream_t s) {
    const SpmvVariant v = spmv_variant();
#define SPMV_CASE(WT, ST, NW, CT)                                                         \
    if (v.wt == WT && v.stages == ST && v.nwarps == NW && v.ctas_per_sm == CT)            \
        return launch_variant<P, WT, ST, NW, CT>(ctx, m, d_x, yt, accumulate, sig, s);
    SPMV_CASE(256, 0, 8, 3)
    SPMV_CASE(256, 2, 8, 3)
    SPMV_CASE(256, 2, 8, 2)
    SPMV_CASE(256, 1, 8, 3)
    SPMV_CASE(256, 1, 8, 4)
    SPMV_CASE(384, 1, 8, 3)
    SPMV_CASE(512, 1, 8, 2)
    SPMV_CASE(256, 0, 8, 4)
    SPMV_CASE(256, 0, 16, 2)
    SPMV_CASE(384, 0, 8, 3)
    SPMV_CASE(512, 0, 8, 2)
    SPMV_CASE(512, 0, 8, 3)
    SPMV_CASE(256, 0, 8, 2)
#undef SPMV_CASE
    SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "unknown SPRS_B200_SPMV_VARIANT");
}

}  // namespace

int spmv_prepare(sprs_b200_ctx* ctx, sprs_b200_csmat* m, cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR) return SPRS_B200_OK;  // CSC mirrors are converted first
    const uint32_t wt = (uint32_t)spmv_variant().wt;
    m->n_tiles = m->nnz == 0 ? 1 : (m->nnz + wt - 1) / wt;
    SPRS_CUDA(ctx, cudaMalloc((void**)&m->d_tile_row, (m->n_tiles + 1) * sizeof(uint32_t)));
    SPRS_CUDA(ctx, cudaMalloc((void**)&m->d_carry, m->n_tiles * sizeof(double)));
    const uint64_t n = m->n_tiles + 1;
    const unsigned grid = (unsigned)((n + 255) / 256);
    if (m->indptr_bytes == 4)
        tile_row_kernel<uint32_t><<<grid, 256, 0, s>>>((const uint32_t*)m->d_indptr,
                                                       (uint32_t)m->rows, m->n_tiles, wt,
                                                       m->d_tile_row);
    else
        tile_row_kernel<uint64_t><<<grid, 256, 0, s>>>((const uint64_t*)m->d_indptr,
                                                       (uint32_t)m->rows, m->n_tiles, wt,
                                                       m->d_tile_row);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

int spmv_launch_targets(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x,
                        const SpmvTargets& yt, int accumulate, cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR)
        SPRS_FAIL(ctx, SPRS_B200_ERR_STORAGE, "Storage mismatch: spmv needs a CSR mirror");
    if (m->rows == 0) return SPRS_B200_OK;
    if (!m->d_tile_row) SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "csmat has no SpMV partition");
    SpmvSignal sig;
    if (spmv_dynamic_mode() == 1 && yt.n <= 1) SPRS_TRY(ctx_tile_counter(ctx, &sig.tile_counter));
    if (m->indptr_bytes == 4)
        SPRS_TRY(launch_dispatch<uint32_t>(ctx, m, d_x, yt, accumulate, sig, s));
    else
        SPRS_TRY(launch_dispatch<uint64_t>(ctx, m, d_x, yt, accumulate, sig, s));
    ctx->launches += 1;
    if (m->n_tiles > 1) {
        const unsigned fgrid = (unsigned)((m->n_tiles - 1 + 255) / 256);
        spmv_fixup_kernel<<<fgrid, 256, 0, s>>>(m->d_tile_row, m->d_carry, yt, m->n_tiles);
        ctx->launches += 1;
    }
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

// ---- pipelined all-gather: SpMV + concurrent put kernel ------------------------------
// The put kernel (a few CTAs on a high-priority side stream, resident BEFORE the SpMV
// starts) follows the SpMV's progress counters chunk by chunk.  Once every tile of chunks
// 0..c has been reduced, each put CTA takes a contiguous share [u_lo, u_hi) of the chunk's
// tiles, applies the carries of the rows that END in those tiles (so rows
// [tile_row[u_lo], tile_row[u_hi]) of the local y are final), and copies those rows into
// every other target buffer -- peer GPUs' y over NVLink, or a pinned host buffer over PCIe --
// with coalesced 8-byte stores, while the SpMV works on the later chunks.  Unlike the fused
// variant (peer stores from the SpMV's own epilogue) the remote traffic never sits in the LSU
// queues of the warps that gather x, and nothing is left to do after the SpMV but the last
// chunk.
namespace {

constexpr int PUT_THREADS = 512;
constexpr long long PUT_TIMEOUT_CYCLES = 6000000000ll;  // ~3 s: trap instead of hanging

// Carries of the row that ends in tile u (u >= 1): the run of tiles [t, u-1] whose carry
// row is tile_row[u] left partial sums in carry[]; they are added in tile order, exactly
// like spmv_fixup_kernel does from the head of the run (same bits).
__device__ __forceinline__ void apply_carries_ending_in(const uint32_t* __restrict__ tile_row,
                                                        const double* carry, double* y,
                                                        uint64_t u) {
    const uint32_t row = tile_row[u];
    if (tile_row[u + 1] == row) return;  // the row continues past tile u: not final yet
    // head of the run: first index f in [1, u] with tile_row[f] >= row, t = f - 1
    uint64_t hi = u, lo = 1;
    for (uint64_t step = 1; step < hi; step <<= 1) {  // gallop down: most runs are 1 tile
        if (tile_row[hi - step] < row) {
            lo = hi - step + 1;
            break;
        }
        hi -= step;
    }
    while (lo < hi) {
        const uint64_t mid = lo + (hi - lo) / 2;
        if (tile_row[mid] >= row)
            hi = mid;
        else
            lo = mid + 1;
    }
    double sum = __ldcg(carry + lo - 1);
    for (uint64_t v = lo; v < u; ++v) sum = __dadd_rn(sum, __ldcg(carry + v));
    y[row] = __dadd_rn(__ldcg(y + row), sum);
}

__global__ void __launch_bounds__(PUT_THREADS)
    stream_put_kernel(double* y_own, SpmvTargets targets, const uint32_t* __restrict__ tile_row,
                      const double* carry, const unsigned long long* progress,
                      unsigned long long epoch, uint64_t n_tiles, int chunk_shift,
                      uint32_t n_chunks, uint32_t rows) {
    const uint64_t tpc = 1ull << chunk_shift;
    for (uint32_t c = 0; c < n_chunks; ++c) {
        const uint64_t t0 = (uint64_t)c << chunk_shift;
        const uint64_t t1 = t0 + tpc < n_tiles ? t0 + tpc : n_tiles;
        if (threadIdx.x == 0) {
            const unsigned long long want = epoch * (unsigned long long)(t1 - t0);
            const volatile unsigned long long* p = progress + c;
            const long long start = clock64();
            unsigned backoff = 64;
            while (*p < want) {
                __nanosleep(backoff);
                if (backoff < 2048) backoff <<= 1;
                if (clock64() - start > PUT_TIMEOUT_CYCLES) __trap();  // the SpMV never ran
            }
            __threadfence();  // acquire: the rows and carries counted above are visible
        }
        __syncthreads();
        // this CTA's share of the chunk's tiles (chunks are visited in order, so every tile
        // before u_hi is complete)
        const uint64_t span = t1 - t0;
        const uint64_t u_lo = t0 + span * blockIdx.x / gridDim.x;
        const uint64_t u_hi = t0 + span * (blockIdx.x + 1) / gridDim.x;
        for (uint64_t u = u_lo + threadIdx.x; u < u_hi; u += PUT_THREADS)
            if (u >= 1) apply_carries_ending_in(tile_row, carry, y_own, u);
        __syncthreads();
        if (targets.n > 1) {
            const uint64_t lo = tile_row[u_lo];  // tile_row[0] == 0, tile_row[n_tiles] == rows
            const uint64_t hi = tile_row[u_hi];
            for (uint64_t i = lo + threadIdx.x; i < hi; i += 4 * PUT_THREADS) {
                double v[4];
#pragma unroll
                for (int w = 0; w < 4; ++w)
                    v[w] = (i + w * PUT_THREADS < hi) ? __ldcg(y_own + i + w * PUT_THREADS) : 0.0;
#pragma unroll
                for (int q = 1; q < SPMV_MAX_TARGETS; ++q)
                    if (q < targets.n) {
#pragma unroll
                        for (int w = 0; w < 4; ++w)
                            if (i + w * PUT_THREADS < hi) targets.p[q][i + w * PUT_THREADS] = v[w];
                    }
            }
        }
        __syncthreads();
    }
    (void)rows;
}

int stream_push_prepare(sprs_b200_ctx* ctx, sprs_b200_csmat* m, cudaStream_t s) {
    SPRS_TRY(ctx_side_stream(ctx));
    if (!m->d_progress) {
        // about 12 chunks: the last chunk's push is what cannot overlap
        int shift = 0;
        while ((m->n_tiles >> shift) > 12) ++shift;
        if (const char* e = getenv("SPRS_B200_PUSH_CHUNK_SHIFT")) shift = atoi(e);
        if (shift < 0) shift = 0;
        if (shift > 40) shift = 40;
        m->chunk_shift = shift;
        m->n_chunks = (uint32_t)((m->n_tiles + (1ull << shift) - 1) >> shift);
        SPRS_CUDA(ctx, cudaMalloc((void**)&m->d_progress, (size_t)m->n_chunks * 8));
        // on the caller's stream: ordered before the fork event both kernels wait behind
        SPRS_CUDA(ctx, cudaMemsetAsync(m->d_progress, 0, (size_t)m->n_chunks * 8, s));
        m->push_epoch = 0;
    }
    return SPRS_B200_OK;
}

}  // namespace

int spmv_launch_stream_push(sprs_b200_ctx* ctx, sprs_b200_csmat* m, const double* d_x,
                            const SpmvTargets& yt, int accumulate, int put_ctas,
                            cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR)
        SPRS_FAIL(ctx, SPRS_B200_ERR_STORAGE, "Storage mismatch: spmv needs a CSR mirror");
    if (m->rows == 0) return SPRS_B200_OK;
    if (!m->d_tile_row) SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "csmat has no SpMV partition");
    if (yt.n <= 1 && !getenv("SPRS_B200_PUSH_ALWAYS"))  // nothing to push to: the plain path
        return spmv_launch_targets(ctx, m, d_x, yt, accumulate, s);
    SPRS_TRY(stream_push_prepare(ctx, m, s));
    m->push_epoch += 1;
    SpmvTargets own;
    own.n = 1;
    own.p[0] = yt.p[0];
    for (int q = 1; q < SPMV_MAX_TARGETS; ++q) own.p[q] = nullptr;
    SpmvSignal sig;
    sig.progress = m->d_progress;
    sig.chunk_shift = m->chunk_shift;
    if (spmv_dynamic_mode() != 0) SPRS_TRY(ctx_tile_counter(ctx, &sig.tile_counter));
    if (put_ctas <= 0) put_ctas = 16;
    if (const char* e = getenv("SPRS_B200_PUSH_CTAS")) put_ctas = atoi(e) > 0 ? atoi(e) : put_ctas;
    if (put_ctas > ctx->sm_count) put_ctas = ctx->sm_count;
    auto launch_put = [&]() -> int {
        stream_put_kernel<<<(unsigned)put_ctas, PUT_THREADS, 0, ctx->side_stream>>>(
            yt.p[0], yt, m->d_tile_row, m->d_carry, m->d_progress,
            (unsigned long long)m->push_epoch, m->n_tiles, m->chunk_shift, m->n_chunks,
            (uint32_t)m->rows);
        ctx->launches += 1;
        SPRS_CUDA(ctx, cudaGetLastError());
        return SPRS_B200_OK;
    };
    auto launch_spmv = [&]() -> int {
        if (m->indptr_bytes == 4)
            SPRS_TRY(launch_dispatch<uint32_t>(ctx, m, d_x, own, accumulate, sig, s));
        else
            SPRS_TRY(launch_dispatch<uint64_t>(ctx, m, d_x, own, accumulate, sig, s));
        ctx->launches += 1;
        SPRS_CUDA(ctx, cudaGetLastError());
        return SPRS_B200_OK;
    };
    // fork: the put kernel starts once the caller's stream has reached this point (y and the
    // peers' buffers are free to be overwritten) -- and it must be RESIDENT before the SpMV
    // fills every SM, hence put first.  (The CPU emulator of tests/emu runs kernels one after
    // the other, so there the order is swapped: the counters are complete when the put runs.)
    SPRS_CUDA(ctx, cudaEventRecord(ctx->ev_fork, s));
    SPRS_CUDA(ctx, cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
#ifdef CUEMU
    SPRS_TRY(launch_spmv());
    SPRS_TRY(launch_put());
#else
    SPRS_TRY(launch_put());
    SPRS_TRY(launch_spmv());
#endif
    // join: the caller's stream continues when the last chunk has been fixed up and pushed
    SPRS_CUDA(ctx, cudaEventRecord(ctx->ev_join, ctx->side_stream));
    SPRS_CUDA(ctx, cudaStreamWaitEvent(s, ctx->ev_join, 0));
    return SPRS_B200_OK;
}

// ---- one chunk of the tile stream (pipelined host path of api.cu) ---------------------
namespace {
__global__ void spmv_fixup_range_kernel(const uint32_t* __restrict__ tile_row, const double* carry,
                                        double* y, uint64_t u_lo, uint64_t u_hi) {
    const uint64_t u = u_lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (u >= 1 && u < u_hi) apply_carries_ending_in(tile_row, carry, y, u);
}
}  // namespace

// SpMV over tiles [t0, t1) followed by the carries of the rows that END in those tiles: once
// this has run for every tile below t1 (chunks in increasing order on one stream), rows
// [0, tile_row[t1]) of y are final -- the same sums in the same order as the one-shot
// spmv_launch (apply_carries_ending_in adds a run's carries in tile order from its head, like
// spmv_fixup_kernel).
int spmv_launch_tile_range(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x,
                           double* d_y, int accumulate, uint64_t t0, uint64_t t1,
                           cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR)
        SPRS_FAIL(ctx, SPRS_B200_ERR_STORAGE, "Storage mismatch: spmv needs a CSR mirror");
    if (!m->d_tile_row) SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "csmat has no SpMV partition");
    if (t1 > m->n_tiles || t0 >= t1) SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "bad tile range");
    SpmvTargets yt;
    yt.n = 1;
    yt.p[0] = d_y;
    for (int q = 1; q < SPMV_MAX_TARGETS; ++q) yt.p[q] = nullptr;
    SpmvSignal sig;
    sig.t0 = t0;
    sig.t1 = t1;
    if (spmv_dynamic_mode() != 0) SPRS_TRY(ctx_tile_counter(ctx, &sig.tile_counter));
    if (m->indptr_bytes == 4)
        SPRS_TRY(launch_dispatch<uint32_t>(ctx, m, d_x, yt, accumulate, sig, s));
    else
        SPRS_TRY(launch_dispatch<uint64_t>(ctx, m, d_x, yt, accumulate, sig, s));
    spmv_fixup_range_kernel<<<(unsigned)((t1 - t0 + 255) / 256), 256, 0, s>>>(
        m->d_tile_row, m->d_carry, d_y, t0, t1);
    ctx->launches += 2;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

int spmv_launch(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x, double* d_y,
                int accumulate, cudaStream_t s) {
    SpmvTargets yt;
    yt.n = 1;
    yt.p[0] = d_y;
    for (int q = 1; q < SPMV_MAX_TARGETS; ++q) yt.p[q] = nullptr;
    return spmv_launch_targets(ctx, m, d_x, yt, accumulate, s);
}
