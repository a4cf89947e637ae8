// spmv.cu -- CSR x dense-vector product for sm_100a (B200).
//
// Replaces prod::mul_acc_mat_vec_csr (sprs/src/sparse/prod.rs:103-127) and the
// one-column case of prod::csr_mulacc_dense_colmaj (prod.rs:274-298), which is what
// `&A * &x` runs (sprs/src/sparse/csmat.rs:2142-2148).
//
// Design (evidence in profiles/r2_spmv_notes.md).  The kernel is bound by the
// L1TEX pipe that serves the x gathers (one 128-byte line per clock per SM) and by the L1 lines
// those gathers hold while their sectors are in flight, not by HBM: the "ceiling" kernels of
// tools/spmv_lab.cu (same streams and gathers, no row logic) top out at 0.55-0.57 of the HBM
// roofline on the 10M R-MAT and 0.45 on uniform columns.  So the kernel is organised to keep
// gathers in flight ALL the time and to spend as few other L1TEX/MIO operations as possible:
//   * the nnz stream is cut into tiles of WT = 32*EPL non-zeros (not rows); a tile belongs to
//     ONE WARP, warps are persistent and autonomous (no CTA-wide barrier anywhere);
//   * no shared-memory staging: indices and values stream through registers with coalesced
//     ld.global.nc.L1::no_allocate (L2 evict_first: the matrix is read exactly once), which
//     leaves the whole L1 to the gathers (a max-shared carve-out costs 3x, lab carve sweep);
//   * software pipeline, one tile deep: the gathers and value loads of tile t+1 (indices
//     prefetched during tile t-1) are issued BEFORE tile t is reduced, so the reduction of a
//     tile -- the part that used to serialise behind the gathers -- runs under the next
//     tile's memory latency; row ranges (tile_row) are fetched two tiles ahead, the row
//     boundaries (indptr) one tile ahead;
//   * register reduction for tiles touching <= 24 rows: lane L holds elements L + 32*i, i.e.
//     32 consecutive non-zeros per register "slab"; slabs are walked in order, a slab without
//     a row end costs one add, a row end inside a slab splits the lanes (warp-uniform control
//     flow), finished per-lane row partials are parked in 4 slots and reduced four rows at a
//     time by one multi-value butterfly (12 shuffles for 4 rows).  Products never touch shared
//     memory; y is written once;
//   * tiles with more rows (short rows, runs of empty rows) store their products to a small
//     per-warp shared buffer and reduce rows with lane groups of G = 1..32 lanes; rows of tiles
//     with mean length <= 6 are summed by one lane in storage order, i.e. bit-identical to the
//     reference's sequential sum; longer rows use trees and agree to rounding (parity gate:
//     |d| <= 1e-6 * sum|terms|);
//   * the row cut by the tile end leaves its partial in carry[t]; a second tiny kernel adds the
//     carries in tile order (deterministic, no atomics).
// Arithmetic is MulAcc::mul_acc's (mul_acc.rs:28-30): unfused multiply, then add.
//
// Algorithmic bytes per nnz: 12 (8 data + 4 index) + 8 per row (y) -- the BASELINE roofline
// 12*nnz + 8*n; indptr (4 B/row), tile_row/carry (12 B per tile) and x gathers are overhead.

#include "common.cuh"
#include "ptx.cuh"

#include <cstdlib>

namespace {

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

constexpr int SPMV_REG_ROWS = 24;  // tiles completing <= this many rows reduce in registers

struct TileCtx {
    uint64_t k0, k1;
    uint32_t r1;  // first row NOT owned (== carry row when < rows)
    double* y;    // this GPU's y (target 0)
    const SpmvTargets* yt;  // MULTI only: the kernel parameter itself (constant bank)
    double* carry_slot;
    int accumulate;
};

// y[r] is written to every target buffer: target 0 is this GPU's own y; targets 1.. are the
// peer GPUs' y buffers (CUDA IPC / VMM mappings) or the NVSwitch multicast address of y (fused
// SpMV + all-gather over NVLink: the result of a row leaves for the peers the moment it is
// reduced, overlapped with the rest of the kernel, instead of a separate collective afterwards).
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

// Shared-memory path: reduce rows [r0, r_last] of one warp tile with groups of G lanes per
// row.  Row boundaries come 31 rows at a time: lane L holds indptr[rbase + L].
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

// What the row emitters need, by value (they are real functions, not inlined: the hot loop stays
// small enough for the instruction cache).
struct RowSink {
    double* y;                  // this GPU's y (target 0)
    const SpmvTargets* yt;      // MULTI only: the kernel parameter itself (constant bank)
    double* carry_slot;
    uint32_t r1;                // first row NOT owned by the tile (== its carry row when < rows)
    int accumulate;
};
template <bool MULTI>
__device__ __forceinline__ void sink_row(const RowSink& k, uint64_t r, double sum) {
    if (r < k.r1) {
        const double v = k.accumulate ? __dadd_rn(k.y[r], sum) : sum;
        k.y[r] = v;
        if (MULTI) {
#pragma unroll
            for (int q = 1; q < SPMV_MAX_TARGETS; ++q)
                if (q < k.yt->n) k.yt->p[q][r] = v;
        }
    } else {
        *k.carry_slot = sum;  // row continues in a later tile: spmv_fixup_kernel adds it
    }
}

// Register path, second half: up to four finished rows sit as per-lane partials in s0..s3
// (row_base + 0..3).  One multi-value butterfly reduces them together: xor 16 halves the live
// values (lanes with bit 4 clear keep rows 0,1, the others rows 2,3), xor 8 halves again, xor
// 4/2/1 finish: 6 double shuffles for 4 rows instead of 20.  Lane 8*j ends up with row j.
template <bool MULTI>
__device__ __noinline__ void flush_slots(RowSink k, double s0, double s1, double s2, double s3,
                                         uint32_t row_base, int n, int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const bool up16 = lane & 16, up8 = lane & 8;
    const double a0 = __dadd_rn(up16 ? s2 : s0, __shfl_xor_sync(FULL, up16 ? s0 : s2, 16));
    const double a1 = __dadd_rn(up16 ? s3 : s1, __shfl_xor_sync(FULL, up16 ? s1 : s3, 16));
    double v = __dadd_rn(up8 ? a1 : a0, __shfl_xor_sync(FULL, up8 ? a0 : a1, 8));
    v = __dadd_rn(v, __shfl_xor_sync(FULL, v, 4));
    v = __dadd_rn(v, __shfl_xor_sync(FULL, v, 2));
    v = __dadd_rn(v, __shfl_xor_sync(FULL, v, 1));
    const int row = lane >> 3;  // 2*bit4 + bit3
    if ((lane & 7) == 0 && row < n) sink_row<MULTI>(k, (uint64_t)row_base + row, v);
}

// Register path: the tile's EPL products per lane (element e = lane + 32*i, i.e. register i is
// the "slab" of 32 consecutive non-zeros 32*i .. 32*i+31) are folded into row sums without
// leaving the registers.  ends: lane L in 1..nrc holds the tile-local END of row r0+L-1, in
// [0, WT]; nrc rows end inside the tile, what is left after the last end belongs to row r0+nrc
// (the tile's carry) when has_tail.  All control flow is warp-uniform AND known to be: the
// slab mask comes from a warp reduction (REDUX), row counts from ballots, so the branches cost
// no divergence bookkeeping.  A slab without a row end costs one add.
template <int EPL, bool MULTI>
__device__ __forceinline__ void reduce_rows_slots(const RowSink& k, const double (&p)[EPL],
                                                  int end_local, uint32_t r0, int nrc, bool has_tail,
                                                  int lane) {
    constexpr unsigned FULL = 0xffffffffu;
    const bool is_end = lane >= 1 && lane <= nrc;
    // slab that holds the row's last element (end 0 = an empty row at the very start: slab 0)
    const int sb = end_local > 0 ? (end_local - 1) >> 5 : 0;
    const int q = end_local - 32 * sb;  // lanes < q of slab sb belong to the row (0 .. 32)
    const unsigned slabs = __reduce_or_sync(FULL, is_end ? 1u << sb : 0u);
    double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0, acc = 0.0;
    int j = 0;
    uint32_t row_base = r0;
#define SPMV_PARK(val)                                                  \
    do {                                                                \
        const double v_ = (val);                                        \
        if (j == 0) s0 = v_;                                            \
        else if (j == 1) s1 = v_;                                       \
        else if (j == 2) s2 = v_;                                       \
        else s3 = v_;                                                   \
        if (++j == 4) {                                                 \
            flush_slots<MULTI>(k, s0, s1, s2, s3, row_base, 4, lane);   \
            row_base += 4;                                              \
            j = 0;                                                      \
        }                                                               \
    } while (0)
#pragma unroll
    for (int i = 0; i < EPL; ++i) {
        double pi = p[i];
        if ((slabs >> i) & 1u) {  // rows end inside this slab: consecutive boundary lanes
            const unsigned m = __ballot_sync(FULL, is_end && sb == i);
            const int first = __ffs(m) - 1, cnt = __popc(m);
            for (int c = 0; c < cnt; ++c) {
                const bool mine = lane < __shfl_sync(FULL, q, first + c);
                SPMV_PARK(__dadd_rn(acc, mine ? pi : 0.0));
                pi = mine ? 0.0 : pi;
                acc = 0.0;
            }
        }
        acc = __dadd_rn(acc, pi);
    }
    if (has_tail) SPMV_PARK(acc);  // row r0+nrc >= r1: sink_row files it as the tile's carry
    if (j > 0) flush_slots<MULTI>(k, s0, j > 1 ? s1 : 0.0, j > 2 ? s2 : 0.0, 0.0, row_base, j, lane);
#undef SPMV_PARK
}

// Shared-memory path (tiles with many rows, and the ragged last tile): products are in the
// warp's buffer; rows are reduced by lane groups sized to the tile's mean row length.
template <typename P, bool MULTI>
__device__ __noinline__ void reduce_rows_smem(RowSink k, const P* __restrict__ indptr,
                                              const double* sprod, uint64_t k0, uint64_t k1,
                                              uint32_t r0, uint32_t rows, uint64_t b_first, int lane) {
    TileCtx tc;
    tc.k0 = k0;
    tc.k1 = k1;
    tc.r1 = k.r1;
    tc.y = k.y;
    tc.yt = k.yt;
    tc.carry_slot = k.carry_slot;
    tc.accumulate = k.accumulate;
    const uint64_t r_last = k.r1 < rows ? (uint64_t)k.r1 : (uint64_t)k.r1 - 1;
    const uint64_t cnt = k1 - k0, nr = r_last - r0 + 1;
    if (cnt <= 6 * nr)
        reduce_rows_warp<P, 1, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
    else if (cnt <= 12 * nr)
        reduce_rows_warp<P, 2, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
    else
        reduce_rows_warp<P, 4, MULTI>(tc, indptr, sprod, r0, r_last, b_first, lane);
}

template <typename P, int EPL, int NWARPS, int MINB, bool MULTI>
__global__ void __launch_bounds__(NWARPS * 32, MINB)
    spmv_pipe_kernel(const P* __restrict__ indptr, const uint32_t* __restrict__ indices,
                     const double* __restrict__ data, const uint32_t* __restrict__ tile_row,
                     const double* __restrict__ x, const __grid_constant__ SpmvTargets yt,
                     double* __restrict__ carry, uint64_t nnz, uint32_t rows, uint32_t t_begin,
                     uint32_t t_end /* this launch covers tiles [t_begin, t_end) */, int accumulate,
                     uint64_t pol_stream /* L2 evict_first */, uint64_t polx /* L2 evict_last */) {
    constexpr int WT = EPL * 32;
    using OFF = P;  // element offsets: 32 bits wide when the indptr is (nnz < 2^32)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sprod = (double*)smem_raw + (size_t)warp * WT;  // many-row tiles only
    const uint32_t GW = gridDim.x * NWARPS;
    uint32_t t = t_begin + blockIdx.x * NWARPS + warp;
    // tiles below n_full are whole: the hot loop never guards a load.  The ragged tail (at most
    // one tile, index n_full) is handled after the loop by the warp the deal gives it to.
    const uint32_t n_full = (uint32_t)(nnz / WT);
    const uint32_t t_hot_end = t_end < n_full ? t_end : n_full;
    // (the two L2 policies are kernel PARAMETERS: values the compiler knows to be warp-uniform
    // go into uniform registers once; as per-thread createpolicy results every hinted load
    // re-materialised its descriptor -- 32 extra instructions per tile)
    RowSink sink;
    sink.y = yt.p[0];
    sink.yt = &yt;
    sink.accumulate = accumulate;

    if (t < t_hot_end) {
        uint32_t c[EPL];
        double xn[EPL], vn[EPL];
        // prologue: tile t fully issued, tile t+GW's indices and row range on their way
        uint32_t tn = t + GW;
        uint32_t r0 = tile_row[t], r1 = tile_row[t + 1];
        uint32_t r0n = 0, r1n = 0;
        if (tn < t_hot_end) {
            r0n = tile_row[tn];
            r1n = tile_row[tn + 1];
        }
        {
            const uint32_t* ip = indices + (OFF)t * WT + lane;
#pragma unroll
            for (int i = 0; i < EPL; ++i) c[i] = ldg_stream_u32(ip + 32 * i, pol_stream);
        }
        // lane L: indptr[r0 + L] (L <= rows ending in the tile, +1 when a row continues)
        P b_first = (r0 + lane <= (r1 < rows ? r1 + 1 : r1)) ? indptr[(size_t)r0 + lane] : (P)0;
        {
            const double* dp = data + (OFF)t * WT + lane;
#pragma unroll
            for (int i = 0; i < EPL; ++i) xn[i] = ldg_f64_hint(x + c[i], polx);
#pragma unroll
            for (int i = 0; i < EPL; ++i) vn[i] = ldg_stream_f64(dp + 32 * i, pol_stream);
        }
        if (tn < t_hot_end) {
            const uint32_t* ip = indices + (OFF)tn * WT + lane;
#pragma unroll
            for (int i = 0; i < EPL; ++i) c[i] = ldg_stream_u32(ip + 32 * i, pol_stream);
        }
        for (;;) {
            // products of tile t (its gathers and values were issued one iteration ago)
            double p[EPL];
#pragma unroll
            for (int i = 0; i < EPL; ++i) p[i] = __dmul_rn(vn[i], xn[i]);
            // tile t+GW: gathers + values in flight while tile t is reduced; then the indices of
            // the tile after that, its row range, and the row boundaries of tile t+GW
            const uint32_t tnn = tn + GW;
            uint32_t r0nn = 0, r1nn = 0;
            P b_next = 0;
            if (tn < t_hot_end) {
                const double* dp = data + (OFF)tn * WT + lane;
#pragma unroll
                for (int i = 0; i < EPL; ++i) xn[i] = ldg_f64_hint(x + c[i], polx);
#pragma unroll
                for (int i = 0; i < EPL; ++i) vn[i] = ldg_stream_f64(dp + 32 * i, pol_stream);
                b_next = (r0n + lane <= (r1n < rows ? r1n + 1 : r1n)) ? indptr[(size_t)r0n + lane] : (P)0;
                if (tnn < t_hot_end) {
                    const uint32_t* ip = indices + (OFF)tnn * WT + lane;
#pragma unroll
                    for (int i = 0; i < EPL; ++i) c[i] = ldg_stream_u32(ip + 32 * i, pol_stream);
                    r0nn = tile_row[tnn];
                    r1nn = tile_row[tnn + 1];
                }
            }
            const OFF k0 = (OFF)t * WT;
            sink.carry_slot = carry + t;
            sink.r1 = r1;
            const uint32_t nrc = r1 - r0;  // rows that END in this tile
            const bool has_tail = r1 < rows;
            // mean row length <= 6 keeps the one-lane-per-row path (storage order: the
            // reference's bits); more than 24 row ends do not fit the boundary lanes
            if (nrc <= (uint32_t)SPMV_REG_ROWS && (uint32_t)WT > 6u * (nrc + (has_tail ? 1u : 0u))) {
                int end_local = 0;  // lane L in 1..nrc: end of row r0+L-1, clamped to the tile
                if (lane <= (int)nrc) {
                    const P bb = b_first > k0 ? (P)(b_first - k0) : (P)0;
                    end_local = bb < (P)WT ? (int)bb : WT;
                }
                reduce_rows_slots<EPL, MULTI>(sink, p, end_local, r0, (int)nrc, has_tail, lane);
            } else {
#pragma unroll
                for (int i = 0; i < EPL; ++i) sprod[lane + 32 * i] = p[i];
                __syncwarp();
                reduce_rows_smem<P, MULTI>(sink, indptr, sprod, (uint64_t)k0, (uint64_t)k0 + WT, r0,
                                           rows, (uint64_t)b_first, lane);
                __syncwarp();
            }
            t = tn;
            if (tn >= t_hot_end) break;
            tn = tnn;
            r0 = r0n;
            r1 = r1n;
            r0n = r0nn;
            r1n = r1nn;
            b_first = b_next;
        }
    }
    // the ragged tail tile (or the single empty tile of a matrix without non-zeros)
    if (t == n_full && t < t_end) {
        const uint64_t k0 = (uint64_t)t * WT;
        const uint32_t r0 = tile_row[t], r1 = tile_row[t + 1];
        for (uint64_t e = k0 + lane; e < nnz; e += 32)
            sprod[e - k0] = __dmul_rn(data[e], ldg_f64_hint(x + indices[e], polx));
        __syncwarp();
        sink.carry_slot = carry + t;
        sink.r1 = r1;
        const uint64_t rl = r1 < rows ? (uint64_t)r1 : (uint64_t)r1 - 1;
        const uint64_t rr = (uint64_t)r0 + lane;
        const uint64_t b_first = rr <= rl + 1 ? (uint64_t)indptr[rr] : 0;
        reduce_rows_smem<P, MULTI>(sink, indptr, sprod, k0, nnz, r0, rows, b_first, lane);
    }
}

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
    // first u with tile_row[u + 1] > row.  Gallop + binary search, then a load-independent sum
    // in tile order (a hub row of 1e6 non-zeros is a run of thousands of tiles).
    uint64_t lo = t + 1, hi = n_tiles - 1;  // candidates for `end` (carry rows exist for u < n_tiles-1)
    for (uint64_t step = 1; lo + step < hi; step <<= 1) {
        if (tile_row[lo + step + 1] > row) {
            hi = lo + step;
            break;
        }
        lo += step;
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

__global__ void spmv_fixup_range_kernel(const uint32_t* __restrict__ tile_row, const double* carry,
                                        double* y, uint64_t u_lo, uint64_t u_hi) {
    const uint64_t u = u_lo + blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (u >= 1 && u < u_hi) apply_carries_ending_in(tile_row, carry, y, u);
}

// ---- launch configuration ---------------------------------------------------------
struct SpmvVariant {
    int epl, ctas_per_sm;
};
// default picked from the round-2 sweeps (profiles/r2_spmv_notes.md);
// SPRS_B200_SPMV_VARIANT="epl,ctas" overrides it for tuning runs (read once per process: the
// tile size 32*epl is baked into every mirror's tile_row).
SpmvVariant spmv_variant() {
    static SpmvVariant v = [] {
        SpmvVariant d{8, 2};
        if (const char* e = getenv("SPRS_B200_SPMV_VARIANT")) {
            int a, b;
            if (sscanf(e, "%d,%d", &a, &b) == 2) d = SpmvVariant{a, b};
        }
        return d;
    }();
    return v;
}

constexpr int SPMV_NWARPS = 8;

template <typename P, int EPL, int CTAS>
int launch_variant(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x,
                   const SpmvTargets& yt, int accumulate, uint64_t t0, uint64_t t1,
                   cudaStream_t s) {
    if (m->n_tiles >= 0xffffffffull)
        SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "spmv: more than 2^32 tiles");
    // CTAS resident CTAs per SM is also the kernel's __launch_bounds__ minBlocks: it sets the
    // register budget of the pipelined operand buffers.
    const bool multi = yt.n > 1;
    auto kern = multi ? spmv_pipe_kernel<P, EPL, SPMV_NWARPS, CTAS, true>
                      : spmv_pipe_kernel<P, EPL, SPMV_NWARPS, CTAS, false>;
    const size_t smem = (size_t)SPMV_NWARPS * EPL * 32 * 8;
    static bool configured_flags[64][2] = {};  // function attributes are per device
    bool& configured = configured_flags[ctx->device & 63][multi ? 1 : 0];
    if (!configured) {
        SPRS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                            (int)smem));
        // Shared-memory carve-out = exactly what CTAS resident CTAs need, everything else stays
        // L1: every in-flight gather holds an L1 line, so the gather rate is bounded by
        // L1 lines / L2 latency (lab carve sweep: 0 % 303, 50 % 275, 100 % 106 Gnnz/s).
        int carve = (int)(((smem + 1024) * CTAS * 100 + 228 * 1024 - 1) / (228 * 1024));
        if (const char* e = getenv("SPRS_B200_SPMV_CARVEOUT")) carve = atoi(e);
        if (carve > 100) carve = 100;
        SPRS_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                            carve));
        configured = true;
    }
    uint64_t grid = (uint64_t)ctx->sm_count * CTAS;
    const uint64_t need = (t1 - t0 + SPMV_NWARPS - 1) / SPMV_NWARPS;
    if (grid > need) grid = need;
    kern<<<(unsigned)grid, SPMV_NWARPS * 32, smem, s>>>((const P*)m->d_indptr, m->d_indices,
                                                        m->d_data, m->d_tile_row, d_x, yt,
                                                        m->d_carry, m->nnz, (uint32_t)m->rows,
                                                        (uint32_t)t0, (uint32_t)t1, accumulate,
                                                        ctx->pol_evict_first, ctx->pol_evict_last);
    return SPRS_B200_OK;
}

template <typename P>
int launch_dispatch(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x,
                    const SpmvTargets& yt, int accumulate, uint64_t t0, uint64_t t1,
                    cudaStream_t s) {
    const SpmvVariant v = spmv_variant();
#define SPMV_CASE(E, CT)                                                          \
    if (v.epl == E && v.ctas_per_sm == CT)                                        \
        return launch_variant<P, E, CT>(ctx, m, d_x, yt, accumulate, t0, t1, s);
    SPMV_CASE(8, 2)
    SPMV_CASE(6, 3)
    SPMV_CASE(10, 2)
    SPMV_CASE(12, 2)
#undef SPMV_CASE
    SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "unknown SPRS_B200_SPMV_VARIANT (epl,ctas)");
}

int check_spmv_args(sprs_b200_ctx* ctx, const sprs_b200_csmat* m) {
    if (m->storage != SPRS_B200_CSR)
        SPRS_FAIL(ctx, SPRS_B200_ERR_STORAGE, "Storage mismatch: spmv needs a CSR mirror");
    if (!m->d_tile_row) SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "csmat has no SpMV partition");
    return SPRS_B200_OK;
}

}  // namespace

int spmv_tile_nnz() { return 32 * spmv_variant().epl; }

int spmv_prepare(sprs_b200_ctx* ctx, sprs_b200_csmat* m, cudaStream_t s) {
    if (m->storage != SPRS_B200_CSR) return SPRS_B200_OK;  // CSC mirrors are converted first
    const uint32_t wt = (uint32_t)spmv_tile_nnz();
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
    SPRS_TRY(check_spmv_args(ctx, m));
    if (m->rows == 0) return SPRS_B200_OK;
    if (m->indptr_bytes == 4)
        SPRS_TRY(launch_dispatch<uint32_t>(ctx, m, d_x, yt, accumulate, 0, m->n_tiles, s));
    else
        SPRS_TRY(launch_dispatch<uint64_t>(ctx, m, d_x, yt, accumulate, 0, m->n_tiles, s));
    ctx->launches += 1;
    if (m->n_tiles > 1) {
        const unsigned fgrid = (unsigned)((m->n_tiles - 1 + 255) / 256);
        spmv_fixup_kernel<<<fgrid, 256, 0, s>>>(m->d_tile_row, m->d_carry, yt, m->n_tiles);
        ctx->launches += 1;
    }
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

// SpMV over tiles [t0, t1) followed by the carries of the rows that END in those tiles: once
// this has run for every tile below t1 (chunks in increasing order on one stream), rows
// [0, tile_row[t1]) of y are final -- the same sums in the same order as the one-shot
// spmv_launch (apply_carries_ending_in adds a run's carries in tile order from its head, like
// spmv_fixup_kernel).  Used to pipeline something behind finished row ranges (the D2H copy of
// the host path, api.cu).
int spmv_launch_tile_range(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, const double* d_x,
                           double* d_y, int accumulate, uint64_t t0, uint64_t t1,
                           cudaStream_t s) {
    SPRS_TRY(check_spmv_args(ctx, m));
    if (t1 > m->n_tiles || t0 >= t1) SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "bad tile range");
    SpmvTargets yt;
    yt.n = 1;
    yt.p[0] = d_y;
    for (int q = 1; q < SPMV_MAX_TARGETS; ++q) yt.p[q] = nullptr;
    if (m->indptr_bytes == 4)
        SPRS_TRY(launch_dispatch<uint32_t>(ctx, m, d_x, yt, accumulate, t0, t1, s));
    else
        SPRS_TRY(launch_dispatch<uint64_t>(ctx, m, d_x, yt, accumulate, t0, t1, s));
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
