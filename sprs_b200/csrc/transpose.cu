// transpose.cu -- CSR <-> CSC conversion on the device.
//
// Replaces CsMatBase::to_other_storage / raw::convert_mat_storage
// (sprs/src/sparse/csmat.rs:1405-1426, 1782-1829): a counting sort of the non-zeros
// by inner index that keeps the outer order inside every bucket, so the result has
// ascending indices per outer dimension (the reference walks outer dims in order,
// csmat.rs:1814-1822).  It is what lets CSC operands use the CSR kernels:
// `csc_mulacc_*` and `mul_acc_mat_vec_csc` (prod.rs:74-99, 219-269) accumulate each
// output element in ascending column order, exactly the order the CSR kernels use on
// the converted matrix.
//
// Device algorithm: stable LSD radix sort of (inner index, source position) pairs,
// 8 bits per pass (ceil(log2(inner)/8) passes), then one gather pass writes
// out_indices[i] = outer(pos) (binary search in indptr) and out_data[i] = data[pos].
// A pass = per-block digit histograms -> device scan -> stable scatter (per-warp
// match_any ranking keeps equal keys in source order).  Deterministic, no atomics on
// the payload.  HBM-bound: 8 B read + 8 B write per non-zero per pass.

#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int RS_NT = 256;
constexpr int RS_IPT = 16;                 // items per thread
constexpr int RS_CHUNK = RS_NT * RS_IPT;   // 4096 items per block
constexpr int RS_WARPS = RS_NT / 32;
constexpr int RS_WCHUNK = RS_CHUNK / RS_WARPS;  // contiguous items per warp (512)

// pass 0 reads keys straight from `indices` and uses pos = position
__global__ void __launch_bounds__(RS_NT)
    rs_hist_kernel(const uint32_t* __restrict__ keys, uint64_t n, int shift,
                   uint32_t* __restrict__ hist /* [256][nblocks] */, uint32_t nblocks) {
    __shared__ uint32_t h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_CHUNK;
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const uint64_t j = base + threadIdx.x + (uint64_t)i * RS_NT;
        if (j < n) atomicAdd(&h[(keys[j] >> shift) & 255u], 1u);
    }
    __syncthreads();
    hist[(uint64_t)threadIdx.x * nblocks + blockIdx.x] = h[threadIdx.x];
}

__global__ void __launch_bounds__(RS_NT)
    rs_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ pos_in,
                      uint64_t n, int shift, const uint64_t* __restrict__ offsets /* scanned */,
                      uint32_t nblocks, uint32_t* __restrict__ keys_out,
                      uint32_t* __restrict__ pos_out) {
    __shared__ uint32_t wh[RS_WARPS][256];   // per-warp digit counters / running offsets
    __shared__ uint64_t gbase[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < RS_WARPS * 256; i += RS_NT) (&wh[0][0])[i] = 0;
    gbase[threadIdx.x] = offsets[(uint64_t)threadIdx.x * nblocks + blockIdx.x];
    __syncthreads();
    // each warp owns a CONTIGUOUS run of the block's items, visited row by row (32 at a time)
    const uint64_t wbase = (uint64_t)blockIdx.x * RS_CHUNK + (uint64_t)warp * RS_WCHUNK;
    uint32_t key[RS_IPT], pos[RS_IPT];
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const uint64_t j = wbase + (uint64_t)i * 32 + lane;
        const bool in = j < n;
        key[i] = in ? keys_in[j] : 0xffffffffu;
        pos[i] = in ? (pos_in ? pos_in[j] : (uint32_t)j) : 0u;
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, in ? d : 256u + lane);
        if (in && lane == __ffs(peers) - 1) wh[warp][d] += __popc(peers);
        __syncwarp();
    }
    __syncthreads();
    {   // exclusive scan over warps for digit = threadIdx.x, seeded with the global offset
        uint32_t run = 0;
#pragma unroll
        for (int w = 0; w < RS_WARPS; ++w) {
            const uint32_t c = wh[w][threadIdx.x];
            wh[w][threadIdx.x] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < RS_IPT; ++i) {
        const uint64_t j = wbase + (uint64_t)i * 32 + lane;
        const bool in = j < n;
        const uint32_t d = (key[i] >> shift) & 255u;
        const uint32_t peers = __match_any_sync(0xffffffffu, in ? d : 256u + lane);
        uint32_t off = 0;
        if (in) off = wh[warp][d] + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
        if (in && lane == __ffs(peers) - 1) wh[warp][d] += __popc(peers);
        __syncwarp();
        if (in) {
            const uint64_t dst = gbase[d] + off;
            keys_out[dst] = key[i];
            pos_out[dst] = pos[i];
        }
    }
}

__global__ void count_inner_kernel(const uint32_t* __restrict__ indices, uint64_t nnz,
                                   uint32_t* __restrict__ counts) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < nnz) atomicAdd(&counts[indices[i]], 1u);
}

template <typename P>
__global__ void gather_transposed_kernel(const uint32_t* __restrict__ pos, uint64_t nnz,
                                         const P* __restrict__ indptr, uint32_t outer,
                                         const double* __restrict__ data,
                                         uint32_t* __restrict__ out_indices,
                                         double* __restrict__ out_data) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= nnz) return;
    const uint32_t p = pos[i];
    uint32_t lo = 0, hi = outer;  // outer index o with indptr[o] <= p < indptr[o+1]
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if ((uint64_t)indptr[(size_t)mid + 1] > p)
            hi = mid;
        else
            lo = mid + 1;
    }
    out_indices[i] = lo;
    out_data[i] = data[p];
}

template <typename TIn, typename TOut>
__global__ void narrow_kernel(const TIn* __restrict__ in, TOut* __restrict__ out, uint64_t n) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (TOut)in[i];
}

inline unsigned grid_for(uint64_t n) { return (unsigned)((n + 255) / 256); }

// Stable LSD radix sort of n (key, pos) pairs on `key_bits` bits; pos_in == nullptr means
// pos = 0..n-1.  Uses the caller's ping-pong buffers; returns pointers to the sorted arrays.
struct SortBuffers {
    uint32_t *kbuf[2] = {nullptr, nullptr}, *pbuf[2] = {nullptr, nullptr}, *hist = nullptr;
    uint64_t* offs = nullptr;
    uint32_t nblocks = 0;
    int alloc(uint64_t n) {
        nblocks = (uint32_t)((n + RS_CHUNK - 1) / RS_CHUNK);
        if (cudaMalloc((void**)&kbuf[0], n * 4) != cudaSuccess ||
            cudaMalloc((void**)&kbuf[1], n * 4) != cudaSuccess ||
            cudaMalloc((void**)&pbuf[0], n * 4) != cudaSuccess ||
            cudaMalloc((void**)&pbuf[1], n * 4) != cudaSuccess ||
            cudaMalloc((void**)&hist, 256ull * nblocks * 4) != cudaSuccess ||
            cudaMalloc((void**)&offs, (256ull * nblocks + 1) * 8) != cudaSuccess)
            return SPRS_B200_ERR_CUDA;
        return SPRS_B200_OK;
    }
    void release() {
        for (int i = 0; i < 2; ++i) {
            if (kbuf[i]) cudaFree(kbuf[i]);
            if (pbuf[i]) cudaFree(pbuf[i]);
            kbuf[i] = pbuf[i] = nullptr;
        }
        if (hist) cudaFree(hist);
        if (offs) cudaFree(offs);
        hist = nullptr;
        offs = nullptr;
    }
};

int bits_for(uint64_t range) {
    int bits = 1;
    while (bits < 32 && (1ull << bits) < range) ++bits;
    return bits;
}

int stable_sort_pairs(sprs_b200_ctx* ctx, SortBuffers& b, const uint32_t* keys_in,
                      const uint32_t* pos_in, uint64_t n, int key_bits, cudaStream_t s,
                      const uint32_t** keys_out, const uint32_t** pos_out) {
    const int passes = (key_bits + 7) / 8;
    const uint32_t* kin = keys_in;
    const uint32_t* pin = pos_in;
    // never scatter into the buffer currently being read
    int w = (kin == b.kbuf[0] || pin == b.pbuf[0]) ? 1 : 0;
    for (int p = 0; p < passes; ++p) {
        rs_hist_kernel<<<b.nblocks, RS_NT, 0, s>>>(kin, n, 8 * p, b.hist, b.nblocks);
        ctx->launches += 1;
        SPRS_TRY((device_exclusive_scan<uint32_t, uint64_t>(ctx, b.hist, 256ull * b.nblocks,
                                                            b.offs, s)));
        rs_scatter_kernel<<<b.nblocks, RS_NT, 0, s>>>(kin, pin, n, 8 * p, b.offs, b.nblocks,
                                                      b.kbuf[w], b.pbuf[w]);
        ctx->launches += 1;
        kin = b.kbuf[w];
        pin = b.pbuf[w];
        w ^= 1;
    }
    *keys_out = kin;
    *pos_out = pin;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

// ---- COO -> CSR helpers (TriMatBase::to_csr, sprs/src/sparse/triplet_iter.rs:127-224)
__global__ void gather_u32_kernel(const uint32_t* __restrict__ src, const uint32_t* __restrict__ pos,
                                  uint64_t n, uint32_t* __restrict__ dst) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[pos[i]];
}
__global__ void head_flags_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
                                  uint64_t n, uint32_t* __restrict__ flag) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n) flag[i] = (i == 0 || row[i] != row[i - 1] || col[i] != col[i - 1]) ? 1u : 0u;
}
// one thread per run head: sums the duplicates in sorted (= insertion) order, like the
// reference's duplicate summation (triplet_iter.rs:143-176), writes the unique entry
__global__ void compress_runs_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
                                     const uint32_t* __restrict__ pos, const uint32_t* __restrict__ flag,
                                     const uint64_t* __restrict__ uidx, const double* __restrict__ vals,
                                     uint64_t n, uint32_t* __restrict__ out_idx,
                                     double* __restrict__ out_val, uint32_t* __restrict__ row_counts) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= n || !flag[i]) return;
    double sum = vals[pos[i]];
    for (uint64_t j = i + 1; j < n && !flag[j]; ++j) sum = __dadd_rn(sum, vals[pos[j]]);
    const uint64_t u = uidx[i];
    out_idx[u] = col[i];
    out_val[u] = sum;
    atomicAdd(&row_counts[row[i]], 1u);
}

}  // namespace

int transpose_launch(sprs_b200_ctx* ctx, const sprs_b200_csmat* m, sprs_b200_csmat* t,
                     cudaStream_t s) {
    if (m->nnz >= 0xffffffffull)
        SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "to_other_storage: nnz >= 2^32 not supported");
    // gh374 (sprs/tests/gh374.rs): the outer indices must fit the index type; device
    // mirrors use u32 indices, so only > 2^32-1 outer dims can fail.
    if (m->outer > 0xffffffffull)
        SPRS_FAIL(ctx, SPRS_B200_ERR_INDEX_RANGE,
                  "Index type is not large enough to hold the number of rows requested");
    t->ctx = ctx;
    t->storage = m->storage == SPRS_B200_CSR ? SPRS_B200_CSC : SPRS_B200_CSR;
    t->rows = m->rows;
    t->cols = m->cols;
    t->nnz = m->nnz;
    t->outer = m->inner;
    t->inner = m->outer;
    t->indptr_bytes = 4;
    t->owns = true;
    const uint64_t nnz = m->nnz, inner = m->inner;
    SPRS_CUDA(ctx, cudaMalloc(&t->d_indptr, (inner + 1) * sizeof(uint32_t) + 16));
    SPRS_CUDA(ctx, cudaMalloc((void**)&t->d_indices, nnz * sizeof(uint32_t) + 16));
    SPRS_CUDA(ctx, cudaMalloc((void**)&t->d_data, nnz * sizeof(double) + 16));

    // ---- new indptr: histogram of inner indices + exclusive scan
    uint32_t* counts = nullptr;
    uint64_t* ip64 = nullptr;
    SPRS_CUDA(ctx, cudaMalloc((void**)&counts, (inner + 1) * sizeof(uint32_t)));
    SPRS_CUDA(ctx, cudaMalloc((void**)&ip64, (inner + 1) * sizeof(uint64_t)));
    int st = SPRS_B200_OK;
    SortBuffers sb;
    do {
        cudaError_t e = cudaMemsetAsync(counts, 0, (inner + 1) * sizeof(uint32_t), s);
        if (e != cudaSuccess) { st = SPRS_B200_ERR_CUDA; break; }
        if (nnz) {
            count_inner_kernel<<<grid_for(nnz), 256, 0, s>>>(m->d_indices, nnz, counts);
            ctx->launches += 1;
        }
        if ((st = device_exclusive_scan<uint32_t, uint64_t>(ctx, counts, inner, ip64, s)) !=
            SPRS_B200_OK)
            break;
        narrow_kernel<uint64_t, uint32_t><<<grid_for(inner + 1), 256, 0, s>>>(
            ip64, (uint32_t*)t->d_indptr, inner + 1);
        ctx->launches += 1;
        if (nnz == 0) break;

        // ---- stable LSD radix sort of (inner index, position)
        if ((st = sb.alloc(nnz)) != SPRS_B200_OK) {
            sprs_b200_set_error(ctx, "to_other_storage: cudaMalloc failed");
            break;
        }
        const uint32_t *kin = nullptr, *pin = nullptr;
        if ((st = stable_sort_pairs(ctx, sb, m->d_indices, nullptr, nnz, bits_for(inner), s, &kin,
                                    &pin)) != SPRS_B200_OK)
            break;
        if (m->indptr_bytes == 4)
            gather_transposed_kernel<uint32_t><<<grid_for(nnz), 256, 0, s>>>(
                pin, nnz, (const uint32_t*)m->d_indptr, (uint32_t)m->outer, m->d_data,
                t->d_indices, t->d_data);
        else
            gather_transposed_kernel<uint64_t><<<grid_for(nnz), 256, 0, s>>>(
                pin, nnz, (const uint64_t*)m->d_indptr, (uint32_t)m->outer, m->d_data,
                t->d_indices, t->d_data);
        ctx->launches += 1;
    } while (0);
    cudaError_t e = cudaStreamSynchronize(s);
    if (st == SPRS_B200_OK && e == cudaSuccess) e = cudaGetLastError();
    if (st == SPRS_B200_OK && e != cudaSuccess) {
        sprs_b200_set_error(ctx, cudaGetErrorString(e));
        st = SPRS_B200_ERR_CUDA;
    }
    cudaFree(counts);
    cudaFree(ip64);
    sb.release();
    return st;
}

// COO (device arrays, any order, duplicates allowed) -> CSR mirror with ascending unique
// column indices per row and duplicates summed: TriMatBase::to_csr
// (sprs/src/sparse/triplet_iter.rs:127-224).  Two stable radix sorts (by column, then by
// row) give the (row, col) order; duplicates are summed in insertion order.
namespace {
__global__ void triplet_bounds_kernel(const uint32_t* __restrict__ row, const uint32_t* __restrict__ col,
                                      uint64_t n, uint32_t rows, uint32_t cols,
                                      unsigned long long* __restrict__ bad) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i < n && (row[i] >= rows || col[i] >= cols)) atomicAdd(bad, 1ull);
}
}  // namespace

int triplets_to_csr_launch(sprs_b200_ctx* ctx, uint64_t rows, uint64_t cols, uint64_t n,
                           const uint32_t* d_row, const uint32_t* d_col, const double* d_val,
                           sprs_b200_csmat* t, cudaStream_t s) {
    if (n >= 0xffffffffull || rows > 0xffffffffull || cols > 0xffffffffull)
        SPRS_FAIL(ctx, SPRS_B200_ERR_INDEX_RANGE, "from_triplets: needs nnz, rows, cols < 2^32");
    if (n) {  // the reference panics on an out-of-range triplet (TriMatBase::add_triplet asserts);
              // here it would be an out-of-bounds device write in the counting pass
        unsigned long long* d_bad = nullptr;
        unsigned long long h_bad = 0;
        SPRS_CUDA(ctx, cudaMalloc((void**)&d_bad, 8));
        cudaMemsetAsync(d_bad, 0, 8, s);
        triplet_bounds_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(d_row, d_col, n, (uint32_t)rows,
                                                                          (uint32_t)cols, d_bad);
        ctx->launches += 1;
        cudaMemcpyAsync(&h_bad, d_bad, 8, cudaMemcpyDeviceToHost, s);
        const cudaError_t e = cudaStreamSynchronize(s);
        cudaFree(d_bad);
        if (e != cudaSuccess) SPRS_FAIL(ctx, SPRS_B200_ERR_CUDA, "from_triplets: %s", cudaGetErrorString(e));
        if (h_bad)
            SPRS_FAIL(ctx, SPRS_B200_ERR_STRUCTURE, "from_triplets: %llu triplet(s) outside %llu x %llu",
                      h_bad, (unsigned long long)rows, (unsigned long long)cols);
    }
    t->ctx = ctx;
    t->storage = SPRS_B200_CSR;
    t->rows = rows;
    t->cols = cols;
    t->outer = rows;
    t->inner = cols;
    t->indptr_bytes = 4;
    t->owns = true;
    t->nnz = 0;
    SPRS_CUDA(ctx, cudaMalloc(&t->d_indptr, (rows + 1) * sizeof(uint32_t) + 16));
    uint32_t *counts = nullptr, *row_s = nullptr, *col_s = nullptr, *flag = nullptr;
    uint64_t *ip64 = nullptr, *uidx = nullptr;
    SortBuffers sb;
    int st = SPRS_B200_OK;
    do {
        if (cudaMalloc((void**)&counts, (rows + 1) * 4) != cudaSuccess ||
            cudaMalloc((void**)&ip64, (rows + 1) * 8) != cudaSuccess) {
            st = SPRS_B200_ERR_CUDA;
            break;
        }
        cudaMemsetAsync(counts, 0, (rows + 1) * 4, s);
        uint64_t n_unique = 0;
        if (n) {
            if (cudaMalloc((void**)&row_s, n * 4) != cudaSuccess ||
                cudaMalloc((void**)&col_s, n * 4) != cudaSuccess ||
                cudaMalloc((void**)&flag, n * 4) != cudaSuccess ||
                cudaMalloc((void**)&uidx, (n + 1) * 8) != cudaSuccess ||
                sb.alloc(n) != SPRS_B200_OK) {
                st = SPRS_B200_ERR_CUDA;
                break;
            }
            const uint32_t *k1 = nullptr, *p1 = nullptr, *k2 = nullptr, *p2 = nullptr;
            if ((st = stable_sort_pairs(ctx, sb, d_col, nullptr, n, bits_for(cols), s, &k1, &p1)) !=
                SPRS_B200_OK)
                break;
            gather_u32_kernel<<<grid_for(n), 256, 0, s>>>(d_row, p1, n, row_s);  // row in col order
            if ((st = stable_sort_pairs(ctx, sb, row_s, p1, n, bits_for(rows), s, &k2, &p2)) !=
                SPRS_B200_OK)
                break;
            gather_u32_kernel<<<grid_for(n), 256, 0, s>>>(d_col, p2, n, col_s);
            head_flags_kernel<<<grid_for(n), 256, 0, s>>>(k2, col_s, n, flag);
            ctx->launches += 3;
            if ((st = device_exclusive_scan<uint32_t, uint64_t>(ctx, flag, n, uidx, s)) != SPRS_B200_OK)
                break;
            if (cudaMemcpyAsync(&n_unique, uidx + n, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess ||
                cudaStreamSynchronize(s) != cudaSuccess) {
                st = SPRS_B200_ERR_CUDA;
                break;
            }
            if (cudaMalloc((void**)&t->d_indices, n_unique * 4 + 16) != cudaSuccess ||
                cudaMalloc((void**)&t->d_data, n_unique * 8 + 16) != cudaSuccess) {
                st = SPRS_B200_ERR_CUDA;
                break;
            }
            compress_runs_kernel<<<grid_for(n), 256, 0, s>>>(k2, col_s, p2, flag, uidx, d_val, n,
                                                            t->d_indices, t->d_data, counts);
            ctx->launches += 1;
        } else {
            if (cudaMalloc((void**)&t->d_indices, 16) != cudaSuccess ||
                cudaMalloc((void**)&t->d_data, 16) != cudaSuccess) {
                st = SPRS_B200_ERR_CUDA;
                break;
            }
        }
        t->nnz = n_unique;
        if ((st = device_exclusive_scan<uint32_t, uint64_t>(ctx, counts, rows, ip64, s)) != SPRS_B200_OK)
            break;
        narrow_kernel<uint64_t, uint32_t><<<grid_for(rows + 1), 256, 0, s>>>(
            ip64, (uint32_t*)t->d_indptr, rows + 1);
        ctx->launches += 1;
    } while (0);
    cudaError_t e = cudaStreamSynchronize(s);
    if (st == SPRS_B200_OK && e == cudaSuccess) e = cudaGetLastError();
    if (st == SPRS_B200_OK && e != cudaSuccess) st = SPRS_B200_ERR_CUDA;
    if (st == SPRS_B200_ERR_CUDA) sprs_b200_set_error(ctx, "from_triplets: CUDA allocation or kernel failed");
    if (counts) cudaFree(counts);
    if (ip64) cudaFree(ip64);
    if (row_s) cudaFree(row_s);
    if (col_s) cudaFree(col_s);
    if (flag) cudaFree(flag);
    if (uidx) cudaFree(uidx);
    sb.release();
    return st;
}
