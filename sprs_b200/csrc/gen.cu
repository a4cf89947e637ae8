// gen.cu -- synthetic sparse inputs generated directly in HBM.
//
// Not part of the product path: these kernels only manufacture benchmark inputs,
// because a 1e9-nnz matrix cannot practically be built on the host and shipped over
// PCIe every run.  Counter-based: candidate edge e of a
// given seed is a pure function of (seed, e), so any rank can regenerate any slice.
//
//  * uniform keys follow sprs-rand's distribution (sprs-rand/src/lib.rs:36-81):
//    a uniform random row per non-zero, uniform columns, duplicates removed later;
//  * R-MAT keys: the reference has no R-MAT generator; this is the
//    Graph500 recursive quadrant choice with probabilities (a, b, c, 1-a-b-c),
//    `scale` levels, candidates with an index >= n rejected.
// Keys are row<<32 | col; the caller sorts and removes duplicates.

#include "common.cuh"

namespace {

__global__ void rmat_keys_kernel(uint64_t seed, int scale, uint64_t n_rows, uint64_t n_cols,
                                 uint32_t ta, uint32_t tab, uint32_t tabc, uint64_t first,
                                 uint64_t count, uint64_t* __restrict__ keys) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t e = first + i;
    const uint64_t base = mix64(seed ^ mix64(e));
    uint64_t row = 0, col = 0, w = 0;
    for (int l = 0; l < scale; ++l) {
        if ((l & 3) == 0) w = mix64(base + (uint64_t)(l >> 2) * 0xD1B54A32D192ED03ull);
        const uint32_t u = (uint32_t)(w >> (16 * (l & 3))) & 0xffffu;
        const uint32_t rbit = u >= tab;                               // quadrants c, d
        const uint32_t cbit = (u >= ta && u < tab) || (u >= tabc);    // quadrants b, d
        row = (row << 1) | rbit;
        col = (col << 1) | cbit;
    }
    keys[i] = (row < n_rows && col < n_cols) ? ((row << 32) | col) : ~0ull;
}

__global__ void uniform_keys_kernel(uint64_t seed, uint64_t n_rows, uint64_t n_cols,
                                    uint64_t first, uint64_t count,
                                    uint64_t* __restrict__ keys) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t e = first + i;
    const uint64_t h1 = mix64(seed ^ mix64(e));
    const uint64_t h2 = mix64(h1 + 0xD1B54A32D192ED03ull);
    const uint64_t row = __umul64hi(h1, n_rows);  // uniform in [0, n_rows)
    const uint64_t col = __umul64hi(h2, n_cols);
    keys[i] = (row << 32) | col;
}

__global__ void normal_from_keys_kernel(uint64_t seed, const uint64_t* __restrict__ keys,
                                        uint64_t count, double* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t key = keys ? keys[i] : i;
    const uint64_t h1 = mix64(seed ^ mix64(key));
    const uint64_t h2 = mix64(h1 + 0xD1B54A32D192ED03ull);
    const double u1 = ((double)(h1 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    const double u2 = ((double)(h2 >> 11) + 0.5) * (1.0 / 9007199254740992.0);
    out[i] = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);  // Box-Muller, N(0,1)
}

__global__ void split_keys_kernel(const uint64_t* __restrict__ keys, uint64_t count,
                                  uint32_t* __restrict__ rows, uint32_t* __restrict__ cols) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t k = keys[i];
    if (rows) rows[i] = (uint32_t)(k >> 32);
    cols[i] = (uint32_t)(k & 0xffffffffull);
}

__global__ void hash_keys_kernel(uint64_t seed, const uint64_t* __restrict__ keys,
                                 uint64_t count, uint64_t* __restrict__ out) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = mix64(seed ^ mix64(keys[i])) >> 1;  // 63 bits: safe as a signed torch int64
}

inline unsigned grid_for(uint64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

extern "C" int sprs_b200_gen_rmat_keys(sprs_b200_ctx* ctx, uint64_t seed, int scale,
                                       uint64_t n_rows, uint64_t n_cols, double a, double b,
                                       double c, uint64_t first, uint64_t count,
                                       uint64_t* d_keys, void* stream) {
    if (!ctx || !d_keys) return SPRS_B200_ERR_ARGUMENT;
    if (scale < 1 || scale > 31 || a <= 0 || b < 0 || c < 0 || a + b + c >= 1.0)
        SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "bad R-MAT parameters");
    if (count == 0) return SPRS_B200_OK;
    if (count > 0x7fffffffull * 256)
        SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "generate in chunks of < 2^39 candidates");
    const uint32_t ta = (uint32_t)(a * 65536.0 + 0.5);
    const uint32_t tab = (uint32_t)((a + b) * 65536.0 + 0.5);
    const uint32_t tabc = (uint32_t)((a + b + c) * 65536.0 + 0.5);
    rmat_keys_kernel<<<grid_for(count), 256, 0, pick_stream(ctx, stream)>>>(
        seed, scale, n_rows, n_cols, ta, tab, tabc, first, count, d_keys);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

extern "C" int sprs_b200_gen_uniform_keys(sprs_b200_ctx* ctx, uint64_t seed, uint64_t n_rows,
                                          uint64_t n_cols, uint64_t first, uint64_t count,
                                          uint64_t* d_keys, void* stream) {
    if (!ctx || !d_keys) return SPRS_B200_ERR_ARGUMENT;
    if (n_rows == 0 || n_cols == 0 || n_rows > 0xffffffffull || n_cols > 0xffffffffull)
        SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "bad shape");
    if (count == 0) return SPRS_B200_OK;
    uniform_keys_kernel<<<grid_for(count), 256, 0, pick_stream(ctx, stream)>>>(
        seed, n_rows, n_cols, first, count, d_keys);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

extern "C" int sprs_b200_gen_normal_from_keys(sprs_b200_ctx* ctx, uint64_t seed,
                                              const uint64_t* d_keys, uint64_t count,
                                              double* d_out, void* stream) {
    if (!ctx || !d_out) return SPRS_B200_ERR_ARGUMENT;
    if (count == 0) return SPRS_B200_OK;
    normal_from_keys_kernel<<<grid_for(count), 256, 0, pick_stream(ctx, stream)>>>(
        seed, d_keys, count, d_out);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

extern "C" int sprs_b200_gen_split_keys(sprs_b200_ctx* ctx, const uint64_t* d_keys,
                                        uint64_t count, uint32_t* d_rows, uint32_t* d_cols,
                                        void* stream) {
    if (!ctx || !d_keys || !d_cols) return SPRS_B200_ERR_ARGUMENT;
    if (count == 0) return SPRS_B200_OK;
    split_keys_kernel<<<grid_for(count), 256, 0, pick_stream(ctx, stream)>>>(d_keys, count,
                                                                             d_rows, d_cols);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

extern "C" int sprs_b200_gen_hash_keys(sprs_b200_ctx* ctx, uint64_t seed,
                                       const uint64_t* d_keys, uint64_t count, uint64_t* d_out,
                                       void* stream) {
    if (!ctx || !d_keys || !d_out) return SPRS_B200_ERR_ARGUMENT;
    if (count == 0) return SPRS_B200_OK;
    hash_keys_kernel<<<grid_for(count), 256, 0, pick_stream(ctx, stream)>>>(seed, d_keys, count,
                                                                            d_out);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}
