// scan.cuh -- device-wide exclusive prefix sum (three small kernels, deterministic).
// out[i] = sum_{j<i} in[j] for i in [0, n]; out has n+1 entries (out[n] = total).
// Used for indptr construction in the transpose and in SpGEMM (the reference's serial
// prefix sums: csmat.rs:1805-1810, smmp.rs:324-331).
#pragma once
#include "common.cuh"

namespace scan_detail {

constexpr int SCAN_NT = 256;
constexpr int SCAN_IPT = 8;
constexpr int SCAN_CHUNK = SCAN_NT * SCAN_IPT;

template <typename T>
__device__ __forceinline__ T block_exclusive_scan(T v, T* total, T* warp_sums /* 32 */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T u = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += u;
    }
    if (lane == 31) warp_sums[warp] = inc;
    __syncthreads();
    if (warp == 0) {
        T w = lane < (blockDim.x >> 5) ? warp_sums[lane] : T(0);
        T winc = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const T u = __shfl_up_sync(0xffffffffu, winc, o);
            if (lane >= o) winc += u;
        }
        warp_sums[lane] = winc - w;  // exclusive prefix of the warp sums
        if (lane == 31) *total = winc;
    }
    __syncthreads();
    const T res = warp_sums[warp] + inc - v;
    __syncthreads();
    return res;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_NT)
    scan_reduce_kernel(const TIn* __restrict__ in, uint64_t n, TOut* __restrict__ block_sums) {
    __shared__ TOut warp_sums[32];
    __shared__ TOut total;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK;
    TOut s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        const uint64_t j = base + threadIdx.x + (uint64_t)i * SCAN_NT;
        if (j < n) s += (TOut)in[j];
    }
    block_exclusive_scan<TOut>(s, &total, warp_sums);
    if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

template <typename TOut>
__global__ void __launch_bounds__(1024)
    scan_block_sums_kernel(TOut* __restrict__ block_sums, uint64_t nblocks,
                           TOut* __restrict__ grand_total) {
    __shared__ TOut warp_sums[32];
    __shared__ TOut total;
    TOut carry = 0;
    for (uint64_t base = 0; base < nblocks; base += 1024) {
        const uint64_t j = base + threadIdx.x;
        const TOut v = j < nblocks ? block_sums[j] : TOut(0);
        const TOut ex = block_exclusive_scan<TOut>(v, &total, warp_sums);
        if (j < nblocks) block_sums[j] = carry + ex;
        carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0) *grand_total = carry;
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(SCAN_NT)
    scan_apply_kernel(const TIn* __restrict__ in, uint64_t n,
                      const TOut* __restrict__ block_sums, TOut* __restrict__ out) {
    __shared__ TOut warp_sums[32];
    __shared__ TOut total;
    const uint64_t base = (uint64_t)blockIdx.x * SCAN_CHUNK + (uint64_t)threadIdx.x * SCAN_IPT;
    TOut v[SCAN_IPT];
    TOut s = 0;
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        v[i] = (base + i < n) ? (TOut)in[base + i] : TOut(0);
        s += v[i];
    }
    TOut run = block_exclusive_scan<TOut>(s, &total, warp_sums) + block_sums[blockIdx.x];
#pragma unroll
    for (int i = 0; i < SCAN_IPT; ++i) {
        if (base + i < n) out[base + i] = run;
        run += v[i];
    }
}

}  // namespace scan_detail

// d_out must hold n+1 entries; may not alias d_in.  Uses ctx scratch slot 3.
template <typename TIn, typename TOut>
int device_exclusive_scan(sprs_b200_ctx* ctx, const TIn* d_in, uint64_t n, TOut* d_out,
                          cudaStream_t s) {
    using namespace scan_detail;
    if (n == 0) {
        SPRS_CUDA(ctx, cudaMemsetAsync(d_out, 0, sizeof(TOut), s));
        return SPRS_B200_OK;
    }
    const uint64_t nblocks = (n + SCAN_CHUNK - 1) / SCAN_CHUNK;
    void* scratch = nullptr;
    SPRS_TRY(ctx_scratch(ctx, 3, nblocks * sizeof(TOut), &scratch));
    TOut* block_sums = (TOut*)scratch;
    scan_reduce_kernel<TIn, TOut><<<(unsigned)nblocks, SCAN_NT, 0, s>>>(d_in, n, block_sums);
    scan_block_sums_kernel<TOut><<<1, 1024, 0, s>>>(block_sums, nblocks, d_out + n);
    scan_apply_kernel<TIn, TOut><<<(unsigned)nblocks, SCAN_NT, 0, s>>>(d_in, n, block_sums, d_out);
    ctx->launches += 3;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}
