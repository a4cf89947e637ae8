// peer.cu -- multi-GPU plumbing of the row-partitioned SpMV (one process per GPU).
//
// The reference has no multi-device code; the natural sharding primitive is
// slice_outer (slicing.rs:65-89).  Each rank owns a contiguous row block of A, x is
// replicated, and the single exchange step is an all-gather of y.  Here the all-gather is
// FUSED into the SpMV kernel: every rank maps the other ranks' y buffers through CUDA IPC
// (NVLink/NVSwitch peer access) and the kernel stores each finished row into all of them,
// so the transfer overlaps the rest of the compute; only a stream-ordered barrier remains.
#include "common.cuh"

namespace {
// Push `count` doubles starting at row_off from this GPU's y into the same position of every
// peer buffer (own all-gather "put"): coalesced 8-byte loads, n_peers coalesced stores each.
__global__ void __launch_bounds__(256)
    peer_push_kernel(const double* __restrict__ src, SpmvTargets dst, uint64_t count) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < count; i += 4 * stride) {
        double v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = (i + u * stride < count) ? src[i + u * stride] : 0.0;
#pragma unroll
        for (int q = 0; q < SPMV_MAX_TARGETS; ++q)
            if (q < dst.n) {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (i + u * stride < count) dst.p[q][i + u * stride] = v[u];
            }
    }
}
}  // namespace

extern "C" {

int sprs_b200_peer_push_dev(sprs_b200_ctx* ctx, const double* d_y_own, uint64_t row_offset,
                            uint64_t rows, int n_peers, double* const* d_y_peers, void* stream) {
    if (!ctx || !d_y_own || (n_peers && !d_y_peers)) return SPRS_B200_ERR_ARGUMENT;
    if (n_peers < 0 || n_peers > SPMV_MAX_TARGETS)
        SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "n_peers must be 0..%d", SPMV_MAX_TARGETS);
    if (rows == 0 || n_peers == 0) return SPRS_B200_OK;
    SpmvTargets dst;
    dst.n = n_peers;
    for (int q = 0; q < SPMV_MAX_TARGETS; ++q)
        dst.p[q] = q < n_peers ? d_y_peers[q] + row_offset : nullptr;
    uint64_t blocks = (rows + 1023) / 1024;
    const uint64_t cap = (uint64_t)ctx->sm_count * 2;
    if (blocks > cap) blocks = cap;
    peer_push_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(d_y_own + row_offset, dst,
                                                                        rows);
    ctx->launches += 1;
    SPRS_CUDA(ctx, cudaGetLastError());
    return SPRS_B200_OK;
}

int sprs_b200_peer_alloc(sprs_b200_ctx* ctx, uint64_t bytes, void** d_ptr,
                         unsigned char ipc_handle[64]) {
    if (!ctx || !d_ptr || !ipc_handle) return SPRS_B200_ERR_ARGUMENT;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    SPRS_CUDA(ctx, cudaSetDevice(ctx->device));
    void* p = nullptr;
    SPRS_CUDA(ctx, cudaMalloc(&p, bytes ? bytes : 16));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
        cudaFree(p);
        SPRS_FAIL(ctx, SPRS_B200_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
    }
    memcpy(ipc_handle, &h, 64);
    *d_ptr = p;
    return SPRS_B200_OK;
}

int sprs_b200_peer_open(sprs_b200_ctx* ctx, const unsigned char ipc_handle[64], void** d_ptr) {
    if (!ctx || !d_ptr || !ipc_handle) return SPRS_B200_ERR_ARGUMENT;
    SPRS_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    SPRS_CUDA(ctx, cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return SPRS_B200_OK;
}

int sprs_b200_peer_close(sprs_b200_ctx* ctx, void* d_ptr) {
    if (!ctx) return SPRS_B200_ERR_ARGUMENT;
    if (d_ptr) SPRS_CUDA(ctx, cudaIpcCloseMemHandle(d_ptr));
    return SPRS_B200_OK;
}

int sprs_b200_peer_free(sprs_b200_ctx* ctx, void* d_ptr) {
    if (!ctx) return SPRS_B200_ERR_ARGUMENT;
    if (d_ptr) SPRS_CUDA(ctx, cudaFree(d_ptr));
    return SPRS_B200_OK;
}

int sprs_b200_copy_dev(sprs_b200_ctx* ctx, void* dst, const void* src, uint64_t bytes,
                       void* stream) {
    if (!ctx || (bytes && (!dst || !src))) return SPRS_B200_ERR_ARGUMENT;
    SPRS_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice,
                                   (cudaStream_t)stream));
    return SPRS_B200_OK;
}

int sprs_b200_spmv_allgather_dev(sprs_b200_ctx* ctx, const sprs_b200_csmat* mat,
                                 const double* d_x, uint64_t row_offset, int n_targets,
                                 double* const* d_y_bufs, int accumulate, void* stream) {
    if (!ctx || !mat || !d_y_bufs) return SPRS_B200_ERR_ARGUMENT;
    if (n_targets < 1 || n_targets > SPMV_MAX_TARGETS)
        SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "n_targets must be 1..%d", SPMV_MAX_TARGETS);
    SpmvTargets yt;
    yt.n = n_targets;
    for (int q = 0; q < SPMV_MAX_TARGETS; ++q)
        yt.p[q] = q < n_targets ? d_y_bufs[q] + row_offset : nullptr;
    return spmv_launch_targets(ctx, mat, d_x, yt, accumulate, (cudaStream_t)stream);
}

int sprs_b200_spmv_stream_push_dev(sprs_b200_ctx* ctx, sprs_b200_csmat* mat, const double* d_x,
                                   uint64_t row_offset, int n_targets, double* const* d_y_bufs,
                                   int accumulate, int put_ctas, void* stream) {
    if (!ctx || !mat || !d_y_bufs) return SPRS_B200_ERR_ARGUMENT;
    if (n_targets < 1 || n_targets > SPMV_MAX_TARGETS)
        SPRS_FAIL(ctx, SPRS_B200_ERR_ARGUMENT, "n_targets must be 1..%d", SPMV_MAX_TARGETS);
    SpmvTargets yt;
    yt.n = n_targets;
    for (int q = 0; q < SPMV_MAX_TARGETS; ++q)
        yt.p[q] = q < n_targets ? d_y_bufs[q] + row_offset : nullptr;
    return spmv_launch_stream_push(ctx, mat, d_x, yt, accumulate, put_ctas,
                                   (cudaStream_t)stream);
}

}  // extern "C"
