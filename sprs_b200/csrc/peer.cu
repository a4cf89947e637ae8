// peer.cu -- multi-GPU plumbing of the row-partitioned SpMV (one process per GPU).
//
// The reference has no multi-device code; the natural sharding primitive is
// slice_outer (slicing.rs:65-89).  Each rank owns a contiguous row block of A, x is
// replicated, and the single exchange step is an all-gather of y.  Here the all-gather is
// FUSED into the SpMV kernel: every rank maps the other ranks' y buffers through CUDA IPC
// (NVLink/NVSwitch peer access) and the kernel stores each finished row into all of them,
// so the transfer overlaps the rest of the compute; only a stream-ordered barrier remains.
#include "common.cuh"

extern "C" {

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

}  // extern "C"
