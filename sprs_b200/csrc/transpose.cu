// transpose.cu -- placeholder until the device counting-sort lands (next commit).
#include "common.cuh"
int transpose_launch(sprs_b200_ctx* ctx, const sprs_b200_csmat*, sprs_b200_csmat*, cudaStream_t) {
    SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "to_other_storage: not built yet");
}
