// spgemm.cu -- placeholder until the two-phase hash SpGEMM lands.
#include "common.cuh"
struct sprs_b200_spgemm { int dummy; };
extern "C" {
int sprs_b200_spgemm_symbolic(sprs_b200_ctx* ctx, const sprs_b200_csmat*, const sprs_b200_csmat*,
                              sprs_b200_spgemm**, uint64_t*) {
    if (!ctx) return SPRS_B200_ERR_ARGUMENT;
    SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "spgemm: not built yet");
}
int sprs_b200_spgemm_numeric(sprs_b200_ctx* ctx, sprs_b200_spgemm*, void*, int, void*, int, double*) {
    if (!ctx) return SPRS_B200_ERR_ARGUMENT;
    SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "spgemm: not built yet");
}
int sprs_b200_spgemm_numeric_dev(sprs_b200_ctx* ctx, sprs_b200_spgemm*, sprs_b200_csmat**) {
    if (!ctx) return SPRS_B200_ERR_ARGUMENT;
    SPRS_FAIL(ctx, SPRS_B200_ERR_UNSUPPORTED, "spgemm: not built yet");
}
uint64_t sprs_b200_spgemm_nprod(const sprs_b200_spgemm*) { return 0; }
int sprs_b200_spgemm_free(sprs_b200_spgemm* p) { delete p; return SPRS_B200_OK; }
}
