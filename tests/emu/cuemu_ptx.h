// cuemu_ptx.h -- TEST INFRASTRUCTURE ONLY: stand-ins for the inline-PTX wrappers at the top
// of sprs_b200/csrc/spmv.cu (mbarrier, 1-D TMA bulk copy, L2 policies, hinted loads).
// transform.py swaps that block for this include.  Semantics kept: an mbarrier phase
// completes when its arrival count reaches zero AND all expected bytes have landed; a bulk
// copy lands immediately (there is no asynchrony to emulate, only the bookkeeping).
#pragma once
#include <unordered_map>

namespace cuemu_ptx {
struct Mbar {
    uint32_t phase = 0;
    int32_t count = 0, pending = 0;
    int64_t tx = 0;
};
static inline std::unordered_map<const void*, Mbar>& table() {
    static std::unordered_map<const void*, Mbar> t;
    return t;
}
static inline void settle(Mbar& b) {
    if (b.pending == 0 && b.tx == 0) {
        b.phase ^= 1u;
        b.pending = b.count;
        cuemu::note_progress();
    }
}
}  // namespace cuemu_ptx

static inline void mbar_init(uint64_t* bar, uint32_t count) {
    cuemu_ptx::Mbar b;
    b.count = b.pending = (int32_t)count;
    cuemu_ptx::table()[bar] = b;
}
static inline void fence_mbar_init() {}
static inline void fence_proxy_async() {}
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    cuemu_ptx::Mbar& b = cuemu_ptx::table().at(bar);
    b.tx += bytes;
    b.pending -= 1;
    cuemu_ptx::settle(b);
}
static inline void mbar_wait(uint64_t* bar, uint32_t parity) {
    // try_wait.parity(P) succeeds once the phase of parity P has completed
    while (cuemu_ptx::table().at(bar).phase == parity) cuemu::yield();
}
static inline void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t) {
    if (bytes % 16 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) {
        fprintf(stderr, "cuemu: cp.async.bulk needs 16-byte aligned addresses and size "
                        "(dst %p src %p bytes %u)\n", dst, src, bytes);
        abort();
    }
    std::memcpy(dst, src, bytes);
    cuemu_ptx::Mbar& b = cuemu_ptx::table().at(bar);
    b.tx -= bytes;
    cuemu_ptx::settle(b);
}
static inline uint64_t policy_evict_first() { return 0; }
static inline uint64_t policy_evict_last() { return 0; }
static inline double ldg_f64_hint(const double* p, uint64_t) { return *p; }
static inline uint32_t ldg_stream_u32(const uint32_t* p, uint64_t) { return *p; }
static inline double ldg_stream_f64(const double* p, uint64_t) { return *p; }
static inline void bulk_s2g(void* dst, const void* src, uint32_t bytes) {
    if (bytes % 16 || ((uintptr_t)dst & 15) || ((uintptr_t)src & 15)) {
        fprintf(stderr, "cuemu: cp.async.bulk (s2g) needs 16-byte aligned addresses and size "
                        "(dst %p src %p bytes %u)\n", dst, src, bytes);
        abort();
    }
    std::memcpy(dst, src, bytes);
}
static inline void bulk_commit_group() {}
static inline void bulk_wait_group_read0() {}
static inline void bulk_wait_group0() {}
