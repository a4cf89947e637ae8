// cuemu_ptx.h -- TEST INFRASTRUCTURE ONLY: stand-ins for the inline-PTX wrappers of
// sprs_b200/csrc/ptx.cuh (L2 policies, hinted loads, TMA bulk store).  transform.py swaps that
// header for this one.  A bulk store lands immediately (there is no asynchrony to emulate); its
// alignment rules are checked.
#pragma once

static inline void fence_proxy_async() {}
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
