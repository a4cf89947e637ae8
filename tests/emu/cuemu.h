// cuemu.h -- TEST INFRASTRUCTURE ONLY: a CPU emulation of the small CUDA subset the kernels
// in sprs_b200/csrc use, so that kernel LOGIC (indexing, barriers, warp collectives, the
// host-side launch sequences) can be exercised where no GPU is attached.
//
// It is never part of the product: libsprs_b200.so is built by nvcc from the unmodified
// sources; tests/emu/Makefile builds a separate libsprs_b200_emu.so from a mechanically
// transformed copy (tests/emu/transform.py) that only tests/test_emu_*.py load.  Nothing
// measured, benchmarked or shipped goes through it, and it says nothing about performance
// or about hardware behaviour (memory model, TMA, caches).
//
// Model: one fiber per CUDA thread, blocks run one after the other on the calling OS thread.
// A fiber runs until it reaches a block barrier, a warp collective or an mbarrier wait and
// then yields; collectives rendezvous over the lanes named in the mask that are still alive.
// A round in which no fiber makes progress is reported as a deadlock (abort).
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define CUEMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) alignas(n)
#define __grid_constant__

struct uint3 {
    unsigned x, y, z;
};
struct dim3 {
    unsigned x = 1, y = 1, z = 1;
    dim3() = default;
    dim3(unsigned long long x_, unsigned y_ = 1, unsigned z_ = 1) : x((unsigned)x_), y(y_), z(z_) {}
};
struct double2 {
    double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint2 {
    unsigned x, y;
};
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
struct uint4 {
    unsigned x, y, z, w;
};
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) {
    return uint4{x, y, z, w};
}

// ---- runtime API subset -------------------------------------------------------------------
typedef int cudaError_t;
enum { cudaSuccess = 0, cudaErrorMemoryAllocation = 2, cudaErrorNotSupported = 801 };
typedef struct cuemu_stream* cudaStream_t;
typedef struct cuemu_event* cudaEvent_t;
enum { cudaEventDisableTiming = 2 };
enum cudaMemcpyKind {
    cudaMemcpyHostToHost = 0,
    cudaMemcpyHostToDevice = 1,
    cudaMemcpyDeviceToHost = 2,
    cudaMemcpyDeviceToDevice = 3
};
enum { cudaStreamNonBlocking = 1 };
enum cudaFuncAttribute {
    cudaFuncAttributeMaxDynamicSharedMemorySize = 8,
    cudaFuncAttributePreferredSharedMemoryCarveout = 9
};
struct cudaDeviceProp {
    char name[256];
    int multiProcessorCount;
    int l2CacheSize;
    size_t totalGlobalMem;
    int major, minor;
};
enum cudaMemoryType { cudaMemoryTypeUnregistered = 0, cudaMemoryTypeHost = 1, cudaMemoryTypeDevice = 2 };
struct cudaPointerAttributes {
    cudaMemoryType type;
    int device;
    void* devicePointer;
    void* hostPointer;
};
struct cudaIpcMemHandle_t {
    char reserved[64];
};
enum { cudaIpcMemLazyEnablePeerAccess = 1 };

const char* cudaGetErrorString(cudaError_t e);
cudaError_t cudaGetLastError();
cudaError_t cudaMalloc(void** p, size_t bytes);
template <class T>
static inline cudaError_t cudaMalloc(T** p, size_t bytes) {
    return cudaMalloc((void**)p, bytes);
}
cudaError_t cudaFree(void* p);
// stream-ordered allocator: plain allocations here (launches are synchronous)
static inline cudaError_t cudaMallocAsync(void** p, size_t bytes, cudaStream_t) { return cudaMalloc(p, bytes); }
template <class T>
static inline cudaError_t cudaMallocAsync(T** p, size_t bytes, cudaStream_t s) {
    return cudaMallocAsync((void**)p, bytes, s);
}
static inline cudaError_t cudaFreeAsync(void* p, cudaStream_t) { return cudaFree(p); }
typedef void* cudaMemPool_t;
enum cudaMemPoolAttr { cudaMemPoolAttrReleaseThreshold = 4 };
static inline cudaError_t cudaDeviceGetDefaultMemPool(cudaMemPool_t* p, int) { *p = nullptr; return cudaSuccess; }
static inline cudaError_t cudaMemPoolSetAttribute(cudaMemPool_t, cudaMemPoolAttr, void*) { return cudaSuccess; }
static inline cudaError_t cudaMemPoolTrimTo(cudaMemPool_t, size_t) { return cudaSuccess; }
cudaError_t cudaMallocHost(void** p, size_t bytes);
template <class T>
static inline cudaError_t cudaMallocHost(T** p, size_t bytes) {
    return cudaMallocHost((void**)p, bytes);
}
cudaError_t cudaFreeHost(void* p);
cudaError_t cudaMemcpy(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind);
cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind,
                            cudaStream_t s = nullptr);
cudaError_t cudaMemsetAsync(void* dst, int value, size_t bytes, cudaStream_t s = nullptr);
cudaError_t cudaMemset(void* dst, int value, size_t bytes);
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned flags);
cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, unsigned flags, int priority);
cudaError_t cudaDeviceGetStreamPriorityRange(int* least, int* greatest);
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned flags);
cudaError_t cudaEventDestroy(cudaEvent_t e);
cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t s);
cudaError_t cudaStreamWaitEvent(cudaStream_t s, cudaEvent_t e, unsigned flags);
cudaError_t cudaStreamDestroy(cudaStream_t s);
cudaError_t cudaStreamSynchronize(cudaStream_t s);
static inline cudaError_t cudaDeviceSynchronize() { return cudaSuccess; }  // launches are synchronous here
cudaError_t cudaSetDevice(int d);
cudaError_t cudaGetDeviceCount(int* n);
cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int d);
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p);
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t* h, void* p);
cudaError_t cudaIpcOpenMemHandle(void** p, cudaIpcMemHandle_t h, unsigned flags);
cudaError_t cudaIpcCloseMemHandle(void* p);
template <class F>
static inline cudaError_t cudaFuncSetAttribute(F, cudaFuncAttribute, int) {
    return cudaSuccess;
}

// ---- execution model ----------------------------------------------------------------------
namespace cuemu {

struct Fiber;
extern Fiber* g_cur;          // the running CUDA thread
extern dim3 g_block_dim, g_grid_dim;
extern uint3 g_block_idx;

struct Fiber {
    void* sp = nullptr;       // saved stack pointer
    char* stack = nullptr;
    bool done = true;
    uint3 tid{0, 0, 0};
    unsigned linear = 0, lane = 0, warp = 0;
};

struct Cfg {
    dim3 grid, block;
    size_t smem;
};
static inline Cfg cfg(dim3 grid, dim3 block, size_t smem = 0, cudaStream_t = nullptr) {
    return Cfg{grid, block, smem};
}
void launch(const Cfg& c, const std::function<void()>& thread_body);
void yield();                 // let the other threads of the block run
void* dyn_smem();             // dynamic shared memory of the running block (128-byte aligned)
void block_barrier();
// warp collectives: exchange one 64-bit payload per lane
uint64_t warp_exchange(unsigned mask, uint64_t mine, int src_lane);
unsigned warp_ballot(unsigned mask, bool pred);
unsigned warp_match_any(unsigned mask, uint64_t value);
void warp_barrier(unsigned mask);
void note_progress();

}  // namespace cuemu

#define threadIdx (cuemu::g_cur->tid)
#define blockIdx (cuemu::g_block_idx)
#define blockDim (cuemu::g_block_dim)
#define gridDim (cuemu::g_grid_dim)
#define warpSize 32

static inline void __syncthreads() { cuemu::block_barrier(); }
static inline void __syncwarp(unsigned mask = 0xffffffffu) { cuemu::warp_barrier(mask); }
static inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)p; }
static inline void __threadfence() {}
static inline void __nanosleep(unsigned) { cuemu::yield(); }
static inline void __trap() {
    fprintf(stderr, "cuemu: __trap()\n");
    abort();
}
long long clock64();
static inline void __threadfence_block() {}

namespace cuemu {
template <class T>
static inline uint64_t to_bits(T v) {
    static_assert(sizeof(T) <= 8, "warp payloads are at most 64 bits");
    uint64_t b = 0;
    std::memcpy(&b, &v, sizeof(T));
    return b;
}
template <class T>
static inline T from_bits(uint64_t b) {
    T v;
    std::memcpy(&v, &b, sizeof(T));
    return v;
}
static inline int lane_id() { return (int)g_cur->lane; }
}  // namespace cuemu

template <class T>
static inline T __shfl_sync(unsigned mask, T v, int src, int width = 32) {
    const int lane = cuemu::lane_id();
    const int s = (lane / width) * width + (((src % width) + width) % width);
    return cuemu::from_bits<T>(cuemu::warp_exchange(mask, cuemu::to_bits(v), s));
}
template <class T>
static inline T __shfl_xor_sync(unsigned mask, T v, int lane_mask, int width = 32) {
    const int lane = cuemu::lane_id();
    int s = lane ^ lane_mask;
    if (s / width != lane / width) s = lane;  // outside the segment: own value
    return cuemu::from_bits<T>(cuemu::warp_exchange(mask, cuemu::to_bits(v), s));
}
template <class T>
static inline T __shfl_up_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = cuemu::lane_id();
    int s = lane - (int)delta;
    if (s < (lane / width) * width) s = lane;
    return cuemu::from_bits<T>(cuemu::warp_exchange(mask, cuemu::to_bits(v), s));
}
template <class T>
static inline T __shfl_down_sync(unsigned mask, T v, unsigned delta, int width = 32) {
    const int lane = cuemu::lane_id();
    int s = lane + (int)delta;
    if (s >= (lane / width + 1) * width) s = lane;
    return cuemu::from_bits<T>(cuemu::warp_exchange(mask, cuemu::to_bits(v), s));
}
static inline unsigned __ballot_sync(unsigned mask, int pred) {
    return cuemu::warp_ballot(mask, pred != 0);
}
static inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0; }
static inline unsigned __reduce_or_sync(unsigned mask, unsigned v) {  // full-warp masks only
    for (int o = 16; o > 0; o >>= 1)
        v |= cuemu::from_bits<unsigned>(cuemu::warp_exchange(mask, cuemu::to_bits(v), cuemu::lane_id() ^ o));
    return v;
}
static inline int __all_sync(unsigned mask, int pred) {
    return cuemu::warp_ballot(mask, pred == 0) == 0;
}
template <class T>
static inline unsigned __match_any_sync(unsigned mask, T v) {
    return cuemu::warp_match_any(mask, cuemu::to_bits(v));
}

// ---- scalar intrinsics ---------------------------------------------------------------------
// (the emulated build uses -ffp-contract=off: a*b+c is never fused, like --fmad=false)
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
static inline int __ffs(int v) { return __builtin_ffs(v); }
static inline int __clz(int v) { return v == 0 ? 32 : __builtin_clz((unsigned)v); }
static inline unsigned long long __umul64hi(unsigned long long a, unsigned long long b) {
    return (unsigned long long)(((unsigned __int128)a * b) >> 64);
}
template <class T>
static inline T __ldg(const T* p) {
    return *p;
}
template <class T>
static inline T __ldcg(const T* p) {
    return *p;
}
template <class T>
static inline T __ldcs(const T* p) {
    return *p;
}
template <class T, class U>
static inline void __stcg(T* p, U v) {
    *p = (T)v;
}
template <class T, class U>
static inline void __stcs(T* p, U v) {
    *p = (T)v;
}
static inline double cospi(double x) { return std::cos(3.14159265358979323846 * x); }
static inline double sinpi(double x) { return std::sin(3.14159265358979323846 * x); }
static inline long long __double_as_longlong(double v) { return cuemu::from_bits<long long>(cuemu::to_bits(v)); }
static inline double __longlong_as_double(long long v) { return cuemu::from_bits<double>(cuemu::to_bits(v)); }

// atomics: one OS thread runs every CUDA thread, so plain read-modify-write is atomic
template <class T, class U>
static inline T atomicAdd(T* p, U v) {
    const T old = *p;
    *p = (T)(old + (T)v);
    cuemu::note_progress();
    return old;
}
template <class T, class U>
static inline T atomicOr(T* p, U v) {
    const T old = *p;
    *p = (T)(old | (T)v);
    cuemu::note_progress();
    return old;
}
template <class T, class U>
static inline T atomicMax(T* p, U v) {
    const T old = *p;
    if ((T)v > old) *p = (T)v;
    cuemu::note_progress();
    return old;
}
template <class T, class U>
static inline T atomicMin(T* p, U v) {
    const T old = *p;
    if ((T)v < old) *p = (T)v;
    cuemu::note_progress();
    return old;
}
template <class T, class U, class V>
static inline T atomicCAS(T* p, U cmp, V val) {
    const T old = *p;
    if (old == (T)cmp) *p = (T)val;
    cuemu::note_progress();
    return old;
}
template <class T, class U>
static inline T atomicExch(T* p, U v) {
    const T old = *p;
    *p = (T)v;
    cuemu::note_progress();
    return old;
}
