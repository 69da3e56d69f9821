// Shared helpers for libct3d (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <atomic>

#include "../../include/ct3d.h"

namespace ct {

void set_error(const char* fmt, ...);
extern std::atomic<unsigned long long> g_launches;
extern thread_local int g_reserved_sms;      // see ct_set_reserved_sms (per host thread)

inline int check_cuda(cudaError_t e, const char* what) {
    if (e != cudaSuccess) {
        // The tcgen05 kernels bound every mbarrier wait (tc_ptx.cuh: trap after 4e9 clocks) so that a protocol error ends
        // as a launch failure instead of a hung GPU; the error is sticky for the context and surfaces at a later call.
        const bool trapped = (e == cudaErrorLaunchFailure || e == cudaErrorIllegalInstruction || e == cudaErrorAssert);
        set_error("%s: %s%s", what, cudaGetErrorString(e),
                  trapped ? " (a kernel of an EARLIER launch trapped -- a bounded mbarrier wait of a convolution kernel timed out, or a "
                            "device-side check failed; the CUDA context is unusable from here on)" : "");
        return 1;
    }
    return 0;
}

// Count a kernel launch and check the launch status (asynchronous errors surface later).
#define CT_LAUNCHED(name)                                                        \
    do {                                                                         \
        ::ct::g_launches.fetch_add(1, std::memory_order_relaxed);                \
        if (::ct::check_cuda(cudaGetLastError(), name)) return 1;                \
    } while (0)

#define CT_CUDA(expr)                                                            \
    do {                                                                         \
        if (::ct::check_cuda((expr), #expr)) return 1;                           \
    } while (0)

#define CT_REQUIRE(cond, ...)                                                    \
    do {                                                                         \
        if (!(cond)) {                                                           \
            ::ct::set_error(__VA_ARGS__);                                        \
            return 1;                                                            \
        }                                                                        \
    } while (0)

// Profiler tags (ct_profile_read)
enum ProfTag { PROF_CONV = 1, PROF_EM = 2, PROF_FFN = 3, PROF_LCN = 4, PROF_UNET_AUX = 5, PROF_WATERSHED = 6 };
struct ProfScope {
    ProfScope(int tag, cudaStream_t s);
    ~ProfScope();
    int tag_;
    cudaStream_t s_;
    bool on_;
    cudaEvent_t a_;
};

// Small host -> device uploads (problem descriptors) through a per-thread ring of PINNED staging slots.  A
// cudaMemcpyAsync from pageable memory is staged by the driver and "might be synchronous with respect to the host": in
// the frame pipeline the five descriptor uploads of a fit (one per PR-GLS repetition) tied the host to the EM chain and
// produced 60-140 ms gaps in its launch loop [measured].  From pinned memory the copy is a plain stream operation.
int stage_h2d(void* dst, const void* src, size_t bytes, cudaStream_t s);

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }
inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// Bump allocator over a caller-provided workspace.
struct Arena {
    char* base;
    size_t size;
    size_t off = 0;
    Arena(void* p, size_t n) : base(static_cast<char*>(p)), size(n) {}
    template <typename T>
    T* take(size_t count) {
        off = align_up(off, 256);
        T* r = reinterpret_cast<T*>(base + off);
        off += count * sizeof(T);
        return r;
    }
    bool ok() const { return off <= size; }
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace ct
