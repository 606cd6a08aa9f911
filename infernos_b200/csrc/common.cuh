// Shared helpers for the infernos_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>
#include <utility>

namespace b2 {

extern thread_local std::string g_last_error;
extern std::atomic<uint64_t> g_launches;

int set_error(const char *fmt, ...);

inline void count_launch(int n = 1) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

#define B2_CUDA_OK(expr)                                                                          \
    do {                                                                                          \
        cudaError_t _e = (expr);                                                                  \
        if (_e != cudaSuccess)                                                                    \
            return b2::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define B2_LAUNCH_OK(name)                                                                        \
    do {                                                                                          \
        cudaError_t _e = cudaGetLastError();                                                      \
        if (_e != cudaSuccess)                                                                    \
            return b2::set_error("launch of %s failed: %s (%s:%d)", name, cudaGetErrorString(_e), __FILE__, __LINE__); \
        b2::count_launch();                                                                       \
    } while (0)

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// Programmatic dependent launch (PDL).  The tail is a chain of ~42 kernels on one stream (a decoder step: ~74): launched with the
// programmatic-stream-serialization attribute, kernel N+1 may become resident while kernel N is still running; it announces itself
// (pdl_trigger) and must not touch anything kernel N produces -- or still reads -- before pdl_wait returns (kernel N complete and flushed).
// Without the attribute both instructions are no-ops, so every kernel can carry them.  B2_PDL=0 launches everything the plain way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// A kernel launched with the attribute is RESIDENT while its predecessor still runs: nothing a predecessor writes may be read before
// griddepcontrol.wait returns.  The compiler does not know that -- a load through a `const __restrict__` parameter is an ld.global.nc of memory it takes to
// be immutable for the kernel's lifetime, and nvcc did hoist one above the wait (k_attend's rowslot[m], read a pass too early: a 2.8 dB loss in
// tests/test_gpu_decoder.py the day the kernel body changed).  So pdl_wait(p, q, ...) passes every pointer it is given through an empty asm AFTER the
// wait: loads through them depend on that asm's output and cannot be scheduled above it.  Hand it every pointer to data another kernel produces;
// tests/test_host_api.py scans the SASS of the library for global loads ahead of the wait.
template <typename P> __device__ __forceinline__ void pdl_fresh(P &p) { asm volatile("" : "+l"(p) : : "memory"); }
template <typename... P> __device__ __forceinline__ void pdl_wait(P &...p) {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    (pdl_fresh(p), ...);
}
#endif
bool pdl_enabled();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args &&...args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int sm_count();   // SMs of the current device (cached)

}  // namespace b2
