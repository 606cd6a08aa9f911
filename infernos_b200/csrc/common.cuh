// Shared helpers for the infernos_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <atomic>

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

int sm_count();   // SMs of the current device (cached)

}  // namespace b2
