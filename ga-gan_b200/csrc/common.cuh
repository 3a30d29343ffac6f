// Shared host/device helpers for libgagan_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <atomic>
#include "../../include/gagan_b200.h"

#define GG_NUM_SMS 148  // B200: 2 dies x 74 SMs; grids are sized in multiples of this

namespace gg {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
        return GG_ECUDA;
    }
    return GG_OK;
}

// Per-device "done once" flag (function attributes such as the dynamic shared-memory opt-in belong to the device's context, and one
// process may drive several GPUs).  Usage:  if (!done_on_this_device(flag)) { ...; mark_done_on_this_device(flag); }
inline uint64_t current_device_bit() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0) dev = 0;
    return 1ull << (dev & 63);
}
inline bool done_on_this_device(const std::atomic<uint64_t>& flag) { return (flag.load(std::memory_order_acquire) & current_device_bit()) != 0; }
inline void mark_done_on_this_device(std::atomic<uint64_t>& flag) { flag.fetch_or(current_device_bit(), std::memory_order_release); }

#define GG_REQUIRE(cond, ...)            \
    do {                                 \
        if (!(cond)) {                   \
            gg::set_error(__VA_ARGS__);  \
            return GG_EINVAL;            \
        }                                \
    } while (0)

#define GG_CUDA(call)                                                              \
    do {                                                                           \
        cudaError_t e__ = (call);                                                  \
        if (e__ != cudaSuccess) {                                                  \
            gg::set_error("%s failed: %s", #call, cudaGetErrorString(e__));        \
            return GG_ECUDA;                                                       \
        }                                                                          \
    } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// floor division / positive modulo for possibly negative numerators (index math must be exact)
__host__ __device__ __forceinline__ int floordiv(int a, int b) {
    int q = a / b;
    return (a % b != 0 && ((a < 0) != (b < 0))) ? q - 1 : q;
}
__host__ __device__ __forceinline__ int posmod(int a, int b) {
    int r = a % b;
    return r < 0 ? r + b : r;
}

}  // namespace gg
