// TEST INFRASTRUCTURE: a minimal CUDA execution-model shim for running SIMT kernel SOURCE on the CPU (g++ -std=c++20 -pthread).
//
// One std::thread per CUDA thread of a block, blocks one after the other.  __syncthreads() / __syncwarp() are real barriers over the
// block's / the warp's threads, __shared__ variables are function-local statics (one block at a time), blockIdx / threadIdx are
// thread-local.  Between two barriers the threads interleave freely -- as the lanes of a warp may since independent thread
// scheduling -- so a missing __syncwarp() is a data race that ThreadSanitizer reports (build the test with -fsanitize=thread).
// Covers kernels written with: __global__/__device__ functions and lambdas, __shared__ arrays, __ldg, float2/float4, barriers,
// full-mask __shfl_xor_sync (an exchange through a per-warp scratch line between two warp barriers: a lane that does not take part
// hangs the run, as a divergent full-mask shuffle would be undefined on the device) and float atomicAdd (std::atomic_ref).
// Not covered: votes, partial-mask shuffles, TMA, mbarriers, tensor cores.
#pragma once
#include <algorithm>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
static inline float2 make_float2(float a, float b) { float2 v; v.x = a; v.y = b; return v; }
static inline float4 make_float4(float a, float b, float c, float d) { float4 v; v.x = a; v.y = b; v.z = c; v.w = d; return v; }
using std::max;
using std::min;

static thread_local dim3 blockIdx, threadIdx;
static dim3 blockDim, gridDim;
static std::barrier<>* shim_block_barrier = nullptr;
static std::vector<std::unique_ptr<std::barrier<>>> shim_warp_barriers;
static inline void __syncthreads() { shim_block_barrier->arrive_and_wait(); }
static inline void __syncwarp() { shim_warp_barriers[threadIdx.x / 32]->arrive_and_wait(); }

static float shim_dynamic_smem[12288];         // what `extern __shared__ float name[];` is bound to (48 KB)
static float shim_shfl_line[32][32];          // [warp of the block][lane]
static inline float __shfl_xor_sync(unsigned /*mask: full*/, float v, int lane_mask) {
    const unsigned w = threadIdx.x / 32, l = threadIdx.x % 32;
    shim_shfl_line[w][l] = v;
    __syncwarp();
    const float r = shim_shfl_line[w][l ^ (unsigned)lane_mask];
    __syncwarp();
    return r;
}
static inline float atomicAdd(float* p, float v) { return std::atomic_ref<float>(*p).fetch_add(v, std::memory_order_relaxed); }

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T* p) { return *p; }

typedef void* cudaStream_t;
typedef void* gg_stream_t;
typedef int cudaError_t;
#define cudaSuccess 0
static inline const char* cudaGetErrorString(cudaError_t) { return "shim"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t bytes, cudaStream_t) { memset(p, v, bytes); return cudaSuccess; }
#define GG_API
#define GG_OK 0
#define GG_EINVAL (-1)
#define GG_ECUDA (-2)
#ifndef GG_NUM_SMS
#define GG_NUM_SMS 148
#endif
static char shim_err[512];
#ifdef SHIM_MULTI_UNIT      // several translation units linked into one library (tests/emulated_lib.py): api.cu's own definitions are used
namespace gg {
void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
static inline int check_launch(const char*) { g_launches.fetch_add(1, std::memory_order_relaxed); return GG_OK; }
}  // namespace gg
#else
namespace gg {
static inline void set_error(const char* fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(shim_err, sizeof shim_err, fmt, ap); va_end(ap); }
static inline int check_launch(const char*) { return GG_OK; }
}  // namespace gg
#endif
#define GG_REQUIRE(cond, ...) do { if (!(cond)) { gg::set_error(__VA_ARGS__); return GG_EINVAL; } } while (0)
#define GG_CUDA(call) do { if ((call) != cudaSuccess) return GG_ECUDA; } while (0)

static long shim_blocks_launched = 0, shim_blocks_total = 0, shim_block_threads = 0;
static void (*shim_block_hook)() = nullptr;         // called after every block (tc_cpu_shim.h resets its per-CTA objects there)

// SHIM_LAUNCH((kernel<...>), grid, block, args...): what `kernel<...><<<grid, block, 0, stream>>>(args...)` does.  The block's threads are
// created once per launch and walk the blocks together (two rendezvous per block instead of a thread spawn per CUDA thread); the
// block / warp barriers are fresh objects for every block, because a thread that has returned from the kernel drops out of them.
static std::unique_ptr<std::barrier<>> shim_block_barrier_owner;
template <class K, class... A>
static void shim_launch(K kernel, dim3 grid, dim3 block, A... args) {
    gridDim = grid; blockDim = block;
    shim_blocks_launched = (long)grid.x * grid.y * grid.z;      // of the last launch
    shim_blocks_total += shim_blocks_launched;                  // of all launches since shim_reset()
    shim_block_threads = block.x;
    const unsigned nt = block.x * block.y * block.z;
    if (shim_blocks_launched == 0 || nt == 0) return;
    auto fresh = [nt]() {
        shim_block_barrier_owner.reset(new std::barrier<>(nt));
        shim_block_barrier = shim_block_barrier_owner.get();
        shim_warp_barriers.clear();
        for (unsigned w = 0; w * 32 < nt; ++w) shim_warp_barriers.emplace_back(new std::barrier<>(std::min(32u, nt - w * 32)));
    };
    fresh();
    std::barrier<> rendezvous(nt);
    std::vector<std::thread> threads;
    for (unsigned t = 0; t < nt; ++t)
        threads.emplace_back([=, &rendezvous, &fresh]() {
            for (unsigned bz = 0; bz < grid.z; ++bz)
                for (unsigned by = 0; by < grid.y; ++by)
                    for (unsigned bx = 0; bx < grid.x; ++bx) {
                        blockIdx = dim3(bx, by, bz);
                        threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
                        kernel(args...);
                        shim_warp_barriers[t / 32]->arrive_and_drop();       // a thread that has returned no longer takes part in barriers
                        shim_block_barrier->arrive_and_drop();
                        rendezvous.arrive_and_wait();                        // the whole block is done
                        if (t == 0) { if (shim_block_hook) shim_block_hook(); fresh(); }
                        rendezvous.arrive_and_wait();                        // fresh barriers, __shared__ statics free for the next block
                    }
        });
    for (auto& th : threads) th.join();
}
// the kernel is called by NAME inside a generic lambda, so that overload resolution and argument-dependent lookup see what the
// `<<<>>>` call of the source sees (conv_thin.cu launches a kernel that shares its name with the host function around the launch)
#define SHIM_UNPAREN(...) __VA_ARGS__
#define SHIM_LAUNCH(kernel, grid, block, ...) shim_launch([=](auto... shim_a) { SHIM_UNPAREN kernel(shim_a...); }, dim3(grid), dim3(block), __VA_ARGS__)
#ifndef SHIM_MULTI_UNIT
extern "C" long shim_blocks() { return shim_blocks_launched; }
extern "C" long shim_blocks_since_reset() { return shim_blocks_total; }
extern "C" void shim_reset() { shim_blocks_launched = shim_blocks_total = 0; shim_err[0] = 0; }
extern "C" long shim_threads() { return shim_block_threads; }
extern "C" const char* shim_error() { return shim_err; }
#endif
