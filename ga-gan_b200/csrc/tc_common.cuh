// PTX wrappers shared by the tcgen05 kernels (conv_tc.cu, wgrad_tc.cu): mbarrier, TMA, TMEM, UMMA descriptors, 3xTF32 split.
#pragma once
#include "common.cuh"
#include <cuda.h>

namespace gg {
// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda link dependency); defined in conv_tc.cu
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn get_encode_fn();
unsigned long long* watchdog_host_record();          // api.cu: 64 bytes of host-mapped memory, zeroed (nullptr if unavailable)
}  // namespace gg

namespace ggtc {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Watchdog shared by the wait loops: off the fast path (the clock is read once per 1024 polls -- reading it around every wait
// was a visible share of the single-thread issue loops), ~4 s of SM clocks, then a trap instead of a hung GPU.  Before it traps
// the thread leaves a record in HOST-mapped memory (it survives the faulted context): which kernel family, which barrier (its
// shared-memory address), which CTA / thread, how long it waited.  gg_watchdog_report() formats it; the Python host appends it
// to the error it raises.  Every translation unit has its own copy of the pointer (no relocatable device code) and sets it
// with wd_arm() before its first launch on a device.
#ifndef GG_TU_TAG
#define GG_TU_TAG 0
#endif
static __device__ unsigned long long* gg_wd_ptr = nullptr;
// The expiry path has to stay TINY: it is inlined at every wait site of kernels whose consumer warps sit at the 168-register cap.  A
// first version (atomicCAS for "first expiry wins", 64-bit fields, a system fence) cost conv_tc 8 .. 232 bytes of spill stack and
// 11 % of the in-step convolution throughput (141 vs 159 TFLOP/s) for code that never runs; a __noinline__ call did not help (ABI
// frame + spills).  Now: four 32-bit volatile stores through one pointer (last expiry wins -- concurrent expiries tell the same story).
__device__ __forceinline__ void mbar_watchdog(uint32_t it, long long& t0, uint32_t bar) {
    if ((it & 1023u) != 1023u) return;
    const long long now = clock64();
    if (t0 == 0) t0 = now;
    else if (now - t0 > 8000000000LL) {
        volatile unsigned int* r = reinterpret_cast<volatile unsigned int*>(gg_wd_ptr);
        if (r != nullptr) {
            r[2] = blockIdx.x;
            r[3] = threadIdx.x;
            r[4] = bar;
            r[0] = 0x57440000u | (unsigned)GG_TU_TAG;                                  // "WD" + kernel family, written last
        }
        __trap();
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        mbar_watchdog(it, t0, bar);
    }
}
// One lane of a converged warp; ptxas treats the region guarded by elect.sync as single-threaded, which lets it keep the
// UMMA descriptors in uniform registers instead of broadcasting them per instruction.
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
// Pure spin on mbarrier.test_wait (no hardware suspend): lower wake-up latency for the fine-grained producer/consumer
// handshakes of the weight-gradient kernel, at the price of issue slots.
__device__ __forceinline__ void mbar_wait_spin(uint32_t bar, uint32_t parity) {
    long long t0 = 0;
    for (uint32_t it = 0;; ++it) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        mbar_watchdog(it, t0, bar);
    }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// the same load without the wait: issue several, then tmem_wait_ld() once, so that their latencies overlap
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t* v) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                   "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// registers -> TMEM: 16 consecutive 32-bit columns of this thread's lane (warp w may only touch lanes 32*(w%4) .. +31)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 :: "r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]),
                    "r"(v[9]), "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]) : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]: the A operand (M = 128 rows = TMEM lanes, K = 8 tf32 values = 8 consecutive columns) is read
// from tensor memory, so only the B rows cross the shared-memory port (cute SM100_MMA_TF32_TS)
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_NONE ("interleave") canonical layouts (cute/arch/mma_sm100_desc.hpp):
// bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout type 0.
//   K-major : ((8,m),(4,2)) : rows 16 B apart inside a core matrix, SBO between 8-row groups, LBO between the 16-byte K chunks.
// Instruction descriptor (UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format TF32 (2) @7/@10, a/b major @15/@16 (0 = K),
// N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 3xTF32 operand split.  Both parts are rounded to nearest tf32 (cvt.rna) rather than left to the tensor core's
// truncation: |lo| <= 2^-12 |v| and lo itself carries a 2^-12 relative rounding error, so hi*hi + hi*lo + lo*hi
// reproduces the fp32 product to ~2^-22 (the dropped lo*lo term is 2^-24).
__device__ __forceinline__ float round_tf32(float v) {
    // round-to-nearest (ties away from zero) to 10 explicit mantissa bits in the integer domain: 2 instructions.
    // (cvt.rna.tf32.f32 expands to a ~12-instruction sequence on sm_100a and dominated the operand converters.)
    return __uint_as_float((__float_as_uint(v) + 0x00001000u) & 0xFFFFE000u);
}
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    hi = round_tf32(v);
    lo = round_tf32(v - hi);          // v - hi is exact in fp32
}


// Expected-value compensation of the tensor core's truncating fp32 accumulator.  Every accumulate step chops the running
// sum P toward zero by 0.5 ulp(P) on average.  Because sign(P)*|P| = P the CONDITIONAL MEAN of the accumulated loss, given
// the final chunk sum S, is -beta * S * sum_j(P_j / S) whatever the signs of the data (a bridge from 0 to S has E[P_j] =
// S*j/m).  With nprod accumulates per hi*hi product and m hi*hi products per accumulator and chunk, sum_j P_j / S =
// nprod*(m+1)/2.  beta = 0.5 * E[ulp(P)/|P|] ~ 0.5 * 2^-23 * 0.72, calibrated on B200 (tools/tc_rounding.py: -1.37e-6 at
// m = 18, nprod = 3; wgrad -5.3e-6 at m = 64) to 5.0e-8.  The drained chunk sum is multiplied by 1 + kappa; what is left of
// the truncation is zero-mean noise of ~1e-7 (profiles/r1_tc_rounding.txt).
__device__ __forceinline__ float rz_compensation(int m, int nprod) {
    return 1.0f + 5.0e-8f * (float)nprod * 0.5f * (float)(m + 1);
}


// Optional fused epilogue of the convolution kernels: y = clamp(act(conv + (bias[o] + noise[pixel])) * gain), the bias_act pass
// (torch_utils/ops/bias_act.py; reference kernel bias_act.cu:56-142) of the layers whose convolution is the LAST operator of
// conv2d_resample.  act = 0: no epilogue; 1 linear, 2 relu, 3 lrelu (the reference's cuda_idx; same strict comparisons).
struct ConvEpilogue {
    const float* bias;          // [O] or null
    const float* noise;         // [OH*OW] (noise_bs == 0) or [N, OH*OW] (noise_bs == OH*OW) or null
    long long noise_bs;
    int act;
    float alpha, gain, clamp;   // clamp < 0: off
};
// The per-element work sits in the consumer warps' store loop, i.e. on the critical path of the small-K layers: keep it at
// add, compare, select, multiply (+ the clamp only when there is one).  act(v) * gain == v * (v > 0 ? gain : gain * slope) with slope 1
// (linear), 0 (relu) or alpha (lrelu); the product gain * alpha is rounded once instead of twice (<= 1 ulp from the two-step form).
struct EpilogueScalars { float g_pos, g_neg, clamp; };
__device__ __forceinline__ EpilogueScalars epilogue_scalars(const ConvEpilogue& e) {
    EpilogueScalars s;
    s.g_pos = e.gain;
    s.g_neg = e.act == 1 ? e.gain : (e.act == 2 ? 0.f : e.gain * e.alpha);
    s.clamp = e.clamp;
    return s;
}
__device__ __forceinline__ float epilogue_apply(float v, float bias_plus_noise, const EpilogueScalars& s) {
    v += bias_plus_noise;
    v *= (v > 0.f) ? s.g_pos : s.g_neg;
    if (s.clamp >= 0.f) v = (v > -s.clamp && v < s.clamp) ? v : (v >= 0.f) ? s.clamp : -s.clamp;
    return v;
}


// host: point this translation unit's watchdog at the host-mapped record (once per device)
static inline void wd_arm() {
    static std::atomic<uint64_t> armed{0};
    if (gg::done_on_this_device(armed)) return;
    unsigned long long* rec = gg::watchdog_host_record();
    if (rec != nullptr) cudaMemcpyToSymbol(gg_wd_ptr, &rec, sizeof(rec));
    gg::mark_done_on_this_device(armed);
}

}  // namespace ggtc
