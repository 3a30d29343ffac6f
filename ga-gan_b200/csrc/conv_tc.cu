// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a, fp32 accuracy through 3xTF32.
//
// Serves the stride-1 convolutions of the StyleGAN2 hot path (3x3 pad 1 and 1x1: SynthesisLayer conv1, ToRGB-sized
// 1x1s with O>=16, Discriminator conv0, and their data gradients, which are stride-1 correlations with the transposed,
// flipped kernel), i.e. the cuDNN calls behind conv2d_gradfix (torch_utils/ops/conv2d_gradfix.py:141-146) plus the
// modulated_conv2d scale passes (training/networks.py:642,648-651) folded into the operand conversion (in_scale) and
// the epilogue (out_scale).
//
// GEMM view      D[m = output pixel, n = out channel] = sum_{k = (tap, in channel)} A[m,k] * B[n,k]
// CTA tile       16x16 output pixels (two 8x16 UMMA M-tiles of 128 pixels) x NT<=128 output channels.
// A operand      the input halo tile (18x20 pixels x 16 channels per K-block) arrives by ONE 4-D TMA box load with
//                hardware zero fill outside the image (= the conv padding), is converted ONCE by four SIMT warps
//                (x in_scale, split into tf32 hi + lo) into the UMMA no-swizzle K-major layout
//                [chunk of 4 channels][pixel][4 channels]; in that layout rows (pixels) are 16 B apart, so each of the
//                nine filter taps is just a different descriptor START ADDRESS into the same tile -- the im2col is free
//                and every input element is converted once, not nine times.
// B operand      weights pre-split (hi/lo) and pre-packed per (n-tile, K-block, tap) by pack_weights_kernel into the
//                exact shared-memory image, streamed by 1-D bulk TMA copies through a 4-stage ring.
// MMA            tcgen05.mma.cta_group::1.kind::tf32, M=128, N=NT, K=8; accumulators in TMEM (2 x NT columns);
//                3 products per MAC (hi*hi + hi*lo + lo*hi) reproduce fp32 to ~1e-6.
// Epilogue       tcgen05.ld 32x32b -> registers -> x out_scale -> NCHW stores (32 B runs per lane group).
// Warp roles     w0 TMA(x)  w1 TMA(weights)  w2 MMA issue + TMEM alloc  w3..w6 convert, then epilogue.
#include "common.cuh"
#include <cuda.h>
#include <stdlib.h>

namespace {

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
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000LL) {   // ~2 s: a pipeline bug must not hang the GPU
            // distinguishable from a hardware fault: a watchdog expiry shows up as "illegal memory access" (null store)
            *reinterpret_cast<volatile int*>(8) = (int)bar;
            __trap();
        }
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

// UMMA shared-memory descriptor, SWIZZLE_NONE ("interleave") canonical layouts (cute/arch/mma_sm100_desc.hpp):
// bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, [46,48) version=1, [61,64) layout type 0.
//   K-major : ((8,m),(4,2)) : rows 16 B apart inside a core matrix, SBO between 8-row groups, LBO between the 16-byte K chunks.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
// Instruction descriptor (UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format TF32 (2) @7/@10, a/b major @15/@16 (0 = K),
// N>>3 @17, M>>4 @24.
__host__ __device__ constexpr uint32_t umma_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// 3xTF32 operand split.  Both parts are rounded to nearest tf32 (cvt.rna) rather than left to the tensor core's
// truncation: |lo| <= 2^-12 |v| and lo itself carries a 2^-12 relative rounding error, so hi*hi + hi*lo + lo*hi
// reproduces the fp32 product to ~2^-22 (the dropped lo*lo term is 2^-24).
__device__ __forceinline__ void split_tf32(float v, float& hi, float& lo) {
    uint32_t h, l;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
    hi = __uint_as_float(h);
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(l) : "f"(v - hi));
    lo = __uint_as_float(l);
}

// ------------------------------------------------------------------------------------------------ weight packing
constexpr int KB_CH = 16;   // input channels per K-block (two UMMA K=8 steps)

struct PackP {
    const float* w; float* wp;
    int O, I, K, NT, n_tiles, num_kb, flip, w_is_IO;
};

// wp[n_tile][kb][tap][half: hi,lo][chunk 0..3][n 0..NT-1][4 channels]  -- per (n_tile,kb,tap) exactly the smem image
__global__ void pack_weights_kernel(PackP p) {
    const int KK = p.K * p.K;
    const int64_t total = (int64_t)p.n_tiles * p.num_kb * KK * 4 * p.NT;     // one thread = one (chunk, n) = 4 channels, hi and lo
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(idx % p.NT); int64_t r = idx / p.NT;
        int chunk = (int)(r % 4); r /= 4;
        int tap = (int)(r % KK); r /= KK;
        int kb = (int)(r % p.num_kb); int nt = (int)(r / p.num_kb);
        int o = nt * p.NT + n;
        int ky = tap / p.K, kx = tap - ky * p.K;
        if (p.flip) { ky = p.K - 1 - ky; kx = p.K - 1 - kx; }
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int i = kb * KB_CH + chunk * 4 + j;
            float v = 0.f;
            if (o < p.O && i < p.I) {
                int64_t src = p.w_is_IO ? (((int64_t)i * p.O + o) * p.K + ky) * p.K + kx : (((int64_t)o * p.I + i) * p.K + ky) * p.K + kx;
                v = __ldg(p.w + src);
            }
            split_tf32(v, hi[j], lo[j]);
        }
        int64_t blk = (((int64_t)nt * p.num_kb + kb) * KK + tap) * (2 * 4 * p.NT * 4);
        int64_t off = ((int64_t)chunk * p.NT + n) * 4;
        *reinterpret_cast<float4*>(p.wp + blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(p.wp + blk + 4 * p.NT * 4 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ------------------------------------------------------------------------------------------------ the conv kernel
constexpr int TILE_W = 16, TILE_H = 16;   // output pixels per CTA: two 8x16 UMMA M-tiles side by side
constexpr int W_STAGES = 4;
constexpr int NUM_THREADS = 7 * 32;
constexpr int CVT_THREADS = 128;

struct TcP {
    const float* wp; float* y; const float* in_scale; const float* out_scale;
    int Nimg, I, O, OH, OW, K, pad_y, pad_x;
    int tiles_x, tiles_y, NT, num_kb, nprod;
    int boxW, boxH;            // converted halo tile in pixels; boxW is its smem pixel pitch
    int rawW;                  // width of the raw TMA box (>= boxW: the box must start on a 16-byte boundary in global memory)
    uint32_t tmem_cols;
    int dbg;                   // GG_TC_DBG bitmask (debug experiments only)
};

struct SmemLayout {   // byte offsets from the 128-byte aligned dynamic smem base
    uint32_t raw[2], cvt[2][2], wst[W_STAGES], bars, tmem_slot, total;
};

__host__ __device__ inline SmemLayout make_layout(int boxW, int boxH, int rawW, int NT) {
    SmemLayout L;
    uint32_t tile = (uint32_t)(KB_CH * boxW * boxH * 4), rtile = (uint32_t)(KB_CH * rawW * boxH * 4);
    tile = (tile + 127) & ~127u;
    rtile = (rtile + 127) & ~127u;
    uint32_t off = 0;
    for (int s = 0; s < 2; ++s) { L.raw[s] = off; off += rtile; }
    for (int s = 0; s < 2; ++s) for (int h = 0; h < 2; ++h) { L.cvt[s][h] = off; off += tile; }
    for (int s = 0; s < W_STAGES; ++s) { L.wst[s] = off; off += (uint32_t)(2 * 4 * NT * 16); }
    L.bars = off; off += 256;          // 2 raw_full, 2 raw_empty, 2 cvt_full, 2 cvt_empty, W_STAGES w_full, W_STAGES w_empty, acc_full
    L.tmem_slot = off; off += 16;
    L.total = off;
    return L;
}

__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap xmap, TcP p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const SmemLayout L = make_layout(p.boxW, p.boxH, p.rawW, p.NT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t bar0 = base + L.bars;
    auto BAR_RAW_FULL = [&](int s) { return bar0 + 8u * s; };
    auto BAR_RAW_EMPTY = [&](int s) { return bar0 + 8u * (2 + s); };
    auto BAR_CVT_FULL = [&](int s) { return bar0 + 8u * (4 + s); };
    auto BAR_CVT_EMPTY = [&](int s) { return bar0 + 8u * (6 + s); };
    auto BAR_W_FULL = [&](int s) { return bar0 + 8u * (8 + s); };
    auto BAR_W_EMPTY = [&](int s) { return bar0 + 8u * (8 + W_STAGES + s); };
    const uint32_t BAR_ACC_FULL = bar0 + 8u * (8 + 2 * W_STAGES);

    // tile coordinates
    int bid = blockIdx.x;
    const int tx = bid % p.tiles_x; bid /= p.tiles_x;
    const int ty = bid % p.tiles_y;
    const int img = bid / p.tiles_y;
    const int nt = blockIdx.y;
    const int ox0 = tx * TILE_W, oy0 = ty * TILE_H;
    const int KK = p.K * p.K;
    const int npix = p.boxW * p.boxH;
    const int rpix = p.rawW * p.boxH;
    const uint32_t raw_bytes = (uint32_t)(KB_CH * rpix * 4);
    // TMA box origin: x aligned down to 4 floats (16 B) and both clamped to >= 0; the converter undoes the shift
    const int cx = max((ox0 - p.pad_x) & ~3, 0), cy = max(oy0 - p.pad_y, 0);
    const uint32_t w_bytes = (uint32_t)(2 * 4 * p.NT * 16);

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(BAR_RAW_FULL(s), 1); mbar_init(BAR_RAW_EMPTY(s), CVT_THREADS);
            mbar_init(BAR_CVT_FULL(s), CVT_THREADS); mbar_init(BAR_CVT_EMPTY(s), 1);
        }
        for (int s = 0; s < W_STAGES; ++s) { mbar_init(BAR_W_FULL(s), 1); mbar_init(BAR_W_EMPTY(s), 1); }
        mbar_init(BAR_ACC_FULL, 1);
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(base + L.tmem_slot, p.tmem_cols);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + L.tmem_slot);

    if (warp == 0) {
        // ===== x producer: one 4-D TMA box per K-block.  Boxes that stick out on the right / bottom are zero-filled by the
        // TMA unit (= the conv padding).  The innermost start coordinate must land on a 16-byte boundary (an unaligned or
        // negative one faults with "illegal instruction"), so the box starts at the aligned, clamped (cx, cy) and the
        // converter shifts it back and re-creates the left / top padding (see cvt_src below).
        if (lane == 0) {
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int s = kb & 1;
                mbar_wait(BAR_RAW_EMPTY(s), ((kb >> 1) & 1) ^ 1);
                if (p.dbg & 16) { mbar_arrive(BAR_RAW_FULL(s)); continue; }
                mbar_expect_tx(BAR_RAW_FULL(s), raw_bytes);
                tma_load_4d(base + L.raw[s], &xmap, BAR_RAW_FULL(s), cx, cy, kb * KB_CH, img);
            }
        }
    } else if (warp == 1) {
        // ===== weight producer: one bulk copy (hi+lo image of this tap) per ring stage
        if (lane == 0) {
            const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)nt * p.num_kb * KK * w_bytes;
            const int total = p.num_kb * KK;
            for (int g = 0; g < total; ++g) {
                const int s = g % W_STAGES;
                mbar_wait(BAR_W_EMPTY(s), ((g / W_STAGES) & 1) ^ 1);
                mbar_expect_tx(BAR_W_FULL(s), w_bytes);
                bulk_load(base + L.wst[s], src + (size_t)g * w_bytes, w_bytes, BAR_W_FULL(s));
            }
        }
    } else if (warp == 2) {
        // ===== MMA issuer (one thread)
        if (lane == 0) {
            const uint32_t idesc = umma_idesc_tf32(128, p.NT, 0, 0);
            const uint32_t a_lbo = (uint32_t)npix * 16u, a_sbo = (uint32_t)p.boxW * 16u;
            const uint32_t b_lbo = (uint32_t)p.NT * 16u, b_sbo = 128u;
            int g = 0;
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const int cs = kb & 1;
                mbar_wait(BAR_CVT_FULL(cs), (kb >> 1) & 1);
                tc_fence_after();
                const uint32_t a_hi0 = base + L.cvt[cs][0], a_lo0 = base + L.cvt[cs][1];
                for (int tap = 0; tap < KK; ++tap, ++g) {
                    const int ws = g % W_STAGES;
                    mbar_wait(BAR_W_FULL(ws), (g / W_STAGES) & 1);
                    tc_fence_after();
                    const int ky = tap / p.K, kx = tap - ky * p.K;
                    const uint32_t b_hi0 = base + L.wst[ws], b_lo0 = b_hi0 + 4u * p.NT * 16u;
#pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
#pragma unroll
                        for (int ks = 0; ks < 2; ++ks) {
                            const uint32_t a_off = (uint32_t)(((p.dbg & 1) ? 8 * sub : (ky * p.boxW + 8 * sub + kx)) * 16) + (uint32_t)(2 * ks) * a_lbo;
                            const uint32_t b_off = (uint32_t)(2 * ks) * b_lbo;
                            const uint64_t a_hi = umma_desc(a_hi0 + a_off, a_lbo, a_sbo);
                            const uint64_t b_hi = umma_desc(b_hi0 + b_off, b_lbo, b_sbo);
                            const uint32_t d = tmem_base + (uint32_t)(sub * p.NT);
                            if (p.dbg & 8) continue;
                            umma_tf32(d, a_hi, b_hi, idesc, (kb | tap | ks) != 0 ? 1u : 0u);
                            if (p.nprod == 3) {
                                const uint64_t a_lo = umma_desc(a_lo0 + a_off, a_lbo, a_sbo);
                                const uint64_t b_lo = umma_desc(b_lo0 + b_off, b_lbo, b_sbo);
                                umma_tf32(d, a_hi, b_lo, idesc, 1u);
                                umma_tf32(d, a_lo, b_hi, idesc, 1u);
                            }
                        }
                    }
                    umma_commit(BAR_W_EMPTY(ws));      // frees the weight stage when these MMAs have read it
                }
                umma_commit(BAR_CVT_EMPTY(cs));        // frees the converted tile
            }
            umma_commit(BAR_ACC_FULL);
        }
    } else {
        // ===== converter warps (w3..w6), then epilogue
        const int ct = threadIdx.x - 3 * 32;            // 0..127
        const float* sc = p.in_scale ? p.in_scale + (size_t)img * p.I : nullptr;
        // source index inside the raw box for each converted pixel this thread owns (-1 = conv padding -> 0)
        constexpr int MAXPX = 4;
        int cvt_src[MAXPX];
        {
            const int dx = (ox0 - p.pad_x) - cx, dy = (oy0 - p.pad_y) - cy;     // converted (r,c) <- raw (r+dy, c+dx)
#pragma unroll
            for (int t = 0; t < MAXPX; ++t) {
                const int px = ct + t * CVT_THREADS;
                const int rr = px / p.boxW + dy, cc = px % p.boxW + dx;
                cvt_src[t] = (px < npix && rr >= 0 && cc >= 0 && rr < p.boxH && cc < p.rawW) ? rr * p.rawW + cc : -1;
            }
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
            const int s = kb & 1;
            mbar_wait(BAR_RAW_FULL(s), (kb >> 1) & 1);
            mbar_wait(BAR_CVT_EMPTY(s), ((kb >> 1) & 1) ^ 1);
            const float* raw = reinterpret_cast<const float*>(gbase + L.raw[s]);
            float4* hi = reinterpret_cast<float4*>(gbase + L.cvt[s][0]);
            float4* lo = reinterpret_cast<float4*>(gbase + L.cvt[s][1]);
#pragma unroll 1
            for (int chunk = 0; chunk < 4; ++chunk) {
                float s0 = 1.f, s1 = 1.f, s2 = 1.f, s3 = 1.f;
                if (sc) {
                    const int c = kb * KB_CH + chunk * 4;
                    s0 = (c + 0 < p.I) ? __ldg(sc + c + 0) : 0.f; s1 = (c + 1 < p.I) ? __ldg(sc + c + 1) : 0.f;
                    s2 = (c + 2 < p.I) ? __ldg(sc + c + 2) : 0.f; s3 = (c + 3 < p.I) ? __ldg(sc + c + 3) : 0.f;
                }
                const float* r0 = raw + (chunk * 4) * rpix;
#pragma unroll
                for (int t = 0; t < MAXPX; ++t) {
                    const int px = ct + t * CVT_THREADS;
                    if (px >= npix) break;
                    const int src = cvt_src[t];
                    float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
                    if (src >= 0) {
                        split_tf32(r0[src] * s0, h.x, l.x);
                        split_tf32(r0[rpix + src] * s1, h.y, l.y);
                        split_tf32(r0[2 * rpix + src] * s2, h.z, l.z);
                        split_tf32(r0[3 * rpix + src] * s3, h.w, l.w);
                    }
                    hi[chunk * npix + px] = h;
                    lo[chunk * npix + px] = l;
                }
            }
            fence_proxy_async();                        // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(BAR_CVT_FULL(s));
            mbar_arrive(BAR_RAW_EMPTY(s));
        }

        // epilogue: this warp may touch TMEM lanes 32*(warp%4) .. +31
        mbar_wait(BAR_ACC_FULL, 0);
        tc_fence_after();
        const int q = warp & 3;
        const int m = q * 32 + lane;                    // accumulator row = pixel inside the 8x16 sub-tile
        const int r = m >> 3, c = m & 7;
        const int oy = oy0 + r;
        const int n0 = nt * p.NT;
        const size_t plane = (size_t)p.OH * p.OW;
        for (int sub = 0; sub < 2; ++sub) {
            const int ox = ox0 + 8 * sub + c;
            const bool pix_ok = (oy < p.OH) && (ox < p.OW);
            float* yp = p.y + ((size_t)img * p.O + n0) * plane + (size_t)oy * p.OW + ox;
            for (int cb = 0; cb < p.NT; cb += 16) {
                uint32_t v[16];
                tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(sub * p.NT + cb), v);
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int o = n0 + cb + j;
                    if (pix_ok && o < p.O) {
                        float val = __uint_as_float(v[j]);
                        if (p.out_scale) val *= __ldg(p.out_scale + (size_t)img * p.O + o);
                        yp[(size_t)(cb + j) * plane] = val;
                    }
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, p.tmem_cols);
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::atomic<int> state{0};
    if (state.load() == 2) return fn;
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(ptr);
    state.store(2);
    return fn;
}

}  // namespace

namespace gg {

bool conv2d_tc_eligible(int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int stride, int pad_y, int pad_x,
                        int transposed) {
    if (stride != 1 || KH != KW || (KH != 1 && KH != 3)) return false;
    if (transposed && (pad_y > KH - 1 || pad_x > KW - 1)) return false;
    if (N < 1 || I < 16 || O < 16) return false;       // 3-channel fromRGB / ToRGB stay on the FFMA path (HBM-bound, AI ~ 1.4)
    if (W % 4 != 0) return false;                      // TMA global strides must be multiples of 16 bytes
    if (OW < 16 || OH < 16) return false;              // 4x4 / 8x8 maps: < 0.3 % of the FLOPs, served by the FFMA kernel
    return true;
}

int conv2d_tc(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int K, int /*KW*/, int pad_y, int pad_x,
              int flip_w, int w_is_IO, const float* in_scale, const float* out_scale, int nprod, cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) { set_error("conv2d(tc): x must be 16-byte aligned"); return GG_EINVAL; }
    const int OH = H + 2 * pad_y - K + 1, OW = W + 2 * pad_x - K + 1;
    const int NT = O >= 128 ? 128 : ((O + 15) / 16) * 16;
    const int n_tiles = (O + NT - 1) / NT;
    const int num_kb = (I + KB_CH - 1) / KB_CH;
    const int KK = K * K;
    const char* dbg_env = getenv("GG_TC_DBG");
    const int dbg = dbg_env ? atoi(dbg_env) : 0;
    const int boxW = ((TILE_W + K - 1) + 3) / 4 * 4, boxH = TILE_H + K - 1;
    const int rawW = (TILE_W + K - 1 + ((4 - pad_x % 4) % 4) + 3) / 4 * 4;   // aligned-down start => up to 3 extra columns

    // 1. pack + split the weights (tiny; stream-ordered scratch)
    const size_t wp_floats = (size_t)n_tiles * num_kb * KK * 2 * 4 * NT * 4;
    float* wp = nullptr;
    GG_CUDA(cudaMallocAsync(&wp, wp_floats * sizeof(float), st));
    PackP pp{w, wp, O, I, K, NT, n_tiles, num_kb, flip_w, w_is_IO};
    {
        int64_t threads = (int64_t)n_tiles * num_kb * KK * 4 * NT;
        int grid = (int)((threads + 255) / 256);
        if (grid > GG_NUM_SMS * 8) grid = GG_NUM_SMS * 8;
        pack_weights_kernel<<<grid, 256, 0, st>>>(pp);
        int rc = check_launch("conv2d(tc) pack_weights");
        if (rc != GG_OK) { cudaFreeAsync(wp, st); return rc; }
    }

    // 2. TMA descriptor of x as a 4-D tensor {W, H, C, N}; box {boxW, boxH, 16, 1}; OOB elements read as zero
    EncodeTiledFn encode = get_encode_fn();
    if (!encode) { cudaFreeAsync(wp, st); set_error("conv2d(tc): cuTensorMapEncodeTiled is unavailable"); return GG_ECUDA; }
    CUtensorMap xmap;
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)I, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * I * 4};
    cuuint32_t box[4] = {(cuuint32_t)rawW, (cuuint32_t)boxH, (cuuint32_t)KB_CH, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { cudaFreeAsync(wp, st); set_error("conv2d(tc): cuTensorMapEncodeTiled failed (%d)", (int)cr); return GG_ECUDA; }

    // 3. launch
    TcP p{wp, y, in_scale, out_scale, N, I, O, OH, OW, K, pad_y, pad_x, (OW + TILE_W - 1) / TILE_W, (OH + TILE_H - 1) / TILE_H,
          NT, num_kb, (nprod == GG_PREC_TF32X1 || (dbg & 2)) ? 1 : 3, boxW, boxH, rawW, 0, dbg};
    uint32_t cols = 32;
    while (cols < (uint32_t)(2 * NT)) cols <<= 1;
    p.tmem_cols = cols;
    const SmemLayout L = make_layout(boxW, boxH, rawW, NT);
    const size_t smem = L.total + 128;
    static std::atomic<int> attr_set{0};
    if (!attr_set.load()) {
        GG_CUDA(cudaFuncSetAttribute(conv_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        attr_set.store(1);
    }
    dim3 grid((unsigned)((size_t)p.tiles_x * p.tiles_y * N), (unsigned)n_tiles);
    conv_tc_kernel<<<grid, NUM_THREADS, smem, st>>>(xmap, p);
    int rc = check_launch("conv2d(tc)");
    cudaFreeAsync(wp, st);
    return rc;
}

bool wgrad_tc_eligible(int, int, int, int, int, int, int, int, int, int, int, int) { return false; }
int wgrad_tc(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, const float*, const float*,
             int, cudaStream_t) {
    set_error("wgrad(tc): not built yet");
    return GG_EUNSUPPORTED;
}

}  // namespace gg
