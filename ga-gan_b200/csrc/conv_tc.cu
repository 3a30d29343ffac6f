// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a, fp32-faithful through 3xTF32 + chunked accumulation.
//
// Serves every stride-1 correlation of the StyleGAN2 hot path with a k x k kernel, k in {1,2,3}: SynthesisLayer conv1,
// Discriminator conv0, 1x1 layers with >= 16 channels, their data gradients (stride-1 correlations with the transposed,
// flipped kernel) and -- through the phase-major (space-to-depth) formulation in torch_utils/ops/conv2d_resample.py --
// the stride-2 up/down layers, i.e. the cuDNN calls behind conv2d_gradfix (torch_utils/ops/conv2d_gradfix.py:141-146)
// plus the modulated_conv2d scale passes (training/networks.py:642,648-651) folded into the operand conversion
// (in_scale) and the epilogue (out_scale).
//
// GEMM view      D[m = output pixel, n = out channel] = sum_{k = (tap, in channel)} A[m,k] * B[n,k]
// CTA            persistent (one per SM), walks 16x16-pixel x NT-channel output tiles; NT in {32,64,128}.
// A operand      the input halo tile ((16+k-1)^2 pixels x 16 channels per K-block) arrives by ONE 4-D TMA box load with
//                hardware zero fill outside the image (= the conv padding), is converted ONCE by the consumer warps
//                (x in_scale, split into tf32 hi + lo, cvt.rna) into the UMMA no-swizzle K-major layout
//                [chunk of 4 channels][pixel][4 channels]; rows (pixels) are 16 B apart, so each filter tap is just a
//                different descriptor START ADDRESS into the same tile -- the im2col is free and every input element
//                is converted once, not k*k times.
// B operand      weights pre-split (hi/lo) and pre-packed per (n-tile, K-block, tap) by pack_weights_kernel into the exact
//                shared-memory image, streamed by 1-D bulk TMA copies through a 4-stage ring.  pack_weights_kernel also
//                flags all-zero (n-tile, K-block, tap) blocks; they are skipped by every role (the phase-major stride-2
//                formulation has 7 structurally zero blocks out of 16).
// MMA            tcgen05.mma.cta_group::1.kind::tf32, M=128, N=NT, K=8; 3 products per MAC (hi*hi + hi*lo + lo*hi).
// Accumulation   The tensor core adds into its fp32 accumulator with truncation toward zero (measured:
//                tools/tc_rounding.py), a systematic error growing with the chain length.  So only ONE K-block
//                (<= 54 MMA steps) is accumulated in TMEM; two TMEM accumulator sets ping-pong, and the consumer warps
//                drain each finished partial sum (tcgen05.ld) into fp32 REGISTERS with round-to-nearest adds while the
//                tensor core works on the next K-block / the next tile.  The epilogue (x out_scale) runs from registers.
// Warp roles     w0 TMA(x)  w1 TMA(weights)  w2 MMA issue + TMEM alloc  (w3 idle)  w4..w11 convert + drain + epilogue;
//                setmaxnreg moves registers from the producer warpgroup to the two consumer warpgroups (128 accumulators/thread).
#define GG_TU_TAG 1
#include "tc_common.cuh"
#include <stdlib.h>
#include <limits.h>

using namespace ggtc;

namespace {

// ------------------------------------------------------------------------------------------------ weight packing
constexpr int KB_CH = 16;      // input channels per K-block (two UMMA K=8 steps)
constexpr int MAX_KB = 128;    // K-blocks per conv (I <= 2048); zero-block masks live in 4 registers per lane

struct PackP {
    const float* w; float* wp; uint32_t* mask;
    int O, I, K, NT, n_tiles, num_kb, flip, w_is_IO;
};

// wp[n_tile][kb][tap][half: hi,lo][chunk 0..3][n 0..NT-1][4 channels]  -- per (n_tile,kb,tap) exactly the smem image
// mask[n_tile][kb] bit tap = 1 iff the block holds a non-zero weight
// One thread = one (n, chunk, kb, n_tile) = 4 channels x all taps: for the [O,I,k,k] layout that is one contiguous run of
// 4*k*k floats, for [I,O,k,k] the lanes of a warp (consecutive n) read consecutive k*k-float runs.
__global__ void pack_weights_kernel(PackP p) {
    const int KK = p.K * p.K;
    const int64_t total = (int64_t)p.n_tiles * p.num_kb * 4 * p.NT;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(idx % p.NT); int64_t r = idx / p.NT;
        int chunk = (int)(r % 4); r /= 4;
        int kb = (int)(r % p.num_kb); int nt = (int)(r / p.num_kb);
        const int o = nt * p.NT + n;
        const int i0 = kb * KB_CH + chunk * 4;
        uint32_t live = 0u;
        for (int tap = 0; tap < KK; ++tap) {
            int ky = tap / p.K, kx = tap - ky * p.K;
            if (p.flip) { ky = p.K - 1 - ky; kx = p.K - 1 - kx; }
            float hi[4], lo[4];
            bool nz = false;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int i = i0 + j;
                float v = 0.f;
                if (o < p.O && i < p.I) {
                    int64_t src = p.w_is_IO ? (((int64_t)i * p.O + o) * p.K + ky) * p.K + kx : (((int64_t)o * p.I + i) * p.K + ky) * p.K + kx;
                    v = __ldg(p.w + src);
                }
                nz |= (v != 0.f);
                split_tf32(v, hi[j], lo[j]);
            }
            int64_t blk = (((int64_t)nt * p.num_kb + kb) * KK + tap) * (2 * 4 * p.NT * 4);
            int64_t off = ((int64_t)chunk * p.NT + n) * 4;
            *reinterpret_cast<float4*>(p.wp + blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<float4*>(p.wp + blk + 4 * p.NT * 4 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
            // NT is a multiple of 32: the 32 lanes of a warp always share (nt, kb, chunk)
            if (__ballot_sync(0xffffffffu, nz) != 0u) live |= 1u << tap;
        }
        if (live != 0u && (threadIdx.x & 31) == 0) atomicOr(p.mask + (int64_t)nt * p.num_kb + kb, live);
    }
}

// ------------------------------------------------------------------------------------------------ the conv kernel
constexpr int TILE_H = 16;                // a tile is SUBS 8x16-pixel UMMA M-tiles side by side: 16x16 (SUBS = 2) or 8x16 pixels
constexpr int MAX_W_STAGES = 4;
// Tile configurations <NT, SUBS>: <32|64|128, 2> and <256, 1>.  The operand fetch of an MMA reads 128 A rows whatever N is, and the
// kernel is shared-memory-bandwidth bound, so layers with >= 256 output channels use N = 256 (half the A traffic per FLOP).
__host__ __device__ constexpr int w_stages_for(int NT) { return NT == 256 ? 3 : 4; }
constexpr int CHUNK_TAPS = 3;   // live taps accumulated in TMEM before the partial sum is drained (18 truncating accumulates).
// Longer chunks were tried (GG_CHUNK_TAPS, up to 9 = one drain per K-block: +5..25 % throughput) and rejected: truncation
// removes ~0.5 ulp of the RUNNING sum at every accumulate, i.e. term i of an n-term chain ends up weighted by 1 - beta (n - i):
// after the mean compensation a linear ramp of +-beta n/2 over the chain is left, which acts like a tiny but COHERENT
// perturbation of the filter taps.  A single conv stays at 2e-6, but gradient sums over all pixels (noise strengths, biases,
// weights) add that error coherently against a random-walk-sized signal: the Gmain parity test goes from 4e-6 to 1.3e-3.
constexpr int CONS_WARPS = 8;
constexpr int CONS_THREADS = CONS_WARPS * 32;
constexpr int PROD_WARPS = 4;           // one warpgroup: TMA(x), TMA(weights), MMA, idle -- so that setmaxnreg can shift registers
constexpr int NUM_THREADS = (PROD_WARPS + CONS_WARPS) * 32;
constexpr int MAX_SCALE_CH = MAX_KB * KB_CH;
constexpr int CVT_ITEMS = (4 * 20 * 18 + CONS_THREADS - 1) / CONS_THREADS;   // (chunk, pixel) items per consumer thread, worst case

struct TcP {
    const float* wp; const uint32_t* mask; float* y; const float* in_scale; const float* out_scale;
    int Nimg, I, O, OH, OW, K, pad_y, pad_x;
    int tiles_x, tiles_y, n_tiles, num_kb, nprod, total_tiles;
    int boxW, boxH;            // converted halo tile in pixels; boxW is its smem pixel pitch
    int chunk_taps;            // live taps accumulated in TMEM before the partial sum is drained (<= CHUNK_TAPS)
    int rawW;                  // width of the raw TMA box (>= boxW: the box must start on a 16-byte boundary in global memory)
    ConvEpilogue epi;          // fused bias / noise / activation (act == 0: none)
};

struct SmemLayout {   // byte offsets from the 128-byte aligned dynamic smem base (arithmetic, so that stage indices stay in registers)
    uint32_t raw0, rtile, cvt0, tile, w0, wbytes, scale, bars, tmem_slot, total;
    __host__ __device__ uint32_t raw(int s) const { return raw0 + (uint32_t)s * rtile; }
    __host__ __device__ uint32_t cvt(int s, int h) const { return cvt0 + (uint32_t)(2 * s + h) * tile; }
    __host__ __device__ uint32_t wst(int s) const { return w0 + (uint32_t)s * wbytes; }
};

__host__ __device__ inline SmemLayout make_layout(int boxW, int boxH, int rawW, int NT) {
    const int W_STAGES = w_stages_for(NT);
    SmemLayout L;
    L.tile = ((uint32_t)(KB_CH * boxW * boxH * 4) + 127) & ~127u;
    L.rtile = ((uint32_t)(KB_CH * rawW * boxH * 4) + 127) & ~127u;
    L.wbytes = (uint32_t)(2 * 4 * NT * 16);
    uint32_t off = 0;
    L.raw0 = off; off += 2 * L.rtile;
    L.cvt0 = off; off += 4 * L.tile;
    L.w0 = off; off += W_STAGES * L.wbytes;
    L.scale = off; off += MAX_SCALE_CH * 4;
    L.bars = off; off += 256;
    L.tmem_slot = off; off += 16;
    L.total = off;
    return L;
}

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct TileCoord { int tx, ty, img, nt; };
__device__ __forceinline__ TileCoord decode_tile(int t, const TcP& p) {
    TileCoord c;
    c.tx = t % p.tiles_x; t /= p.tiles_x;
    c.ty = t % p.tiles_y; t /= p.tiles_y;
    c.img = t % p.Nimg;
    c.nt = t / p.Nimg;
    return c;
}

// Zero-block masks of one n-tile, spread over the lanes of a warp: lane l holds kb = l + 32*i in mk[i].
struct KbMasks {
    uint32_t mk[MAX_KB / 32];
    int last_kb;            // highest K-block with any live tap, -1 if the whole n-tile is zero
    __device__ __forceinline__ void load(const TcP& p, int nt, int lane) {
        last_kb = -1;
#pragma unroll
        for (int i = 0; i < MAX_KB / 32; ++i) {
            const int kb = lane + 32 * i;
            uint32_t m = 0u;
            if (kb < p.num_kb) m = p.mask ? __ldg(p.mask + (size_t)nt * p.num_kb + kb) : 0x1FFu;
            mk[i] = m;
            const unsigned b = __ballot_sync(0xffffffffu, m != 0u);
            if (b) last_kb = 32 * i + 31 - __clz(b);
        }
    }
    __device__ __forceinline__ uint32_t get(int kb) const {
        uint32_t v = 0;
#pragma unroll
        for (int i = 0; i < MAX_KB / 32; ++i) {
            const uint32_t t = __shfl_sync(0xffffffffu, mk[i], kb & 31);
            if ((kb >> 5) == i) v = t;
        }
        return v;
    }
};

template <int NT, int SUBS, bool EPI>
__global__ void __launch_bounds__(NUM_THREADS, 1) conv_tc_kernel(const __grid_constant__ CUtensorMap xmap, TcP p) {
    constexpr int W_STAGES = w_stages_for(NT);
    constexpr int TILE_W = 8 * SUBS;
    extern __shared__ __align__(128) uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 127u) & ~127u;
    uint8_t* gbase = smem_raw + (base - smem_u32(smem_raw));
    const SmemLayout L = make_layout(p.boxW, p.boxH, p.rawW, NT);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t bar0 = base + L.bars;
    auto BAR_RAW_FULL = [&](int s) { return bar0 + 8u * s; };
    auto BAR_RAW_EMPTY = [&](int s) { return bar0 + 8u * (2 + s); };
    auto BAR_CVT_FULL = [&](int s) { return bar0 + 8u * (4 + s); };
    auto BAR_CVT_EMPTY = [&](int s) { return bar0 + 8u * (6 + s); };
    auto BAR_ACC_FULL = [&](int s) { return bar0 + 8u * (8 + s); };
    auto BAR_ACC_EMPTY = [&](int s) { return bar0 + 8u * (10 + s); };
    auto BAR_W_FULL = [&](int s) { return bar0 + 8u * (12 + s); };
    auto BAR_W_EMPTY = [&](int s) { return bar0 + 8u * (12 + MAX_W_STAGES + s); };

    const int KK = p.K * p.K;
    const int npix = p.boxW * p.boxH;
    const int rpix = p.rawW * p.boxH;
    const uint32_t raw_bytes = (uint32_t)(KB_CH * rpix * 4);
    constexpr uint32_t w_bytes = (uint32_t)(2 * 4 * NT * 16);
    constexpr uint32_t SET_COLS = SUBS * NT;    // TMEM columns of one accumulator set
    constexpr uint32_t TMEM_COLS = 2 * SET_COLS; // two sets ping-pong

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(BAR_RAW_FULL(s), 1); mbar_init(BAR_RAW_EMPTY(s), CONS_WARPS);
            mbar_init(BAR_CVT_FULL(s), CONS_WARPS); mbar_init(BAR_CVT_EMPTY(s), SUBS);    // one commit per MMA issuer (one per sub-tile)
            mbar_init(BAR_ACC_FULL(s), SUBS); mbar_init(BAR_ACC_EMPTY(s), CONS_WARPS);
        }
        for (int s = 0; s < W_STAGES; ++s) { mbar_init(BAR_W_FULL(s), 1); mbar_init(BAR_W_EMPTY(s), SUBS); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(base + L.tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + L.tmem_slot);

    if (warp < PROD_WARPS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
    if (warp == 0) {
        // ===== x producer (one thread): one 4-D TMA box per live K-block.  Boxes that stick out on the right / bottom are
        // zero-filled by the TMA unit (= the conv padding).  The innermost start coordinate must land on a 16-byte boundary
        // (an unaligned or negative one faults with "illegal instruction"), so the box starts at the aligned, clamped
        // (cx, cy) and the converter shifts it back and re-creates the left / top padding (see cvt_src below).
        if (elect_one()) {
            uint32_t kbc = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, p);
                const uint32_t* mrow = p.mask + (size_t)tc.nt * p.num_kb;
                const int cx = max((tc.tx * TILE_W - p.pad_x) & ~3, 0), cy = max(tc.ty * TILE_H - p.pad_y, 0);
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    if (__ldg(mrow + kb) == 0u) continue;
                    const int s = kbc & 1;
                    mbar_wait(BAR_RAW_EMPTY(s), ((kbc >> 1) & 1) ^ 1);
                    mbar_expect_tx(BAR_RAW_FULL(s), raw_bytes);
                    tma_load_4d(base + L.raw(s), &xmap, BAR_RAW_FULL(s), cx, cy, kb * KB_CH, tc.img);
                    ++kbc;
                }
            }
        }
    } else if (warp == 1) {
        // ===== weight producer (one thread): one bulk copy (hi+lo image of one live tap) per ring stage
        if (elect_one()) {
            uint32_t ws = 0, wph = 0;
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, p);
                const uint32_t* mrow = p.mask + (size_t)tc.nt * p.num_kb;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wp) + (size_t)tc.nt * p.num_kb * KK * w_bytes;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    const uint32_t m = __ldg(mrow + kb);
                    for (int tap = 0; tap < KK; ++tap) {
                        if (!((m >> tap) & 1u)) continue;
                        mbar_wait(BAR_W_EMPTY(ws), wph ^ 1);
                        mbar_expect_tx(BAR_W_FULL(ws), w_bytes);
                        bulk_load(base + L.wst(ws), src + (size_t)(kb * KK + tap) * w_bytes, w_bytes, BAR_W_FULL(ws));
                        if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 2 || (SUBS == 2 && warp == 3)) {
        // ===== MMA issuers: ONE THREAD PER SUB-TILE (warp 2: sub-tile 0, warp 3: sub-tile 1).  The issue loop is scalar code of a
        // single thread -- ~100 dependent instructions per filter tap for 12 MMAs, ~7 cycles each (ncu source view) -- and for
        // N <= 128 that, not the tensor core, set the pace.  The two sub-tiles have independent accumulators, so two threads on
        // two different schedulers issue them side by side; every "empty" / "accumulator full" barrier counts one
        // tcgen05.commit per issuer (a commit covers the MMAs of the thread that executes it).
        if (elect_one()) {
            const uint32_t sub = (uint32_t)(warp - 2);
            uint32_t kbc = 0, ws = 0, wph = 0, ac = 0;          // K-block, weight-stage and accumulator-chunk counters
            const uint32_t idesc = umma_idesc_tf32(128, NT, 0, 0);
            // descriptor: bits [0,14) start>>4, [16,30) LBO>>4, [32,46) SBO>>4, bit 46 = version 1
            //   A: LBO = npix*16 B (between 4-channel chunks), SBO = boxW*16 B (next pixel row = next 8-row group)
            //   B: LBO = NT*16 B, SBO = 128 B
            const uint64_t a_word = ((uint64_t)(((uint32_t)p.boxW & 0x3FFFu) | (1u << 14)) << 32) | (((uint32_t)npix & 0x3FFFu) << 16);
            const uint64_t b_word = ((uint64_t)(8u | (1u << 14)) << 32) | (((uint32_t)NT & 0x3FFFu) << 16);
            const uint32_t a_ks = 2u * (uint32_t)npix;          // second K=8 step: two 4-channel chunks further (16-byte units)
            constexpr uint32_t b_ks = 2u * NT, b_lo_off = 4u * NT;
            const bool three = p.nprod == 3;
            int live_nt = -1, live_total = 0;                   // live (K-block, tap) pairs of the current n-tile
            for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
                const TileCoord tc = decode_tile(t, p);
                const uint32_t* mrow = p.mask + (size_t)tc.nt * p.num_kb;
                if (tc.nt != live_nt) {
                    live_total = 0;
                    for (int kb = 0; kb < p.num_kb; ++kb) live_total += __popc(__ldg(mrow + kb));
                    live_nt = tc.nt;
                }
                uint32_t m_next = __ldg(mrow);
                // A chunk = up to chunk_taps live taps accumulated in one TMEM set.  It may SPAN K-blocks (the phase-major stride-2
                // weights have K-blocks with one or two live taps: one drain per K-block made those layers drain-bound); it closes
                // when it is full or with the last live tap of the tile.
                uint32_t d = 0, accf = 0u;
                int in_chunk = 0, left = live_total;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    const uint32_t m = m_next;
                    if (kb + 1 < p.num_kb) m_next = __ldg(mrow + kb + 1);
                    if (m == 0u) continue;
                    const uint32_t cs = kbc & 1, cph = (kbc >> 1) & 1;
                    mbar_wait(BAR_CVT_FULL(cs), cph);
                    tc_fence_after();
                    // sub-tile `sub` starts 8 pixels (8 x 16 bytes) to the right
                    const uint64_t a_hi0 = a_word + ((base + L.cvt(cs, 0)) >> 4) + 8u * sub, a_lo0 = a_word + ((base + L.cvt(cs, 1)) >> 4) + 8u * sub;
                    uint32_t tap = 0;
                    for (int ky = 0; ky < p.K; ++ky) {
                        for (int kx = 0; kx < p.K; ++kx, ++tap) {
                            if (!((m >> tap) & 1u)) continue;
                            if (in_chunk == 0) {                      // open a chunk: TMEM accumulator set ac & 1, drained two chunks ago
                                const uint32_t as = ac & 1;
                                mbar_wait(BAR_ACC_EMPTY(as), ((ac >> 1) & 1) ^ 1);
                                d = tmem_base + as * SET_COLS + sub * NT;
                                accf = 0u;
                            }
                            mbar_wait(BAR_W_FULL(ws), wph);
                            tc_fence_after();
                            const uint32_t toff = (uint32_t)(ky * p.boxW + kx);
                            const uint64_t b_hi0 = b_word + ((base + L.wst(ws)) >> 4), b_lo0 = b_hi0 + b_lo_off;
                            const uint64_t ah0 = a_hi0 + toff, al0 = a_lo0 + toff;
                            umma_tf32(d, ah0, b_hi0, idesc, accf);                       // K step 0
                            if (three) { umma_tf32(d, ah0, b_lo0, idesc, 1u); umma_tf32(d, al0, b_hi0, idesc, 1u); }
                            umma_tf32(d, ah0 + a_ks, b_hi0 + b_ks, idesc, 1u);           // K step 1
                            if (three) { umma_tf32(d, ah0 + a_ks, b_lo0 + b_ks, idesc, 1u); umma_tf32(d, al0 + a_ks, b_hi0 + b_ks, idesc, 1u); }
                            umma_commit(BAR_W_EMPTY(ws));          // frees the weight stage when these MMAs have read it
                            if (++ws == W_STAGES) { ws = 0; wph ^= 1; }
                            accf = 1u;
                            --left;
                            if (++in_chunk == p.chunk_taps || left == 0) {   // close the chunk: its partial sums are complete in TMEM
                                umma_commit(BAR_ACC_FULL(ac & 1));
                                ++ac; in_chunk = 0;
                            }
                        }
                    }
                    umma_commit(BAR_CVT_EMPTY(cs));                // frees the converted tile
                    ++kbc;
                }
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 224;");
        // ===== consumer warps (w4..w11): convert K-block j, then drain the partial sums of K-block j-1 into registers
        // (and store the finished tile) while the tensor core works on K-block j.
        const int cw = warp - PROD_WARPS;               // 0..7
        const int ct = threadIdx.x - PROD_WARPS * 32;   // 0..255
        const int q = warp & 3;                         // TMEM lane quarter this warp may access
        const int hcol = cw >> 2;                       // which half of the NT accumulator columns this warp owns
        constexpr int HN = NT / 2;
        float acc[SUBS][HN];
#pragma unroll
        for (int s = 0; s < SUBS; ++s)
#pragma unroll
            for (int j = 0; j < HN; ++j) acc[s][j] = 0.f;
        float* sc_s = reinterpret_cast<float*>(gbase + L.scale);
        KbMasks M; int cur_nt = -1, cur_img = -1;
        uint32_t kbc = 0;
        bool pend = false, pend_last = false;
        uint32_t ac = 0, pend_ac = 0;                   // accumulator-chunk counter (same sequence as the MMA issuer's)
        int pend_n = 0, pend_size = 0;                  // chunks that closed inside the pending K-block, taps of the last of them
        int open_taps = 0;                              // taps already in the chunk that is still open (chunks span K-blocks)
        TileCoord pend_tc{0, 0, 0, 0};
        const size_t plane = (size_t)p.OH * p.OW;
        int cvt_src[CVT_ITEMS];                         // conversion plan: source index in the raw box per (chunk, pixel) item
        int plan_dx = INT_MIN, plan_dy = INT_MIN;

        const EpilogueScalars epi_s = epilogue_scalars(p.epi);
        float tile_nz[SUBS], next_nz[SUBS];             // EPI: the noise value of this thread's output pixel(s), current / upcoming tile
#pragma unroll
        for (int sub = 0; sub < SUBS; ++sub) { tile_nz[sub] = 0.f; next_nz[sub] = 0.f; }
        auto load_nz = [&](const TileCoord& tc, float (&dst)[SUBS]) {
            if (!EPI || p.epi.noise == nullptr) return;
            const int m = q * 32 + lane;
            const int oy = tc.ty * TILE_H + (m >> 3);
#pragma unroll
            for (int sub = 0; sub < SUBS; ++sub) {
                const int ox = tc.tx * TILE_W + 8 * sub + (m & 7);
                dst[sub] = (oy < p.OH && ox < p.OW) ? __ldg(p.epi.noise + (size_t)tc.img * p.epi.noise_bs + (size_t)oy * p.OW + ox) : 0.f;
            }
        };
        auto store_tile = [&](const TileCoord& tc) {
            const int m = q * 32 + lane;                // accumulator row = pixel inside the 8x16 sub-tile
            const int r = m >> 3, c = m & 7;
            const int oy = tc.ty * TILE_H + r;
            const int n0 = tc.nt * NT + hcol * HN;
#pragma unroll
            for (int sub = 0; sub < SUBS; ++sub) {
                const int ox = tc.tx * TILE_W + 8 * sub + c;
                const bool pix_ok = (oy < p.OH) && (ox < p.OW);
                float* yp = p.y + ((size_t)tc.img * p.O + n0) * plane + (size_t)oy * p.OW + ox;
                const float* os = p.out_scale ? p.out_scale + (size_t)tc.img * p.O + n0 : nullptr;
                const int nvalid = pix_ok ? min(HN, p.O - n0) : 0;
                const float nz = EPI ? tile_nz[sub] : 0.f;     // prefetched when the tile was opened (a global load in here would sit on the
                                                                // consumers' critical path)
#pragma unroll
                for (int j = 0; j < HN; ++j) {
                    if (j < nvalid) {
                        float val = acc[sub][j];
                        if (os) val *= __ldg(os + j);
                        if (EPI) val = epilogue_apply(val, (p.epi.bias ? __ldg(p.epi.bias + n0 + j) : 0.f) + nz, epi_s);
                        *yp = val;
                    }
                    yp += plane;                        // running pointer: one live address instead of HN
                    acc[sub][j] = 0.f;
                }
            }
        };
        auto drain_chunk = [&](uint32_t k, float kc) {
            const int s = k & 1;
            mbar_wait(BAR_ACC_FULL(s), (k >> 1) & 1);
            tc_fence_after();
            // TMEM loads are issued in pairs before one wait, so that their latencies overlap (one load + wait at a time made the
            // drain a chain of ~16 dependent round trips at NT = 256)
            constexpr int NLD = SUBS * HN / 16;               // 16-column loads of this thread: (sub, cb) in row-major order
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * SET_COLS + hcol * HN);
#pragma unroll
            for (int l = 0; l < NLD; l += 2) {
                uint32_t v[2][16];
#pragma unroll
                for (int u = 0; u < 2; ++u)
                    if (l + u < NLD) tmem_ld16_nowait(t0 + (uint32_t)(((l + u) / (HN / 16)) * NT + ((l + u) % (HN / 16)) * 16), v[u]);
                tmem_wait_ld();
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    if (l + u < NLD) {
                        const int sub = (l + u) / (HN / 16), cb = ((l + u) % (HN / 16)) * 16;
#pragma unroll
                        for (int j = 0; j < 16; ++j) acc[sub][cb + j] = fmaf(__uint_as_float(v[u][j]), kc, acc[sub][cb + j]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(BAR_ACC_EMPTY(s));
        };

        // the chunks that CLOSED inside one K-block: all full (chunk_taps taps) except possibly the last one of a tile
        auto drain = [&](uint32_t first_chunk, int nclosed, int last_size) {
            for (int c = 0; c < nclosed; ++c)
                drain_chunk(first_chunk + c, rz_compensation(2 * (c == nclosed - 1 ? last_size : p.chunk_taps), p.nprod));
        };

        for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
            const TileCoord tc = decode_tile(t, p);
            if (tc.nt != cur_nt) { M.load(p, tc.nt, lane); cur_nt = tc.nt; }
            if (M.last_kb < 0) {                        // the whole n-tile of weights is zero: the output tile is zero
                if (pend) { drain(pend_ac, pend_n, pend_size); if (pend_last) store_tile(pend_tc); pend = false; }
                load_nz(tc, tile_nz);
                store_tile(tc);
                continue;
            }
            {   // EPI: fetch this tile's noise now; if the previous tile is still waiting for its store, keep its values until then
                if (pend && pend_last) load_nz(tc, next_nz); else load_nz(tc, tile_nz);
            }
            if (p.in_scale && tc.img != cur_img) {      // stage this image's per-channel input scales (styles)
                named_bar_sync(1, CONS_THREADS);
                for (int i = ct; i < p.num_kb * KB_CH; i += CONS_THREADS)
                    sc_s[i] = (i < p.I) ? __ldg(p.in_scale + (size_t)tc.img * p.I + i) : 0.f;
                named_bar_sync(1, CONS_THREADS);
                cur_img = tc.img;
            }
            // source index inside the raw box for each converted (chunk, pixel) item this thread owns (-1 = padding -> 0)
            const int ox0 = tc.tx * TILE_W, oy0 = tc.ty * TILE_H;
            const int cx = max((ox0 - p.pad_x) & ~3, 0), cy = max(oy0 - p.pad_y, 0);
            const int dx = (ox0 - p.pad_x) - cx, dy = (oy0 - p.pad_y) - cy;       // converted (r,c) <- raw (r+dy, c+dx)
            // (the plan only depends on how the box had to be shifted / clamped, i.e. on (dx, dy): interior tiles share one plan,
            //  and its ~60 integer divisions per thread are not repeated for every tile)
            if (dx != plan_dx || dy != plan_dy) {
                plan_dx = dx; plan_dy = dy;
#pragma unroll
                for (int it = 0; it < CVT_ITEMS; ++it) {
                    const int idx = ct + it * CONS_THREADS;
                    const int chunk = idx / npix, px = idx - chunk * npix;
                    const int rr = px / p.boxW + dy, cc = px % p.boxW + dx;
                    // bits 0..23: source index in the raw box, bits 24..25: the 4-channel chunk (so that the hot loop has no division)
                    cvt_src[it] = (chunk < 4 && rr >= 0 && cc >= 0 && rr < p.boxH && cc < p.rawW) ? (((chunk * 4) * rpix + rr * p.rawW + cc) | (chunk << 24))
                                                                                                   : (chunk < 4 ? -1 : -2);
                }
            }
            for (int kb = 0; kb < p.num_kb; ++kb) {
                const uint32_t live = M.get(kb);
                if (live == 0u) continue;
                const int s = kbc & 1;
                mbar_wait(BAR_RAW_FULL(s), (kbc >> 1) & 1);
                mbar_wait(BAR_CVT_EMPTY(s), ((kbc >> 1) & 1) ^ 1);
                const float* raw = reinterpret_cast<const float*>(gbase + L.raw(s));
                float4* hi = reinterpret_cast<float4*>(gbase + L.cvt(s, 0));
                float4* lo = reinterpret_cast<float4*>(gbase + L.cvt(s, 1));
                // K-block j is converted and published FIRST, then the partial sums of K-block j-1 are drained: the tensor core goes
                // from the MMAs of j-1 straight to those of j (other TMEM set) while the drain of j-1 runs next to them.
#pragma unroll
                for (int it = 0; it < CVT_ITEMS; ++it) {
                    const int plan = cvt_src[it], src = plan < 0 ? plan : (plan & 0xFFFFFF);
                    if (src != -2) {                    // -2: past the end of the tile
                        const int idx = ct + it * CONS_THREADS;
                        float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
                        if (src >= 0) {
                            float v0 = raw[src], v1 = raw[src + rpix], v2 = raw[src + 2 * rpix], v3 = raw[src + 3 * rpix];
                            if (p.in_scale) {
                                const int chunk = plan >> 24;
                                const float4 sv = *reinterpret_cast<const float4*>(sc_s + kb * KB_CH + chunk * 4);
                                v0 *= sv.x; v1 *= sv.y; v2 *= sv.z; v3 *= sv.w;
                            }
                            split_tf32(v0, h.x, l.x); split_tf32(v1, h.y, l.y);
                            split_tf32(v2, h.z, l.z); split_tf32(v3, h.w, l.w);
                        }
                        hi[idx] = h;
                        lo[idx] = l;
                    }
                }
                fence_proxy_async();                    // generic-proxy writes -> visible to the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) { mbar_arrive(BAR_CVT_FULL(s)); mbar_arrive(BAR_RAW_EMPTY(s)); }
                if (pend) {
                    drain(pend_ac, pend_n, pend_size);
                    if (pend_last) {
                        store_tile(pend_tc);
                        if (EPI && (pend_tc.tx != tc.tx || pend_tc.ty != tc.ty || pend_tc.img != tc.img || pend_tc.nt != tc.nt)) {
#pragma unroll
                            for (int sub = 0; sub < SUBS; ++sub) tile_nz[sub] = next_nz[sub];
                        }
                    }
                }
                pend = true; pend_last = (kb == M.last_kb); pend_tc = tc;
                pend_ac = ac;
                {   // same chunk sequence as the issuer: full chunks close as they fill up, the remainder closes with the tile
                    const int total = open_taps + __popc(live);
                    pend_n = pend_last ? (total + p.chunk_taps - 1) / p.chunk_taps : total / p.chunk_taps;
                    const int rem = total % p.chunk_taps;
                    pend_size = (pend_last && rem != 0) ? rem : p.chunk_taps;
                    open_taps = pend_last ? 0 : rem;
                }
                ac += pend_n;
                ++kbc;
            }
        }
        if (pend) { drain(pend_ac, pend_n, pend_size); if (pend_last) store_tile(pend_tc); }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <int NT, int SUBS, bool EPI>
int launch_conv_tc(const CUtensorMap& xmap, const TcP& p, cudaStream_t st) {
    const SmemLayout L = make_layout(p.boxW, p.boxH, p.rawW, NT);
    const size_t smem = L.total + 128;
    if (smem > 227 * 1024) { gg::set_error("conv2d(tc): shared-memory layout of %zu bytes does not fit", smem); return GG_EUNSUPPORTED; }
    static std::atomic<uint64_t> attr_set{0};           // one bit per device: the opt-in belongs to the device's context
    if (!gg::done_on_this_device(attr_set)) {
        GG_CUDA(cudaFuncSetAttribute(conv_tc_kernel<NT, SUBS, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        gg::mark_done_on_this_device(attr_set);
    }
    const int grid = p.total_tiles < GG_NUM_SMS ? p.total_tiles : GG_NUM_SMS;
    wd_arm();
    conv_tc_kernel<NT, SUBS, EPI><<<grid, NUM_THREADS, smem, st>>>(xmap, p);
    return gg::check_launch("conv2d(tc)");
}

}  // namespace

gg::EncodeTiledFn gg::get_encode_fn() {
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

namespace gg {

// The stream-ordered scratch comes from the device's default memory pool.  Its default release threshold of 0 hands the
// memory back to the driver at every synchronisation, which makes the next cudaMallocAsync cost ~1 ms; keep it cached.
void keep_pool_memory() {
    static std::atomic<uint64_t> done{0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return;
    if (done.load() & (1ull << dev)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done.fetch_or(1ull << dev);
}

bool conv2d_tc_eligible(int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int stride, int pad_y, int pad_x,
                        int transposed) {
    if (stride != 1 || KH != KW || KH < 1 || KH > 3) return false;
    if (pad_y > KH - 1 || pad_x > KW - 1) return false;
    if (N < 1 || I < 16 || O < 16) return false;       // 3-channel fromRGB / ToRGB stay on the FFMA path (HBM-bound, AI ~ 1.4)
    if (I > MAX_KB * KB_CH) return false;
    if (W % 4 != 0) return false;                      // TMA global strides must be multiples of 16 bytes
    // small maps (4x4, 5x8 ...) also run here: one mostly empty 16x16 tile per image and n-tile still beats an FFMA kernel that
    // can only spread N*OH*OW <= 128 output pixels over a handful of CTAs while walking K = 9*512 sequentially.
    (void)H; (void)OH; (void)OW; (void)transposed;
    return true;
}

int conv2d_tc(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int K, int /*KW*/, int OH, int OW, int pad_y,
              int pad_x, int flip_w, int w_is_IO, const float* in_scale, const float* out_scale, int nprod, const ConvEpilogue* epi,
              cudaStream_t st) {
    if ((reinterpret_cast<uintptr_t>(x) & 15) != 0) { set_error("conv2d(tc): x must be 16-byte aligned"); return GG_EINVAL; }
    int NT = O > 128 ? 256 : (O > 64 ? 128 : (O > 32 ? 64 : 32));
    {
        // Small problems (the 4^2 .. 32^2 maps, small batches) do not fill one wave of CTAs with the widest n-tile, and the K loop of a
        // single CTA is then the whole cost of the launch.  Model: cost = waves * (M sub-tiles per CTA) * max(NT, 64) -- an MMA takes
        // ~NT cycles but never less than the 64 it needs to fetch its 128 A rows -- and take the cheapest n-tile (ties: the widest).
        auto cost = [&](int nt) {
            const int tw = nt == 256 ? 8 : 16;
            const int64_t ctas = (int64_t)((OW + tw - 1) / tw) * ((OH + TILE_H - 1) / TILE_H) * N * ((O + nt - 1) / nt);
            const int64_t waves = (ctas + GG_NUM_SMS - 1) / GG_NUM_SMS;
            return waves * (nt == 256 ? 1 : 2) * (nt > 64 ? nt : 64);
        };
        int best = NT;
        for (int nt = NT / 2; nt >= 32; nt /= 2)
            if (cost(nt) < cost(best)) best = nt;
        if (cost(NT) <= 4 * 256) NT = best;      // only launches of a few waves: large ones stay on the widest tile (weights fetched once)
    }
    const int TILE_W = NT == 256 ? 8 : 16;
    const int n_tiles = (O + NT - 1) / NT;
    const int num_kb = (I + KB_CH - 1) / KB_CH;
    const int KK = K * K;
    const int boxW = ((TILE_W + K - 1) + 3) / 4 * 4, boxH = TILE_H + K - 1;
    const int rawW = (TILE_W + K - 1 + ((4 - pad_x % 4) % 4) + 3) / 4 * 4;   // aligned-down start => up to 3 extra columns

    // 1. pack + split the weights, flag the all-zero blocks (tiny; stream-ordered scratch)
    keep_pool_memory();
    const size_t wp_floats = (size_t)n_tiles * num_kb * KK * 2 * 4 * NT * 4;
    const size_t mask_words = (size_t)n_tiles * num_kb;
    float* wp = nullptr;
    GG_CUDA(cudaMallocAsync(&wp, wp_floats * sizeof(float) + mask_words * sizeof(uint32_t), st));
    uint32_t* mask = reinterpret_cast<uint32_t*>(wp + wp_floats);
    cudaMemsetAsync(mask, 0, mask_words * sizeof(uint32_t), st);
    PackP pp{w, wp, mask, O, I, K, NT, n_tiles, num_kb, flip_w, w_is_IO};
    {
        int64_t threads = (int64_t)n_tiles * num_kb * 4 * NT;
        int grid = (int)((threads + 255) / 256);
        if (grid > GG_NUM_SMS * 8) grid = GG_NUM_SMS * 8;
        pack_weights_kernel<<<grid, 256, 0, st>>>(pp);
        int rc = check_launch("conv2d(tc) pack_weights");
        if (rc != GG_OK) { cudaFreeAsync(wp, st); return rc; }
    }

    // 2. TMA descriptor of x as a 4-D tensor {W, H, C, N}; box {rawW, boxH, 16, 1}; OOB elements read as zero
    gg::EncodeTiledFn encode = gg::get_encode_fn();
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

    // 3. launch (persistent: one CTA per SM walks the tiles, n-tile slowest so that concurrent CTAs share weights in L2)
    TcP p{};
    p.wp = wp; p.mask = mask; p.y = y; p.in_scale = in_scale; p.out_scale = out_scale;
    p.Nimg = N; p.I = I; p.O = O; p.OH = OH; p.OW = OW; p.K = K; p.pad_y = pad_y; p.pad_x = pad_x;
    p.tiles_x = (OW + TILE_W - 1) / TILE_W; p.tiles_y = (OH + TILE_H - 1) / TILE_H; p.n_tiles = n_tiles; p.num_kb = num_kb;
    p.nprod = (nprod == GG_PREC_TF32X1) ? 1 : 3;
    const int64_t total = (int64_t)p.tiles_x * p.tiles_y * N * n_tiles;
    if (total > 0x7fffffffLL) { cudaFreeAsync(wp, st); set_error("conv2d(tc): too many tiles"); return GG_EINVAL; }
    p.total_tiles = (int)total;
    p.boxW = boxW; p.boxH = boxH; p.rawW = rawW;
    p.chunk_taps = CHUNK_TAPS;
    if (epi) p.epi = *epi;             // (value-initialised otherwise: act == 0)
    int rc;
    const bool e = p.epi.act != 0;      // the fused-epilogue instantiations are separate kernels: the plain ones stay as they were
    if (NT == 256) rc = e ? launch_conv_tc<256, 1, true>(xmap, p, st) : launch_conv_tc<256, 1, false>(xmap, p, st);
    else if (NT == 128) rc = e ? launch_conv_tc<128, 2, true>(xmap, p, st) : launch_conv_tc<128, 2, false>(xmap, p, st);
    else if (NT == 64) rc = e ? launch_conv_tc<64, 2, true>(xmap, p, st) : launch_conv_tc<64, 2, false>(xmap, p, st);
    else rc = e ? launch_conv_tc<32, 2, true>(xmap, p, st) : launch_conv_tc<32, 2, false>(xmap, p, st);
    cudaFreeAsync(wp, st);
    return rc;
}

}  // namespace gg
