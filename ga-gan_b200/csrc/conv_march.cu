// Row-marching tcgen05 convolution for layers with FEW output channels (O <= 64): the filter ROWS ride in the MMA N dimension.
//
// conv_tc.cu maps output channels to N.  At 32 or 64 channels an M=128 x N=32/64 x K=8 instruction does 16 / 32 cycles of math
// but its operand fetch reads 128 A rows of 32 bytes whatever N is (>= 32 cycles, ~58 from the 16-byte-shifted tap views): the
// 32-channel layers of the 1024^2 networks ran at 80 TFLOP/s, a third of what the >= 128-channel layers reach.  Here
//
//   M = 128 consecutive pixels of ONE image row (a "segment"),   N = (ky, o) = K * NT accumulator columns,   K = input channels
//   D_r[x, (ky, o)] = sum_{kx, i}  in[r, x - pad_x + kx, i] * v[o, i, ky, kx]                      (input row r)
//   out[y, x, o]    = sum_ky D_{y - pad_y + ky}[x, (ky, o)]
//
// so one instruction serves all K filter rows (N = 96 / 192: the tensor pipe, not the operand fetch, sets the pace), every input
// row is loaded and converted exactly ONCE per segment (no halo rows: the kernel marches down the image and the K partial rows
// that are still open live in the consumer threads' REGISTERS -- pixel x is TMEM lane x is thread x, so the sum over ky never
// crosses a thread), and the kx shift is a descriptor start address as in conv_tc.cu.
//
// A operand      one TMA box per (row, 32-channel K-block): [32 ch][136 px] raw fp32, hardware zero fill right of / below the
//                image; converted once (x in_scale, tf32 hi + lo) into the SWIZZLE_128B K-major layout: ONE 128-byte row per pixel
//                (32 channels), 16-byte chunk c of pixel j stored at chunk position c ^ (j & 7).  A pixel shift is a 128-byte
//                shift: every tap view starts on a full shared-memory line (the 16-byte pixel pitch of conv_tc.cu's no-swizzle
//                layout made the kx != 0 views straddle lines, ~1.8x the operand-fetch cycles).
// B operand      weights packed per (16-channel group, kx) as the no-swizzle K-major image [hi|lo][4-ch chunk][(ky,o)][4 ch],
//                streamed through a 4-stage bulk-TMA ring.
// Accumulation   as conv_tc.cu: a chunk = 3 kx taps x 16 channels (18 truncating accumulates) is summed in TMEM, two TMEM sets
//                ping-pong, the consumer warps drain each chunk into fp32 registers (round to nearest, expected-value
//                compensation of the truncation, tc_common.cuh) while the tensor core works on the next chunk.
// Warp roles     w0 TMA(x)  w1 TMA(weights)  w2 MMA issue + TMEM alloc  (w3 idle)  w4..w11 convert + drain + store.
#define GG_TU_TAG 2
#include "tc_common.cuh"
#include <limits.h>

using namespace ggtc;

namespace {

constexpr int MK = 32;                    // input channels per K-block = one 128-byte operand row per pixel
constexpr int SEG = 128;                  // pixels per segment = UMMA M
constexpr int RAWW = SEG + 8;             // raw box width: up to 3 columns of alignment slack + K-1 halo, multiple of 4
constexpr int CVT_PIX = SEG + 8;          // pixel rows of a converted buffer (whole 8-row swizzle atoms)
constexpr int MW_STAGES = 6;             // ring stages; with <= 6 (16-channel group, kx) blocks the weights stay resident
constexpr int M_PROD_WARPS = 4, M_CONS_WARPS = 8;
constexpr int M_CONS_THREADS = M_CONS_WARPS * 32;
constexpr int M_THREADS = (M_PROD_WARPS + M_CONS_WARPS) * 32;
constexpr int M_MAX_CH = 2048;
constexpr uint32_t RAW_BYTES = MK * RAWW * 4;          // 17408
constexpr uint32_t CVT_BYTES = CVT_PIX * 128;          // 17408 = 17 atoms of 1024 bytes

struct MarchP {
    const float* wp; float* y; const float* in_scale; const float* out_scale;
    int Nimg, I, O, H, W, OH, OW, pad_y, pad_x;
    int num_kb;            // 32-channel K-blocks
    int num_g16;           // 16-channel groups (= weight stages per kx)
    int strips_x, bands, band_rows, total_units, nprod;
    int resident;          // all num_g16*K weight stages fit the ring: loaded once per CTA, never recycled
    ConvEpilogue epi;      // fused bias / noise / activation (act == 0: none)
};

__host__ __device__ constexpr int march_stages(int N) { return N > 128 ? 4 : MW_STAGES; }

struct MLayout {
    uint32_t raw0, cvt0, w0, wbytes, scale, bars, tmem_slot, total;
    __host__ __device__ uint32_t raw(int s) const { return raw0 + (uint32_t)s * RAW_BYTES; }
    __host__ __device__ uint32_t cvt(int s, int h) const { return cvt0 + (uint32_t)(2 * s + h) * CVT_BYTES; }
    __host__ __device__ uint32_t wst(int s) const { return w0 + (uint32_t)s * wbytes; }
};

__host__ __device__ inline MLayout make_mlayout(int N) {      // offsets from the 1024-byte aligned base
    MLayout L;
    uint32_t off = 0;
    L.cvt0 = off; off += 4 * CVT_BYTES;                        // first: swizzled buffers need 1024-byte alignment
    L.raw0 = off; off += 2 * RAW_BYTES;
    L.wbytes = (uint32_t)(2 * 4 * N * 16);
    L.w0 = off; off += (uint32_t)march_stages(N) * L.wbytes;   // 4 x 24 KB (N = 192) or 6 x 12 KB (N = 96)
    L.scale = off; off += M_MAX_CH * 4 + 256;                  // in_scale of the image's channels, then the fused bias of <= 64 outputs
    L.bars = off; off += 256;
    L.tmem_slot = off; off += 16;
    L.total = off;
    return L;
}

// ------------------------------------------------------------------------------------------------ weight packing
struct MPackP { const float* w; float* wp; int O, I, K, NT, num_g16, flip, w_is_IO; };

// wp[g16][kx][half][chunk 0..3][n = ky*NT + o][4 channels]; v[o,i,ky,kx] = w read as [O,I,K,K] or [I,O,K,K], flipped if asked
__global__ void march_pack_weights(MPackP p) {
    const int N = p.K * p.NT;
    const int64_t total = (int64_t)p.num_g16 * p.K * 4 * N;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int n = (int)(idx % N); int64_t r = idx / N;
        const int chunk = (int)(r % 4); r /= 4;
        const int kx = (int)(r % p.K); const int g = (int)(r / p.K);
        const int ky = n / p.NT, o = n - ky * p.NT;
        int sy = ky, sx = kx;
        if (p.flip) { sy = p.K - 1 - ky; sx = p.K - 1 - kx; }
        float hi[4], lo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int i = g * 16 + chunk * 4 + j;
            float v = 0.f;
            if (o < p.O && i < p.I) {
                const int64_t src = p.w_is_IO ? (((int64_t)i * p.O + o) * p.K + sy) * p.K + sx : (((int64_t)o * p.I + i) * p.K + sy) * p.K + sx;
                v = __ldg(p.w + src);
            }
            split_tf32(v, hi[j], lo[j]);
        }
        const int64_t blk = ((int64_t)g * p.K + kx) * (2 * 4 * N * 4);
        const int64_t off = ((int64_t)chunk * N + n) * 4;
        *reinterpret_cast<float4*>(p.wp + blk + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<float4*>(p.wp + blk + 4 * N * 4 + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ------------------------------------------------------------------------------------------------ the kernel
struct Unit { int x0, y0, y1, img; };
__device__ __forceinline__ Unit decode_unit(int u, const MarchP& p) {
    Unit c;
    const int sx = u % p.strips_x; u /= p.strips_x;
    const int b = u % p.bands; c.img = u / p.bands;
    c.x0 = sx * SEG; c.y0 = b * p.band_rows;
    c.y1 = min(c.y0 + p.band_rows, p.OH);
    return c;
}

__device__ __forceinline__ void mnamed_bar_sync(int id, int nthreads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory"); }

template <int NT, int K, bool EPI>
__global__ void __launch_bounds__(M_THREADS, 1) conv_march_kernel(const __grid_constant__ CUtensorMap xmap, MarchP p) {
    constexpr int N = K * NT;                       // accumulator columns of one set: (ky, o)
    constexpr uint32_t TMEM_COLS = (2 * N <= 256) ? 256 : 512;
    constexpr uint32_t w_bytes = (uint32_t)(2 * 4 * N * 16);
    constexpr int STAGES = march_stages(N);
    extern __shared__ __align__(1024) uint8_t msmem_raw[];
    const uint32_t base = (smem_u32(msmem_raw) + 1023u) & ~1023u;
    uint8_t* gbase = msmem_raw + (base - smem_u32(msmem_raw));
    const MLayout L = make_mlayout(N);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    const uint32_t bar0 = base + L.bars;
    auto BAR_RAW_FULL = [&](int s) { return bar0 + 8u * s; };
    auto BAR_RAW_EMPTY = [&](int s) { return bar0 + 8u * (2 + s); };
    auto BAR_CVT_FULL = [&](int s) { return bar0 + 8u * (4 + s); };
    auto BAR_CVT_EMPTY = [&](int s) { return bar0 + 8u * (6 + s); };
    auto BAR_ACC_FULL = [&](int s) { return bar0 + 8u * (8 + s); };
    auto BAR_ACC_EMPTY = [&](int s) { return bar0 + 8u * (10 + s); };
    auto BAR_W_FULL = [&](int s) { return bar0 + 8u * (12 + s); };
    auto BAR_W_EMPTY = [&](int s) { return bar0 + 8u * (12 + MW_STAGES + s); };

    if (threadIdx.x == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(BAR_RAW_FULL(s), 1); mbar_init(BAR_RAW_EMPTY(s), M_CONS_WARPS);
            mbar_init(BAR_CVT_FULL(s), M_CONS_WARPS); mbar_init(BAR_CVT_EMPTY(s), 1);
            mbar_init(BAR_ACC_FULL(s), 1); mbar_init(BAR_ACC_EMPTY(s), M_CONS_WARPS);
        }
        for (int s = 0; s < MW_STAGES; ++s) { mbar_init(BAR_W_FULL(s), 1); mbar_init(BAR_W_EMPTY(s), 1); }
        fence_barrier_init();
    }
    if (warp == 2) tmem_alloc(base + L.tmem_slot, TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gbase + L.tmem_slot);

    // Every role walks the same sequence: units -> t = y0 .. y1+K-2 (t = input row + pad_y) -> K-blocks of a row inside the image.
    if (warp < M_PROD_WARPS) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 56;");
        if (warp == 0) {
            if (elect_one()) {
                uint32_t sc = 0;                                         // step counter
                for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                    const Unit un = decode_unit(u, p);
                    const int cx = max((un.x0 - p.pad_x) & ~3, 0);
                    for (int t = un.y0; t < un.y1 + K - 1; ++t) {
                        const int r = t - p.pad_y;
                        if (r < 0 || r >= p.H) continue;
                        for (int kb = 0; kb < p.num_kb; ++kb) {
                            const int s = sc & 1;
                            mbar_wait(BAR_RAW_EMPTY(s), ((sc >> 1) & 1) ^ 1);
                            mbar_expect_tx(BAR_RAW_FULL(s), RAW_BYTES);
                            tma_load_4d(base + L.raw(s), &xmap, BAR_RAW_FULL(s), cx, r, kb * MK, un.img);
                            ++sc;
                        }
                    }
                }
            }
        } else if (warp == 1) {
            if (elect_one()) {
                uint32_t ws = 0, wph = 0;
                const uint8_t* src = reinterpret_cast<const uint8_t*>(p.wp);
                if (p.resident) {                                         // the whole weight image stays in shared memory
                    for (int st = 0; st < p.num_g16 * K; ++st) {
                        mbar_expect_tx(BAR_W_FULL(st), w_bytes);
                        bulk_load(base + L.wst(st), src + (size_t)st * w_bytes, w_bytes, BAR_W_FULL(st));
                    }
                }
                for (int u = blockIdx.x; !p.resident && u < p.total_units; u += gridDim.x) {
                    const Unit un = decode_unit(u, p);
                    for (int t = un.y0; t < un.y1 + K - 1; ++t) {
                        const int r = t - p.pad_y;
                        if (r < 0 || r >= p.H) continue;
                        for (int g = 0; g < p.num_g16; ++g)
                            for (int kx = 0; kx < K; ++kx) {
                                mbar_wait(BAR_W_EMPTY(ws), wph ^ 1);
                                mbar_expect_tx(BAR_W_FULL(ws), w_bytes);
                                bulk_load(base + L.wst(ws), src + (size_t)(g * K + kx) * w_bytes, w_bytes, BAR_W_FULL(ws));
                                if (++ws == STAGES) { ws = 0; wph ^= 1; }
                            }
                    }
                }
            }
        } else if (warp == 2) {
            if (elect_one()) {
                uint32_t sc = 0, ws = 0, wph = 0, ac = 0;
                bool w_ready = false;                                     // resident weights: waited for once
                const uint32_t idesc = umma_idesc_tf32(128, N, 0, 0);
                // A: SWIZZLE_128B K-major.  start>>4 | LBO (unused, 1) << 16 | SBO = 1024 B (8 pixel rows) << 32 | version 1 << 46 |
                //    layout type 2 << 61.  base_offset (bits 49..51) stays 0 although the tap views start 128*kx bytes into a
                //    swizzle atom: measured on B200, the XOR pattern is taken from the ABSOLUTE shared-memory address bits 7..9 (the
                //    buffers are 1024-byte aligned and written with chunk ^ (pixel & 7)); base_offset = (start >> 7) & 7 shifts it twice.
                const uint64_t a_word = ((uint64_t)2 << 61) | ((uint64_t)1 << 46) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 16);
                // B: no swizzle K-major: LBO = N*16 B between 4-channel chunks, SBO = 128 B between 8-row groups
                const uint64_t b_word = ((uint64_t)(8u | (1u << 14)) << 32) | (((uint64_t)N & 0x3FFFu) << 16);
                constexpr uint32_t b_ks = 2u * N, b_lo_off = 4u * N;
                const bool three = p.nprod == 3;
                for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
                    const Unit un = decode_unit(u, p);
                    for (int t = un.y0; t < un.y1 + K - 1; ++t) {
                        const int r = t - p.pad_y;
                        if (r < 0 || r >= p.H) continue;
                        for (int kb = 0; kb < p.num_kb; ++kb) {
                            const uint32_t cs = sc & 1, cph = (sc >> 1) & 1;
                            mbar_wait(BAR_CVT_FULL(cs), cph);
                            tc_fence_after();
                            const uint32_t a_hi_base = base + L.cvt(cs, 0), a_lo_base = base + L.cvt(cs, 1);
                            const int n16 = min(2, p.num_g16 - 2 * kb);
                            for (int c16 = 0; c16 < n16; ++c16) {
                                const uint32_t as = ac & 1;
                                mbar_wait(BAR_ACC_EMPTY(as), ((ac >> 1) & 1) ^ 1);
                                tc_fence_after();
                                const uint32_t d = tmem_base + as * N;
                                uint32_t accf = 0u;
                                for (int kx = 0; kx < K; ++kx) {
                                    if (p.resident) {
                                        ws = (uint32_t)((kb * 2 + c16) * K + kx);
                                        if (!w_ready) mbar_wait(BAR_W_FULL(ws), 0);
                                    } else {
                                        mbar_wait(BAR_W_FULL(ws), wph);
                                    }
                                    tc_fence_after();
                                    const uint64_t b_hi0 = b_word + ((base + L.wst(ws)) >> 4), b_lo0 = b_hi0 + b_lo_off;
#pragma unroll
                                    for (int ks = 0; ks < 2; ++ks) {
                                        const uint32_t aoff = (uint32_t)kx * 128u + (uint32_t)(c16 * 2 + ks) * 32u;
                                        const uint32_t ah = a_hi_base + aoff, al = a_lo_base + aoff;
                                        const uint64_t adh = a_word | (uint64_t)((ah >> 4) & 0x3FFFu) | 0ull;
                                        const uint64_t adl = a_word | (uint64_t)((al >> 4) & 0x3FFFu) | 0ull;
                                        umma_tf32(d, adh, b_hi0 + ks * b_ks, idesc, accf);
                                        if (three) { umma_tf32(d, adh, b_lo0 + ks * b_ks, idesc, 1u); umma_tf32(d, adl, b_hi0 + ks * b_ks, idesc, 1u); }
                                        accf = 1u;
                                    }
                                    if (!p.resident) {
                                        umma_commit(BAR_W_EMPTY(ws));
                                        if (++ws == STAGES) { ws = 0; wph ^= 1; }
                                    }
                                }
                                umma_commit(BAR_ACC_FULL(as));
                                ++ac;
                            }
                            umma_commit(BAR_CVT_EMPTY(cs));
                            ++sc;
                            if (kb == p.num_kb - 1) w_ready = true;        // every stage has been seen once
                        }
                    }
                }
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
        const int cw = warp - M_PROD_WARPS;               // 0..7
        const int ct = threadIdx.x - M_PROD_WARPS * 32;   // 0..255
        const int q = warp & 3;                           // TMEM lane quarter of this warp
        const int hcol = cw >> 2;                         // which half of the NT output channels this warp owns
        constexpr int HN = NT / 2;
        float acc[K][HN];                                 // the K output rows that are still open, slot = (output row - y0) mod K
#pragma unroll
        for (int s = 0; s < K; ++s)
#pragma unroll
            for (int j = 0; j < HN; ++j) acc[s][j] = 0.f;
        float osc[HN];                                    // out_scale of this thread's channels for the current image
#pragma unroll
        for (int j = 0; j < HN; ++j) osc[j] = 1.f;
        int osc_img = -1;
        float* bsc = reinterpret_cast<float*>(gbase + L.scale) + M_MAX_CH + hcol * HN;   // EPI: the bias of this thread's channels (shared memory)
        if (EPI) {
            if (ct < NT) reinterpret_cast<float*>(gbase + L.scale)[M_MAX_CH + ct] = (p.epi.bias && ct < p.O) ? __ldg(p.epi.bias + ct) : 0.f;
            mnamed_bar_sync(1, M_CONS_THREADS);
        }
        const EpilogueScalars epi_s = epilogue_scalars(p.epi);
        auto fetch_nz = [&](const Unit& un, int y) -> float {     // the noise value of this thread's pixel in output row y
            if (!EPI || p.epi.noise == nullptr) return 0.f;
            const int x = un.x0 + q * 32 + lane;
            return (y >= un.y0 && y < un.y1 && x < p.OW) ? __ldg(p.epi.noise + (size_t)un.img * p.epi.noise_bs + (size_t)y * p.OW + x) : 0.f;
        };
        float* sc_s = reinterpret_cast<float*>(gbase + L.scale);
        const size_t plane = (size_t)p.OH * p.OW;
        const float kcomp = rz_compensation(2 * K, p.nprod);
        int cur_img = -1;
        uint32_t sc = 0, ac = 0;

        // one chunk: K slices of HN columns; slice ky of the row with phase PH = (t - y0) mod K belongs to slot (PH - ky) mod K
        auto drain_chunk = [&](uint32_t k, int ph) {
            const int s = k & 1;
            mbar_wait(BAR_ACC_FULL(s), (k >> 1) & 1);
            tc_fence_after();
            const uint32_t t0 = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(s * N + hcol * HN);
#pragma unroll
            for (int ky = 0; ky < K; ++ky) {
                uint32_t v[HN];
#pragma unroll
                for (int c = 0; c < HN; c += 16) tmem_ld16_nowait(t0 + (uint32_t)(ky * NT + c), &v[c]);
                tmem_wait_ld();
                if (ky == K - 1) {                       // everything of this chunk is in registers: hand the TMEM set back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(BAR_ACC_EMPTY(s));
                }
#pragma unroll
                for (int pp = 0; pp < K; ++pp) {
                    if (ph == pp) {
                        const int slot = (pp - ky + K) % K;
#pragma unroll
                        for (int j = 0; j < HN; ++j) acc[slot][j] = fmaf(__uint_as_float(v[j]), kcomp, acc[slot][j]);
                    }
                }
            }
        };
        // the output row completed by the input row with phase ph: slot (ph + 1) mod K; stored, then cleared for row y + K
        auto store_row = [&](const Unit& un, int y, int ph, float row_nz) {
            const int x = un.x0 + q * 32 + lane;
            const int n0 = hcol * HN;
            const bool ok = (y >= un.y0) && (y < un.y1) && (x < p.OW);
            float* yp = p.y + ((size_t)un.img * p.O + n0) * plane + (size_t)y * p.OW + x;
            const int nvalid = ok ? min(HN, p.O - n0) : 0;
            const float nz = EPI ? row_nz : 0.f;          // fetched when this row's input was converted (not here: critical path)
#pragma unroll
            for (int pp = 0; pp < K; ++pp) {
                if (ph == pp) {
                    const int slot = (pp + 1) % K;
#pragma unroll
                    for (int j = 0; j < HN; ++j) {
                        if (j < nvalid) {
                            float val = acc[slot][j] * osc[j];
                            if (EPI) val = epilogue_apply(val, bsc[j] + nz, epi_s);
                            *yp = val;
                        }
                        yp += plane;
                        acc[slot][j] = 0.f;
                    }
                }
            }
        };

        // pending work of the previous step: its chunks are drained AFTER the next step has been converted and published
        bool pend = false, pend_last = false;
        uint32_t pend_ac = 0; int pend_n = 0, pend_ph = 0, pend_y = 0; float pend_nz = 0.f;
        Unit pend_un{0, 0, 0, 0};
        auto flush = [&]() {
            if (!pend) return;
            for (int c = 0; c < pend_n; ++c) drain_chunk(pend_ac + c, pend_ph);
            if (pend_last) store_row(pend_un, pend_y, pend_ph, pend_nz);
            pend = false;
        };

        for (int u = blockIdx.x; u < p.total_units; u += gridDim.x) {
            const Unit un = decode_unit(u, p);
            flush();                                                   // a band ends with partial sums of rows beyond it in the slots
#pragma unroll
            for (int s = 0; s < K; ++s)
#pragma unroll
                for (int j = 0; j < HN; ++j) acc[s][j] = 0.f;
            if (p.out_scale && un.img != osc_img) {
#pragma unroll
                for (int j = 0; j < HN; ++j) osc[j] = (hcol * HN + j < p.O) ? __ldg(p.out_scale + (size_t)un.img * p.O + hcol * HN + j) : 0.f;
                osc_img = un.img;
            }
            if (p.in_scale && un.img != cur_img) {
                mnamed_bar_sync(1, M_CONS_THREADS);
                for (int i = ct; i < p.num_kb * MK; i += M_CONS_THREADS)
                    sc_s[i] = (i < p.I) ? __ldg(p.in_scale + (size_t)un.img * p.I + i) : 0.f;
                mnamed_bar_sync(1, M_CONS_THREADS);
                cur_img = un.img;
            }
            const int cx = max((un.x0 - p.pad_x) & ~3, 0);
            const int shift = (un.x0 - p.pad_x) - cx;                  // converted pixel j <- raw column j + shift (may be negative at x0 = 0)
            int ph = 0;
            for (int t = un.y0; t < un.y1 + K - 1; ++t, ph = (ph + 1 == K) ? 0 : ph + 1) {
                const int r = t - p.pad_y;
                const int y = t - (K - 1);                             // the output row this input row completes
                if (r < 0 || r >= p.H) {                               // a row of zero padding: nothing to add, but row y is complete
                    flush();
                    store_row(un, y, ph, fetch_nz(un, y));
                    continue;
                }
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    const int s = sc & 1;
                    mbar_wait(BAR_RAW_FULL(s), (sc >> 1) & 1);
                    mbar_wait(BAR_CVT_EMPTY(s), ((sc >> 1) & 1) ^ 1);
                    const float* raw = reinterpret_cast<const float*>(gbase + L.raw(s));
                    uint8_t* hi = gbase + L.cvt(s, 0);
                    uint8_t* lo = gbase + L.cvt(s, 1);
                    // items: (4-channel chunk c = 0..7, pixel j = 0..CVT_PIX-1); lanes <-> consecutive pixels
#pragma unroll
                    for (int it = 0; it < (8 * CVT_PIX + M_CONS_THREADS - 1) / M_CONS_THREADS; ++it) {
                        const int idx = ct + it * M_CONS_THREADS;
                        if (idx < 8 * CVT_PIX) {
                            const int c = idx / CVT_PIX, j = idx - c * CVT_PIX;
                            const int src = j + shift;
                            float4 h = make_float4(0.f, 0.f, 0.f, 0.f), l = h;
                            if (src >= 0 && src < RAWW && j < SEG + K - 1) {
                                const float* rp = raw + (c * 4) * RAWW + src;
                                float v0 = rp[0], v1 = rp[RAWW], v2 = rp[2 * RAWW], v3 = rp[3 * RAWW];
                                if (p.in_scale) {
                                    const float4 sv = *reinterpret_cast<const float4*>(sc_s + kb * MK + c * 4);
                                    v0 *= sv.x; v1 *= sv.y; v2 *= sv.z; v3 *= sv.w;
                                }
                                split_tf32(v0, h.x, l.x); split_tf32(v1, h.y, l.y);
                                split_tf32(v2, h.z, l.z); split_tf32(v3, h.w, l.w);
                            }
                            const uint32_t off = (uint32_t)j * 128u + (uint32_t)((c ^ (j & 7)) << 4);
                            *reinterpret_cast<float4*>(hi + off) = h;
                            *reinterpret_cast<float4*>(lo + off) = l;
                        }
                    }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) { mbar_arrive(BAR_CVT_FULL(s)); mbar_arrive(BAR_RAW_EMPTY(s)); }
                    const int n16 = min(2, p.num_g16 - 2 * kb);
                    const float nz_here = (kb == p.num_kb - 1) ? fetch_nz(un, y) : 0.f;   // issued before the drains below
                    flush();
                    pend = true; pend_ac = ac; pend_n = n16; pend_ph = ph; pend_un = un; pend_y = y; pend_nz = nz_here;
                    pend_last = (kb == p.num_kb - 1);
                    ac += n16;
                    ++sc;
                }
            }
        }
        flush();
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int NT, int K, bool EPI>
int launch_march(const CUtensorMap& xmap, const MarchP& p, cudaStream_t st) {
    const MLayout L = make_mlayout(K * NT);
    const size_t smem = L.total + 1024;
    if (smem > 227 * 1024) { gg::set_error("conv2d(march): shared-memory layout of %zu bytes does not fit", smem); return GG_EUNSUPPORTED; }
    static std::atomic<uint64_t> attr_set{0};
    if (!gg::done_on_this_device(attr_set)) {
        GG_CUDA(cudaFuncSetAttribute(conv_march_kernel<NT, K, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
        gg::mark_done_on_this_device(attr_set);
    }
    const int grid = p.total_units < GG_NUM_SMS ? p.total_units : GG_NUM_SMS;
    wd_arm();
    conv_march_kernel<NT, K, EPI><<<grid, M_THREADS, smem, st>>>(xmap, p);
    return gg::check_launch("conv2d(march)");
}

}  // namespace

namespace gg {

void keep_pool_memory();   // conv_tc.cu

bool conv2d_march_eligible(const float* x, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int stride, int pad_y, int pad_x) {
    if (stride != 1 || KH != 3 || KW != 3) return false;
    if (pad_y > 2 || pad_x > 2 || pad_y < 0 || pad_x < 0) return false;
    if (N < 1 || I < 16 || I > M_MAX_CH || O < 16 || O > 64) return false;
    if (W % 4 != 0 || OW < 64 || OH < 8) return false;       // narrow maps leave most of a 128-pixel segment empty: conv_tc.cu serves them
    (void)H;
    return (reinterpret_cast<uintptr_t>(x) & 15) == 0;
}

int conv2d_march(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int K, int OH, int OW, int pad_y, int pad_x,
                 int flip_w, int w_is_IO, const float* in_scale, const float* out_scale, int nprod, const ConvEpilogue* epi, cudaStream_t st) {
    const int NT = O > 32 ? 64 : 32;
    const int Ncols = K * NT;
    const int num_g16 = (I + 15) / 16, num_kb = (I + MK - 1) / MK;

    keep_pool_memory();
    const size_t wp_floats = (size_t)num_g16 * K * 2 * 4 * Ncols * 4;
    float* wp = nullptr;
    GG_CUDA(cudaMallocAsync(&wp, wp_floats * sizeof(float), st));
    {
        MPackP pp{w, wp, O, I, K, NT, num_g16, flip_w, w_is_IO};
        const int64_t threads = (int64_t)num_g16 * K * 4 * Ncols;
        int grid = (int)((threads + 255) / 256);
        if (grid > GG_NUM_SMS * 8) grid = GG_NUM_SMS * 8;
        march_pack_weights<<<grid, 256, 0, st>>>(pp);
        int rc = check_launch("conv2d(march) pack_weights");
        if (rc != GG_OK) { cudaFreeAsync(wp, st); return rc; }
    }
    gg::EncodeTiledFn encode = gg::get_encode_fn();
    if (!encode) { cudaFreeAsync(wp, st); set_error("conv2d(march): cuTensorMapEncodeTiled is unavailable"); return GG_ECUDA; }
    CUtensorMap xmap;
    cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)I, (cuuint64_t)N};
    cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4, (cuuint64_t)W * H * I * 4};
    cuuint32_t box[4] = {(cuuint32_t)RAWW, 1, (cuuint32_t)MK, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult cr = encode(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(x), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) { cudaFreeAsync(wp, st); set_error("conv2d(march): cuTensorMapEncodeTiled failed (%d)", (int)cr); return GG_ECUDA; }

    MarchP p{};
    p.wp = wp; p.y = y; p.in_scale = in_scale; p.out_scale = out_scale;
    p.Nimg = N; p.I = I; p.O = O; p.H = H; p.W = W; p.OH = OH; p.OW = OW; p.pad_y = pad_y; p.pad_x = pad_x;
    p.num_kb = num_kb; p.num_g16 = num_g16;
    p.strips_x = (OW + SEG - 1) / SEG;
    int band = 128;                                           // rows per unit: halve until there are >= 2 waves of units
    while (band > 16 && (int64_t)N * p.strips_x * ((OH + band - 1) / band) < 2 * GG_NUM_SMS) band >>= 1;
    p.band_rows = band; p.bands = (OH + band - 1) / band;
    const int64_t total = (int64_t)N * p.strips_x * p.bands;
    if (total > 0x7fffffffLL) { cudaFreeAsync(wp, st); set_error("conv2d(march): too many units"); return GG_EINVAL; }
    p.total_units = (int)total;
    p.nprod = (nprod == GG_PREC_TF32X1) ? 1 : 3;
    p.resident = (num_g16 * K <= march_stages(Ncols)) ? 1 : 0;
    if (epi) p.epi = *epi;
    int rc;
    const bool e = p.epi.act != 0;      // the fused-epilogue instantiations are separate kernels
    if (K == 3 && NT == 32) rc = e ? launch_march<32, 3, true>(xmap, p, st) : launch_march<32, 3, false>(xmap, p, st);
    else if (K == 3 && NT == 64) rc = e ? launch_march<64, 3, true>(xmap, p, st) : launch_march<64, 3, false>(xmap, p, st);
    else { set_error("conv2d(march): unsupported configuration"); rc = GG_EUNSUPPORTED; }
    cudaFreeAsync(wp, st);
    return rc;
}

}  // namespace gg
