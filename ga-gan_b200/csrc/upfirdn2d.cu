// upfirdn2d for sm_100a: pad -> zero-stuff (up) -> FIR -> decimate (down), per channel plane.
//
// Replaces torch_utils/ops/upfirdn2d.cu:29-200 + upfirdn2d.cpp:16-94 of the reference.
//
//   y[oy][ox] = sum_{ky,kx} K[ky][kx] * U[oy*downy + ky][ox*downx + kx]
//   U[uy][ux] = x[(uy-pady0)/upy][(ux-padx0)/upx]  when divisible and inside the image, else 0
//   K[ky][kx] = gain * (flip ? f[ky][kx] : f[fH-1-ky][fW-1-kx])
//
// which is upfirdn2d.py:195-218 (zero-stuff, pad/crop, correlate with the flipped filter, keep every
// down-th sample).  The path is HBM-bound (4 B*(in+out) per plane, 16 MAC/output for the 4x4 filter):
//
//  * fir4_tile<UP,DOWN,PEX,OXT,OYT>: the StyleGAN2 cases (4x4 filter; up2 / down2 / filter-only).  One
//    CTA stages the input footprint of a (32*OXT)x(8*OYT) output tile in shared memory with coalesced
//    loads (zero-filled outside the image = the padding), then every thread produces an OXT x OYT
//    micro-tile from 128/64-bit shared-memory window loads that are reused across the filter taps and
//    across the micro-tile rows, and writes 128-bit rows.  Polyphase: for UP=2 only the live taps
//    ((t+kx) parity == PEX) are evaluated.
//  * upfirdn2d_generic: any filter / factors / padding (separable 12-tap ADA filters, anisotropic
//    scaling, tiny images): one thread per output, filter in shared memory.
#include "common.cuh"

namespace {

struct Params {
    const float* x;
    const float* f;
    float* y;
    int N, C, inH, inW, fH, fW, upx, upy, downx, downy, padx0, pady0, flip;
    float gain;
    int outH, outW;
};

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) upfirdn2d_generic(Params p) {
    extern __shared__ float sK[];
    const int nf = p.fH * p.fW;
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        int ky = i / p.fW, kx = i - ky * p.fW;
        int sy = p.flip ? ky : p.fH - 1 - ky, sx = p.flip ? kx : p.fW - 1 - kx;
        sK[i] = p.gain * __ldg(p.f + sy * p.fW + sx);
    }
    __syncthreads();
    const int64_t total = (int64_t)p.N * p.C * p.outH * p.outW;
    for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        int ox = (int)(idx % p.outW);
        int64_t r = idx / p.outW;
        int oy = (int)(r % p.outH);
        int64_t nc = r / p.outH;
        const float* xp = p.x + nc * (int64_t)p.inH * p.inW;
        // first live tap and its input coordinate (exact integer math; bit-exact with the reference's floor_div)
        int ky0 = gg::posmod(p.pady0 - oy * p.downy, p.upy);
        int iy0 = (oy * p.downy + ky0 - p.pady0) / p.upy;
        int kx0 = gg::posmod(p.padx0 - ox * p.downx, p.upx);
        int ix0 = (ox * p.downx + kx0 - p.padx0) / p.upx;
        float v = 0.f;
        for (int ky = ky0, iy = iy0; ky < p.fH; ky += p.upy, ++iy) {
            if (iy < 0 || iy >= p.inH) continue;
            for (int kx = kx0, ix = ix0; kx < p.fW; kx += p.upx, ++ix) {
                if (ix < 0 || ix >= p.inW) continue;
                v = fmaf(sK[ky * p.fW + kx], __ldg(xp + (int64_t)iy * p.inW + ix), v);
            }
        }
        p.y[idx] = v;
    }
}

// ------------------------------------------------------------------------------------------------
template <int UP, int DOWN, int OXT, int OYT>
struct TileGeom {
    static constexpr int TW = 32 * OXT;                             // output tile width  (one warp = one row strip)
    static constexpr int TH = 8 * OYT;                              // output tile height (8 warps)
    static constexpr int WN = (UP == 1) ? (OXT - 1) * DOWN + 4 : 4; // input columns one thread touches per row
    static constexpr int WNV = (WN + 3) / 4 * 4;                    // ... rounded to whole vector loads
    static constexpr int LSTEP = OXT * DOWN / UP;                   // per-lane column step inside the tile
    static constexpr int TC = ((31 * LSTEP + WNV) + 3) / 4 * 4;     // tile columns (pitch), multiple of 4
    static constexpr int RR = (OYT - 1) * DOWN + 4;                 // u-rows one thread touches
    static constexpr int TR = ((TH - 1) * DOWN + 3) / UP + 2;       // tile rows
};

template <int UP, int DOWN, int PEX, int OXT, int OYT>
__global__ void __launch_bounds__(256) fir4_tile(Params p, int tilesX, int tilesY) {
    using G = TileGeom<UP, DOWN, OXT, OYT>;
    __shared__ __align__(16) float tile[G::TR * G::TC];
    __shared__ float sK[16];

    int64_t bid = blockIdx.x;
    const int tx = (int)(bid % tilesX); bid /= tilesX;
    const int ty = (int)(bid % tilesY);
    const int64_t nc = bid / tilesY;
    const int ox_t0 = tx * G::TW, oy_t0 = ty * G::TH;
    const int ix0 = gg::floordiv(ox_t0 * DOWN - p.padx0 + UP - 1, UP);
    const int iy0 = gg::floordiv(oy_t0 * DOWN - p.pady0 + UP - 1, UP);

    if (threadIdx.x < 16) {
        int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
        int sy = p.flip ? ky : 3 - ky, sx = p.flip ? kx : 3 - kx;
        sK[threadIdx.x] = p.gain * __ldg(p.f + sy * 4 + sx);
    }
    // stage the input footprint (zero outside the image: that is the padding / crop)
    const float* xp = p.x + nc * (int64_t)p.inH * p.inW;
    for (int i = threadIdx.x; i < G::TR * G::TC; i += 256) {
        int r = i / G::TC, c = i - r * G::TC;
        int gy = iy0 + r, gx = ix0 + c;
        float v = 0.f;
        if (gy >= 0 && gy < p.inH && gx >= 0 && gx < p.inW) v = __ldg(xp + (int64_t)gy * p.inW + gx);
        tile[i] = v;
    }
    __syncthreads();

    float K[4][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) K[i >> 2][i & 3] = sK[i];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int oy0 = oy_t0 + warp * OYT;      // first output row of this thread
    const int ox0 = ox_t0 + lane * OXT;      // first output column of this thread
    float acc[OYT][OXT];
#pragma unroll
    for (int r = 0; r < OYT; ++r)
#pragma unroll
        for (int t = 0; t < OXT; ++t) acc[r][t] = 0.f;

    const float* wbase = tile + lane * G::LSTEP;
#pragma unroll
    for (int rr = 0; rr < G::RR; ++rr) {
        const int uy = oy0 * DOWN + rr - p.pady0;
        if (UP == 2 && (uy & 1)) continue;                 // dead u-row (warp-uniform)
        const int trow = (UP == 2 ? (uy >> 1) : uy) - iy0;
        if (trow < 0 || trow >= G::TR) continue;           // only rows that belong to out-of-range outputs
        const float* src = wbase + trow * G::TC;
        float w[G::WNV];
        if (G::LSTEP % 4 == 0) {
#pragma unroll
            for (int j = 0; j < G::WNV; j += 4) {
                float4 v = *reinterpret_cast<const float4*>(src + j);
                w[j] = v.x; w[j + 1] = v.y; w[j + 2] = v.z; w[j + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int j = 0; j < G::WNV; j += 2) {
                float2 v = *reinterpret_cast<const float2*>(src + j);
                w[j] = v.x; w[j + 1] = v.y;
            }
        }
#pragma unroll
        for (int r = 0; r < OYT; ++r) {
            const int ky = rr - r * DOWN;
            if (ky < 0 || ky > 3) continue;                // compile-time after unrolling
#pragma unroll
            for (int t = 0; t < OXT; ++t)
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    if (UP == 1) {
                        acc[r][t] = fmaf(K[ky][kx], w[t * DOWN + kx], acc[r][t]);
                    } else if (((t + kx) & 1) == PEX) {
                        acc[r][t] = fmaf(K[ky][kx], w[(t + kx - PEX) / 2], acc[r][t]);
                    }
                }
        }
    }

    float* yp = p.y + nc * (int64_t)p.outH * p.outW;
    const bool vec = (OXT == 4) && (p.outW % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
#pragma unroll
    for (int r = 0; r < OYT; ++r) {
        const int oy = oy0 + r;
        if (oy >= p.outH) continue;
        float* dst = yp + (int64_t)oy * p.outW + ox0;
        if (vec && ox0 + 3 < p.outW) {
            *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
        } else {
#pragma unroll
            for (int t = 0; t < OXT; ++t)
                if (ox0 + t < p.outW) dst[t] = acc[r][t];
        }
    }
}

template <int UP, int DOWN, int PEX, int OXT, int OYT>
int launch_tile(const Params& p, cudaStream_t st) {
    using G = TileGeom<UP, DOWN, OXT, OYT>;
    int tilesX = (p.outW + G::TW - 1) / G::TW, tilesY = (p.outH + G::TH - 1) / G::TH;
    int64_t blocks = (int64_t)tilesX * tilesY * p.N * p.C;
    if (blocks > 0x7fffffffLL) { gg::set_error("upfirdn2d: grid too large"); return GG_EINVAL; }
    fir4_tile<UP, DOWN, PEX, OXT, OYT><<<(unsigned)blocks, 256, 0, st>>>(p, tilesX, tilesY);
    return gg::check_launch("upfirdn2d(fir4_tile)");
}


// ------------------------------------------------------------------------------------------------
// fir_stream<IN_PM, OUT_PM>: 4x4 FIR at unit rate (up = down = 1) that streams rows through registers and reads / writes
// the PHASE-MAJOR layout of conv2d_resample.py on the fly, so that the space-to-depth / depth-to-space passes around the
// tensor-core convolutions of the stride-2 layers cost no extra HBM traffic:
//
//   phase-major tensor  t_pm[n, (py,px,c), Y, X]  <->  logical image  t[n, c, 2Y+py, 2X+px]
//
//   IN_PM : the input is phase-major [N,4C,pmH,pmW]; logical pixels outside (validH, validW) count as zero
//   OUT_PM: the output is phase-major [N,4C,pmH,pmW]; logical pixels outside (outH, outW) are written as zero
//
// One thread produces OXT consecutive output columns of RY consecutive rows: every input row is loaded once (scalar,
// coalesced across the warp; the 3-column overlap between neighbouring threads is an L1 hit) and feeds the up-to-4 output
// rows it contributes to.  Stores are 128-bit.
struct StreamP {
    const float* x; const float* f; float* y;
    int N, C, inH, inW, padx0, pady0, flip;     // logical input extent (IN_PM: the valid extent)
    float gain;
    int outH, outW;                             // logical (valid) output extent
    int ipH, ipW, opH, opW;                     // phase-major plane sizes (IN_PM / OUT_PM)
};

// LAYOUT 0: plain -> plain, 1: phase-major in -> plain out, 2: plain in -> phase-major out.  PX0 = padx0 (0..3) is a template
// parameter so that the position of every needed input column inside the aligned 128-bit blocks is known at compile time.
// One thread = 8 output columns x 8 output rows; per input row it issues 4 (plain) or 6 (phase-major: 3 per column phase)
// aligned 128-bit loads; rows and column blocks that stick out of the image take a guarded scalar path.
template <int LAYOUT, int PX0>
__global__ void __launch_bounds__(256) fir_stream(StreamP p) {
    constexpr bool IN_PM = LAYOUT == 1, OUT_PM = LAYOUT == 2;
    constexpr int OXT = 8, RY = 8;
    constexpr int NB = IN_PM ? 3 : 4;                 // 128-bit blocks per row (and per column phase for IN_PM)
    __shared__ float sK[16];
    if (threadIdx.x < 16) {
        int ky = threadIdx.x >> 2, kx = threadIdx.x & 3;
        int sy = p.flip ? ky : 3 - ky, sx = p.flip ? kx : 3 - kx;
        sK[threadIdx.x] = p.gain * __ldg(p.f + sy * 4 + sx);
    }
    __syncthreads();
    float K[4][4];
#pragma unroll
    for (int i = 0; i < 16; ++i) K[i >> 2][i & 3] = sK[i];
    // K == fy (x) fx ?  (exact test on the 16 coefficients; warp- and grid-uniform)
    float fx[4], fy[4];
    bool separable = K[0][0] != 0.f;
    {
        float amax = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) amax = fmaxf(amax, fabsf(K[i >> 2][i & 3]));
#pragma unroll
        for (int i = 0; i < 4; ++i) { fx[i] = K[0][i]; fy[i] = separable ? K[i][0] / K[0][0] : 0.f; }
#pragma unroll
        for (int i = 0; i < 16; ++i) separable = separable && fabsf(K[i >> 2][i & 3] - fy[i >> 2] * fx[i & 3]) <= 1e-7f * amax;
    }

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int x0 = (blockIdx.x * 32 + lane) * OXT;
    const int y0 = (blockIdx.y * 8 + warp) * RY;
    const int fullW = OUT_PM ? 2 * p.opW : p.outW, fullH = OUT_PM ? 2 * p.opH : p.outH;
    if (x0 >= fullW || y0 >= fullH) return;
    const int pitch = IN_PM ? p.ipW : p.inW;          // floats per stored input row
    const size_t iplane = IN_PM ? (size_t)p.ipH * p.ipW : (size_t)p.inH * p.inW;
    // first column of block 0 in the stored row: logical x0-4 (plain) or phase-major column x0/2-4
    const int cb0 = IN_PM ? (x0 >> 1) - 4 : x0 - 4;
    const int climit = IN_PM ? p.ipW : p.inW;         // stored columns
    // the valid extent may end inside the stored row (IN_PM): logical col < inW  <=>  2X+px < inW
    for (int nc = blockIdx.z; nc < p.N * p.C; nc += gridDim.z) {
        const int n = nc / p.C, c = nc - n * p.C;
        const float* xin = p.x + (IN_PM ? ((size_t)n * 4 * p.C + c) * iplane : (size_t)nc * iplane);

        float acc[RY][OXT];
#pragma unroll
        for (int r = 0; r < RY; ++r)
#pragma unroll
            for (int t = 0; t < OXT; ++t) acc[r][t] = 0.f;

#pragma unroll
        for (int rr = 0; rr < RY + 3; ++rr) {
            const int r = y0 + rr - p.pady0;                    // logical input row
            const bool row_ok = r >= 0 && r < p.inH;
            float in[OXT + 3];
            if (!IN_PM) {
                float buf[4 * NB];
                const float* rowp = xin + (size_t)(row_ok ? r : 0) * pitch;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const int cb = cb0 + 4 * b;
                    if (row_ok && cb >= 0 && cb + 3 < climit) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(rowp + cb));
                        buf[4 * b] = v.x; buf[4 * b + 1] = v.y; buf[4 * b + 2] = v.z; buf[4 * b + 3] = v.w;
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            buf[4 * b + e] = (row_ok && cb + e >= 0 && cb + e < climit) ? __ldg(rowp + cb + e) : 0.f;
                    }
                }
#pragma unroll
                for (int j = 0; j < OXT + 3; ++j) in[j] = buf[4 - PX0 + j];
            } else {
                // logical column x0 - PX0 + j lives in column phase ph = (j - PX0) & 1 at phase-major column x0/2 + (j - PX0 - ph)/2
                float buf[2][4 * NB];
                const int py = r & 1;
#pragma unroll
                for (int ph = 0; ph < 2; ++ph) {
                    const float* rowp = xin + (size_t)((py * 2 + ph) * p.C) * iplane + (size_t)(row_ok ? (r >> 1) : 0) * pitch;
                    const int vcols = (p.inW - ph + 1) >> 1;    // valid phase-major columns of this phase
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        const int cb = cb0 + 4 * b;
                        if (row_ok && cb >= 0 && cb + 3 < vcols) {
                            const float4 v = __ldg(reinterpret_cast<const float4*>(rowp + cb));
                            buf[ph][4 * b] = v.x; buf[ph][4 * b + 1] = v.y; buf[ph][4 * b + 2] = v.z; buf[ph][4 * b + 3] = v.w;
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                buf[ph][4 * b + e] = (row_ok && cb + e >= 0 && cb + e < vcols) ? __ldg(rowp + cb + e) : 0.f;
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < OXT + 3; ++j) {
                    constexpr int dummy = 0; (void)dummy;
                    const int d = j - PX0;                       // compile-time after unrolling
                    const int ph = d & 1;
                    in[j] = buf[ph][4 + (d - ph) / 2];
                }
            }
            if (separable) {
                // rank-1 filter (setup_filter builds the 2-D [1,3,3,1] filter as an outer product): horizontal pass once per input
                // row, then one FMA per output row -- 8 instead of 16 FMAs per output
                float h[OXT];
#pragma unroll
                for (int t = 0; t < OXT; ++t)
                    h[t] = fmaf(fx[3], in[t + 3], fmaf(fx[2], in[t + 2], fmaf(fx[1], in[t + 1], fx[0] * in[t])));
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const int yy = rr - ky;
                    if (yy < 0 || yy >= RY) continue;
#pragma unroll
                    for (int t = 0; t < OXT; ++t) acc[yy][t] = fmaf(fy[ky], h[t], acc[yy][t]);
                }
            } else {
#pragma unroll
                for (int ky = 0; ky < 4; ++ky) {
                    const int yy = rr - ky;                      // output row (relative) fed through filter row ky
                    if (yy < 0 || yy >= RY) continue;            // compile-time after unrolling
#pragma unroll
                    for (int t = 0; t < OXT; ++t)
#pragma unroll
                        for (int kx = 0; kx < 4; ++kx) acc[yy][t] = fmaf(K[ky][kx], in[t + kx], acc[yy][t]);
                }
            }
        }

#pragma unroll
        for (int r = 0; r < RY; ++r) {
            const int y = y0 + r;
            if (y >= fullH) continue;
            if (OUT_PM) {
                // logical (y, x0 + 2X' + px) -> plane (y&1, px, c), row y>>1, columns x0/2 + X'
                float v[OXT];
#pragma unroll
                for (int t = 0; t < OXT; ++t) v[t] = (y < p.outH && x0 + t < p.outW) ? acc[r][t] : 0.f;
                const size_t oplane = (size_t)p.opH * p.opW;
                float* base = p.y + ((size_t)n * 4 * p.C + (size_t)((y & 1) * 2) * p.C + c) * oplane + (size_t)(y >> 1) * p.opW + (x0 >> 1);
                *reinterpret_cast<float4*>(base) = make_float4(v[0], v[2], v[4], v[6]);
                *reinterpret_cast<float4*>(base + (size_t)p.C * oplane) = make_float4(v[1], v[3], v[5], v[7]);
            } else {
                float* dst = p.y + (size_t)nc * p.outH * p.outW + (size_t)y * p.outW + x0;
                if ((p.outW & 3) == 0 && x0 + 7 < p.outW) {
                    *reinterpret_cast<float4*>(dst) = make_float4(acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                    *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[r][4], acc[r][5], acc[r][6], acc[r][7]);
                } else {
#pragma unroll
                    for (int t = 0; t < OXT; ++t)
                        if (x0 + t < p.outW) dst[t] = acc[r][t];
                }
            }
        }
    }
}

template <int LAYOUT, int PX0>
int launch_stream2(const StreamP& p, cudaStream_t st) {
    constexpr int OXT = 8, RY = 8;
    const int fullW = LAYOUT == 2 ? 2 * p.opW : p.outW, fullH = LAYOUT == 2 ? 2 * p.opH : p.outH;
    dim3 grid((unsigned)((fullW + 32 * OXT - 1) / (32 * OXT)), (unsigned)((fullH + 8 * RY - 1) / (8 * RY)),
              (unsigned)(p.N * p.C < 65535 ? p.N * p.C : 65535));
    fir_stream<LAYOUT, PX0><<<grid, 256, 0, st>>>(p);
    return gg::check_launch("upfirdn2d(fir_stream)");
}

template <int LAYOUT>
int launch_stream(const StreamP& p, cudaStream_t st) {
    switch (p.padx0) {
        case 0: return launch_stream2<LAYOUT, 0>(p, st);
        case 1: return launch_stream2<LAYOUT, 1>(p, st);
        case 2: return launch_stream2<LAYOUT, 2>(p, st);
        default: return launch_stream2<LAYOUT, 3>(p, st);
    }
}

// 128-bit row loads need 16-byte aligned rows; padx0 must be one of the four instantiations
bool stream_ok(const StreamP& p, bool in_pm) {
    const int pitch = in_pm ? p.ipW : p.inW;
    return p.padx0 >= 0 && p.padx0 <= 3 && pitch % 4 == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0;
}

}  // namespace

extern "C" GG_API int gg_upfirdn2d_f32(const float* x, const float* f, float* y, int N, int C, int inH, int inW, int fH, int fW,
                                int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1,
                                int flip, float gain, int outH, int outW, gg_stream_t stream) {
    GG_REQUIRE(x && f && y, "upfirdn2d: null pointer");
    GG_REQUIRE(N >= 0 && C >= 0 && inH >= 1 && inW >= 1, "upfirdn2d: x must be rank 4 with non-empty planes");
    GG_REQUIRE(fH >= 1 && fW >= 1, "upfirdn2d: f must be at least 1x1");                                  // upfirdn2d.cpp:26
    GG_REQUIRE(upx >= 1 && upy >= 1, "upfirdn2d: upsampling factor must be at least 1");                  // :27
    GG_REQUIRE(downx >= 1 && downy >= 1, "upfirdn2d: downsampling factor must be at least 1");            // :28
    const int ew = (inW * upx + padx0 + padx1 - fW + downx) / downx;                                       // :32
    const int eh = (inH * upy + pady0 + pady1 - fH + downy) / downy;                                       // :33
    GG_REQUIRE(ew >= 1 && eh >= 1, "upfirdn2d: output must be at least 1x1");                             // :34
    GG_REQUIRE(ew == outW && eh == outH, "upfirdn2d: output size mismatch (expected %dx%d, got %dx%d)", eh, ew, outH, outW);
    GG_REQUIRE((int64_t)N * C * inH * inW <= 0x7fffffffLL && (int64_t)N * C * outH * outW <= 0x7fffffffLL,
               "upfirdn2d: tensor is too large");                                                          // :22,36
    if ((int64_t)N * C == 0) return GG_OK;
    Params p{x, f, y, N, C, inH, inW, fH, fW, upx, upy, downx, downy, padx0, pady0, flip, gain, outH, outW};
    cudaStream_t st = (cudaStream_t)stream;

    const bool f4 = (fH == 4 && fW == 4) && upx == upy && downx == downy && outW >= 48;
    if (f4 && upx == 1 && downx == 1) {   // unit rate: the register-streaming kernel (also serves the phase-major layouts)
        StreamP sp{x, f, y, N, C, inH, inW, padx0, pady0, flip, gain, outH, outW, 0, 0, 0, 0};
        if (stream_ok(sp, false)) return launch_stream<0>(sp, st);
        return launch_tile<1, 1, 0, 4, 4>(p, st);
    }
    if (f4 && upx == 1 && downx == 2) return launch_tile<1, 2, 0, 2, 2>(p, st);
    if (f4 && upx == 2 && downx == 1) {
        // PEX = parity of (ox0 - padx0) for the first column of any thread (ox0 is a multiple of 4)
        return (padx0 & 1) ? launch_tile<2, 1, 1, 4, 2>(p, st) : launch_tile<2, 1, 0, 4, 2>(p, st);
    }
    GG_REQUIRE((size_t)fH * fW * sizeof(float) <= 48 * 1024, "upfirdn2d: filter too large");
    int64_t total = (int64_t)N * C * outH * outW;
    int64_t grid = (total + 255) / 256;
    if (grid > GG_NUM_SMS * 32) grid = GG_NUM_SMS * 32;
    upfirdn2d_generic<<<(unsigned)grid, 256, (size_t)fH * fW * sizeof(float), st>>>(p);
    return gg::check_launch("upfirdn2d(generic)");
}


extern "C" GG_API int gg_fir4_pm_f32(const float* x, const float* f, float* y, int N, int C, int inH, int inW, int padx0, int pady0,
                              int flip, float gain, int outH, int outW, int in_pm, int in_pmH, int in_pmW, int out_pm, int out_pmH,
                              int out_pmW, gg_stream_t stream) {
    GG_REQUIRE(x && f && y, "fir4_pm: null pointer");
    GG_REQUIRE(N >= 0 && C >= 1 && inH >= 1 && inW >= 1 && outH >= 1 && outW >= 1, "fir4_pm: bad shape");
    GG_REQUIRE(!(in_pm && out_pm), "fir4_pm: at most one side may be phase-major");
    GG_REQUIRE(!in_pm || (in_pmH >= 1 && in_pmW >= 1 && inH <= 2 * in_pmH && inW <= 2 * in_pmW), "fir4_pm: phase-major input planes are smaller than the valid extent");
    GG_REQUIRE(!out_pm || (out_pmH >= 1 && out_pmW >= 1 && out_pmW % 4 == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0),
               "fir4_pm: phase-major output planes must be 16-byte aligned with a width that is a multiple of 4");
    GG_REQUIRE((int64_t)N * C * (in_pm ? 4LL * in_pmH * in_pmW : (int64_t)inH * inW) <= 0x7fffffffLL &&
               (int64_t)N * C * (out_pm ? 4LL * out_pmH * out_pmW : (int64_t)outH * outW) <= 0x7fffffffLL, "fir4_pm: tensor is too large");
    if (N == 0) return GG_OK;
    StreamP p{x, f, y, N, C, inH, inW, padx0, pady0, flip, gain, outH, outW, in_pmH, in_pmW, out_pmH, out_pmW};
    cudaStream_t st = (cudaStream_t)stream;
    GG_REQUIRE(stream_ok(p, in_pm != 0), "fir4_pm: needs 0 <= padx0 <= 3 and 16-byte aligned input rows (width %% 4 == 0)");
    if (in_pm) return launch_stream<1>(p, st);
    if (out_pm) return launch_stream<2>(p, st);
    return launch_stream<0>(p, st);
}
