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
// fir_stream<LAYOUT, PX0>: 4x4 FIR at unit rate (up = down = 1) that reads / writes the PHASE-MAJOR layout of
// conv2d_resample.py on the fly, so that the space-to-depth / depth-to-space passes around the tensor-core convolutions of
// the stride-2 layers cost no extra HBM traffic:
//
//   phase-major tensor  t_pm[n, (py,px,c), Y, X]  <->  logical image  t[n, c, 2Y+py, 2X+px]
//
//   LAYOUT 0: plain -> plain
//   LAYOUT 1: the input is phase-major [N,4C,ipH,ipW]; logical pixels outside (inH, inW) count as zero
//   LAYOUT 2: the output is phase-major [N,4C,opH,opW]; logical pixels outside (outH, outW) are written as zero
//
// One thread owns 8 output columns and MARCHES DOWN a strip of RS rows: every input row is loaded once (the 3 halo rows of a
// strip twice), runs through the horizontal taps, and updates three running partial sums -- the rows of the vertical taps that
// are still open -- so an output row leaves the thread three input rows after its first contribution.  The loads of the row
// two steps ahead are in flight while a row is processed (~200 bytes per thread in flight, 4-5 CTAs of 128 threads per SM:
// enough to cover the HBM latency-bandwidth product).  PX0 = padx0 (0..3) is a template parameter: the position of every
// needed column relative to the thread's aligned 8-column block is a compile-time constant, so a row is 2 aligned 128-bit
// loads plus the left / right halo (a 128-bit load each, of which only the needed lanes are kept); threads whose blocks stick
// out of the image (and rows outside it) take guarded scalar loads -- that is the zero padding.  Index math is integer and exact.
struct StreamP {
    const float* x; const float* f; float* y;
    int N, C, inH, inW, padx0, pady0, flip;     // logical input extent (LAYOUT 1: the valid extent)
    float gain;
    int outH, outW;                             // logical (valid) output extent
    int ipH, ipW, opH, opW;                     // phase-major plane sizes (LAYOUT 1 / 2)
    int RS, ncg, nst;                           // rows per strip, 8-column groups per row, strips per plane
    int vec_in;                                 // plain input rows are 16-byte aligned (width % 4 == 0): 128-bit loads, else scalar
    int vec_out;                                // plain output rows are 16-byte aligned: 128-bit stores
};

// STG (plain -> plain only): rows that are not 16-byte multiples (the reference-shaped (2r+1)^2 maps) are exchanged through a
// per-warp shared-memory buffer so that every global access of the warp is one contiguous 128-byte segment: the warp loads its
// 288-column window of an input row with nine coalesced instructions and hands each lane its 16 columns, and the eight outputs per
// lane are written back as eight coalesced row segments (through per-lane row pointers, so the lanes of a warp may sit on different
// rows).  Every lane runs RS + 3 steps whatever its strip (full-mask __syncwarp); for the input window the launch also pads the
// column groups of a row to a multiple of 32, so that a warp's lanes share one row.
struct StageBuf { float* v; float** ptr; int* nv; };           // per warp: 288 floats, 32 row pointers, 32 valid counts

template <int LAYOUT, int PX0, bool SEP, bool STG>
__device__ __forceinline__ void fir_march(const StreamP& p, const float (&K)[4][4], const float (&fx)[4], const float (&fy)[4],
                                          int x0, int y0, int n, int c, const StageBuf sb) {
    constexpr bool IN_PM = LAYOUT == 1, OUT_PM = LAYOUT == 2;
    static_assert(!STG || LAYOUT == 0, "the staged row exchange serves the plain -> plain layout");
    const int lane = threadIdx.x & 31;
    const bool raw_rows = STG && !p.vec_in;                     // rows travel as the warp's coalesced window (slots 0..8 of the row buffer)
    const int fullH = OUT_PM ? 2 * p.opH : p.outH;
    const int nrows = min(p.RS, fullH - y0);
    const int nt = (STG ? p.RS : nrows) + 3;                    // input rows this strip touches (STG: the same count in every lane of a warp,
                                                                // whatever strip it works on -- rows beyond the image load zeros and store nothing)
    const int rbase = y0 - p.pady0;                             // logical input row of step 0
    const int pitch = IN_PM ? p.ipW : p.inW;
    const size_t iplane = IN_PM ? (size_t)p.ipH * p.ipW : (size_t)p.inH * p.inW;
    const float* xin = p.x + (IN_PM ? ((size_t)n * 4 * p.C + c) * iplane : ((size_t)n * p.C + c) * iplane);
    const int h0 = x0 >> 1;                                     // IN_PM: first phase-major column of the thread's block

    // One input row -> in[0..10] = logical columns x0 - PX0 .. x0 - PX0 + 10 (zero outside the image).
    auto load_row = [&](int rr, float (&in)[11]) {
        const int r = rbase + rr;
        const bool row_ok = (unsigned)r < (unsigned)p.inH;
        if (STG && raw_rows) {
            const float* roww = xin + (size_t)(row_ok ? r : 0) * pitch;
            const int w0 = x0 - 8 * lane - 4;                    // first column of the warp's window
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const int col = w0 + 32 * i + lane;
                in[i] = (row_ok && col >= 0 && col < p.inW) ? __ldg(roww + col) : 0.f;
            }
            in[9] = 0.f; in[10] = 0.f;
        } else if (!IN_PM) {
            const float* rowp = xin + (size_t)(row_ok ? r : 0) * pitch + x0;
            float buf[16];                                       // columns x0-4 .. x0+11
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (b == 0 && PX0 == 0) continue;                // left halo not needed
                if (b == 3 && PX0 == 3) continue;                // right halo not needed
                const int cb = x0 - 4 + 4 * b;
                if (p.vec_in && row_ok && cb >= 0 && cb + 3 < p.inW) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(rowp - 4 + 4 * b));
                    buf[4 * b] = v.x; buf[4 * b + 1] = v.y; buf[4 * b + 2] = v.z; buf[4 * b + 3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        buf[4 * b + e] = (row_ok && cb + e >= 0 && cb + e < p.inW) ? __ldg(rowp - 4 + 4 * b + e) : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 11; ++j) in[j] = buf[4 - PX0 + j];
        } else {
            // logical column x0 - PX0 + j lives in column phase ph = (j - PX0) & 1 at phase-major column h0 + (j - PX0 - ph) / 2
            float buf[2][8];                                     // phase-major columns h0-2 .. h0+5 of both column phases
            const int py = r & 1;
#pragma unroll
            for (int ph = 0; ph < 2; ++ph) {
                const float* rowp = xin + (size_t)((py * 2 + ph) * p.C) * iplane + (size_t)(row_ok ? (r >> 1) : 0) * pitch + h0;
                const int vcols = (p.inW - ph + 1) >> 1;        // valid phase-major columns of this phase
                // left pair (h0-2, h0-1), middle quad (h0 .. h0+3), right pair (h0+4, h0+5)
                if (row_ok && h0 >= 2 && h0 - 1 < vcols) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(rowp - 2));
                    buf[ph][0] = v.x; buf[ph][1] = v.y;
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) buf[ph][e] = (row_ok && h0 - 2 + e >= 0 && h0 - 2 + e < vcols) ? __ldg(rowp - 2 + e) : 0.f;
                }
                if (row_ok && h0 + 3 < vcols) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(rowp));
                    buf[ph][2] = v.x; buf[ph][3] = v.y; buf[ph][4] = v.z; buf[ph][5] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e) buf[ph][2 + e] = (row_ok && h0 + e < vcols) ? __ldg(rowp + e) : 0.f;
                }
                if (row_ok && h0 + 5 < vcols) {
                    const float2 v = __ldg(reinterpret_cast<const float2*>(rowp + 4));
                    buf[ph][6] = v.x; buf[ph][7] = v.y;
                } else {
#pragma unroll
                    for (int e = 0; e < 2; ++e) buf[ph][6 + e] = (row_ok && h0 + 4 + e < vcols) ? __ldg(rowp + 4 + e) : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 11; ++j) {
                const int d = j - PX0;                           // compile-time after unrolling
                const int ph = d & 1;
                in[j] = buf[ph][(d - ph) / 2 + 2];
            }
        }
    };

    float* yout = OUT_PM ? p.y + ((size_t)n * 4 * p.C + c) * ((size_t)p.opH * p.opW) : p.y + ((size_t)n * p.C + c) * ((size_t)p.outH * p.outW);
    const bool vec_out = OUT_PM || ((p.outW & 3) == 0 && x0 + 7 < p.outW && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
    auto store_row = [&](int y, const float (&o)[8]) {
        if (OUT_PM) {
            // logical (y, x0 + 2X' + px) -> plane (y&1, px, c), row y>>1, columns x0/2 + X'
            float v[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) v[t] = (y < p.outH && x0 + t < p.outW) ? o[t] : 0.f;
            const size_t oplane = (size_t)p.opH * p.opW;
            float* base = yout + (size_t)((y & 1) * 2) * p.C * oplane + (size_t)(y >> 1) * p.opW + (x0 >> 1);
            *reinterpret_cast<float4*>(base) = make_float4(v[0], v[2], v[4], v[6]);
            *reinterpret_cast<float4*>(base + (size_t)p.C * oplane) = make_float4(v[1], v[3], v[5], v[7]);
        } else if (STG && !p.vec_out) {
            const int nv = y < p.outH ? min(8, max(0, p.outW - x0)) : 0;
            *reinterpret_cast<float4*>(sb.v + 8 * lane) = make_float4(o[0], o[1], o[2], o[3]);
            *reinterpret_cast<float4*>(sb.v + 8 * lane + 4) = make_float4(o[4], o[5], o[6], o[7]);
            sb.ptr[lane] = yout + (size_t)y * p.outW + x0;       // (only dereferenced for the nv valid columns)
            sb.nv[lane] = nv;
            __syncwarp();
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const int e = 32 * t + lane, src = e >> 3, k = e & 7;
                if (k < sb.nv[src]) sb.ptr[src][k] = sb.v[e];
            }
            __syncwarp();
        } else {
            if (y >= p.outH) return;
            float* dst = yout + (size_t)y * p.outW + x0;
            if (vec_out) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t)
                    if (x0 + t < p.outW) dst[t] = o[t];
            }
        }
    };

    // running partial sums of the three output rows that are still open: A is the oldest (one vertical tap missing)
    float A[8], B[8], Cc[8];
#pragma unroll
    for (int t = 0; t < 8; ++t) { A[t] = 0.f; B[t] = 0.f; Cc[t] = 0.f; }
    auto step = [&](int rr, const float (&in)[11]) {
        float o[8];
        if (SEP) {
            // rank-1 filter (setup_filter builds the 2-D [1,3,3,1] filter as an outer product): horizontal pass once per input
            // row, then one FMA per open output row
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                const float h = fmaf(fx[3], in[t + 3], fmaf(fx[2], in[t + 2], fmaf(fx[1], in[t + 1], fx[0] * in[t])));
                o[t] = fmaf(fy[3], h, A[t]);
                A[t] = fmaf(fy[2], h, B[t]);
                B[t] = fmaf(fy[1], h, Cc[t]);
                Cc[t] = fy[0] * h;
            }
        } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                float o3 = A[t], o2 = B[t], o1 = Cc[t], o0 = 0.f;
#pragma unroll
                for (int kx = 0; kx < 4; ++kx) {
                    o3 = fmaf(K[3][kx], in[t + kx], o3);
                    o2 = fmaf(K[2][kx], in[t + kx], o2);
                    o1 = fmaf(K[1][kx], in[t + kx], o1);
                    o0 = fmaf(K[0][kx], in[t + kx], o0);
                }
                o[t] = o3; A[t] = o2; B[t] = o1; Cc[t] = o0;
            }
        }
        if (rr >= 3) store_row(y0 + rr - 3, o);
    };

    // a row buffer holds either the lane's 11 columns or (raw_rows) the lane's 9 elements of the warp's window: exchange first
    auto consume = [&](int rr, const float (&buf)[11]) {
        if (STG && raw_rows) {
#pragma unroll
            for (int i = 0; i < 9; ++i) sb.v[32 * i + lane] = buf[i];
            __syncwarp();
            float w16[16];                                       // window columns 8 lane .. 8 lane + 15 = image columns x0 - 4 .. x0 + 11
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 t4 = *reinterpret_cast<const float4*>(sb.v + 8 * lane + 4 * q4);
                w16[4 * q4] = t4.x; w16[4 * q4 + 1] = t4.y; w16[4 * q4 + 2] = t4.z; w16[4 * q4 + 3] = t4.w;
            }
            __syncwarp();
            float in[11];
#pragma unroll
            for (int j = 0; j < 11; ++j) in[j] = w16[4 - PX0 + j];
            step(rr, in);
        } else {
            step(rr, buf);
        }
    };
    float r0[11], r1[11], r2[11];
    load_row(0, r0);
    load_row(1, r1);
    for (int rr = 0; rr < nt; rr += 3) {
        if (rr + 2 < nt) load_row(rr + 2, r2);
        consume(rr, r0);
        if (rr + 1 < nt) {
            if (rr + 3 < nt) load_row(rr + 3, r0);
            consume(rr + 1, r1);
        }
        if (rr + 2 < nt) {
            if (rr + 4 < nt) load_row(rr + 4, r1);
            consume(rr + 2, r2);
        }
    }
}

template <int LAYOUT, int PX0, bool STG>
__global__ void __launch_bounds__(128, 4) fir_stream(StreamP p) {
    __shared__ float sK[16];
    __shared__ __align__(16) float s_stage[STG ? 4 * 288 : 1];
    __shared__ float* s_ptr[STG ? 128 : 1];
    __shared__ int s_nv[STG ? 128 : 1];
    const int warp_ = threadIdx.x >> 5;
    const StageBuf sb{s_stage + (STG ? warp_ * 288 : 0), s_ptr + (STG ? warp_ * 32 : 0), s_nv + (STG ? warp_ * 32 : 0)};
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
    // thread -> (8-column group, row strip, plane); column groups fastest, so a warp reads contiguous row segments and narrow
    // images still fill their warps (with neighbouring strips / planes)
    const long long id = (long long)blockIdx.x * 128 + threadIdx.x;
    const int cg = (int)(id % p.ncg);
    const long long r = id / p.ncg;
    const int st = (int)(r % p.nst);
    const long long nc = r / p.nst;
    const bool past_end = nc >= (long long)p.N * p.C;
    if (!STG && past_end) return;
    // STG: the lanes of the grid's last warp that have no work stay for the warp-wide exchanges, parked on a column beyond both images
    // (every load is out of range and reads as zero, every store has zero valid columns)
    const int n = past_end ? 0 : (int)(nc / p.C), c = past_end ? 0 : (int)(nc - (long long)n * p.C);
    const int x0 = past_end ? ((max(p.inW, p.outW) + 31) / 8 * 8 + 64) : cg * 8;
    if (separable) fir_march<LAYOUT, PX0, true, STG>(p, K, fx, fy, x0, st * p.RS, n, c, sb);
    else fir_march<LAYOUT, PX0, false, STG>(p, K, fx, fy, x0, st * p.RS, n, c, sb);
}

template <int LAYOUT, int PX0>
int launch_stream2(StreamP p, cudaStream_t st) {
    const int fullW = LAYOUT == 2 ? 2 * p.opW : p.outW, fullH = LAYOUT == 2 ? 2 * p.opH : p.outH;
    p.RS = fullH >= 128 ? 32 : (fullH >= 32 ? 16 : 8);
    p.vec_in = (LAYOUT != 1 && p.inW % 4 == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0) ? 1 : 0;
    p.vec_out = (LAYOUT == 2 || (p.outW % 4 == 0 && (reinterpret_cast<uintptr_t>(p.y) & 15) == 0)) ? 1 : 0;
    p.ncg = (fullW + 7) / 8;
    p.nst = (fullH + p.RS - 1) / p.RS;
    // plain -> plain with unaligned rows on either side, wide enough that padding the column groups to whole warps is cheap
    // (measured: the input window pays from 257 columns on -- 57 -> 69 % of the HBM peak at N = 8, 67 -> 79 % at 1025 --, the output
    // exchange from 513 on -- 51 -> 55 %, 60 % at 1025 -- and costs a little at 257)
    // Below 256 output columns a row has fewer than 32 column groups and the padding idles up to half of every warp (129 -> 128
    // columns: 63 -> 43 %): those stay on the per-lane scalar loads.
    const bool staged = LAYOUT == 0 && ((!p.vec_in && fullW >= 256) || (!p.vec_out && fullW >= 384));
    // the coalesced input window needs the lanes of a warp on ONE row (column groups padded to whole warps); the output exchange works
    // lane by lane and needs no padding
    if (staged && !p.vec_in) p.ncg = (p.ncg + 31) / 32 * 32;
    const long long threads = (long long)p.ncg * p.nst * p.N * p.C;
    const long long blocks = (threads + 127) / 128;
    if (blocks > 0x7fffffffLL) { gg::set_error("upfirdn2d: grid too large"); return GG_EINVAL; }
    if constexpr (LAYOUT == 0) {
        if (staged) {
            fir_stream<0, PX0, true><<<(unsigned)blocks, 128, 0, st>>>(p);
            return gg::check_launch("upfirdn2d(fir_stream, staged rows)");
        }
    }
    fir_stream<LAYOUT, PX0, false><<<(unsigned)blocks, 128, 0, st>>>(p);
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

// ------------------------------------------------------------------------------------------------
// fir_resample2<UP, PX0>: the true 2x cases of the 4x4 filter -- down2 (UP = false: D skip path, upfirdn2d.py:369-404) and
// up2 (UP = true: image upsampling, the gradient of the skip path, upfirdn2d.py:329-364) -- with the marching scheme of
// fir_stream: one thread walks down the INPUT rows of a strip, 128-bit aligned loads, running partial sums of the output rows
// that are still open, 128-bit stores.
//   down2: thread = 4 output columns (8 input columns + halo); input row j (relative to the first tap row of the strip) feeds
//          filter rows ky = j mod 2 and ky + 2 of two output rows.
//   up2  : thread = 8 output columns (4 input columns + halo); polyphase: every output has 2x2 live taps; input row r
//          completes output rows q = 2r-3, 2r-2 and opens q = 2r-1, 2r   (q = oy - pady0).
struct Res2P {
    const float* x; const float* f; float* y;
    int NC, inH, inW, padx0, pady0, flip;
    float gain;
    int outH, outW, RS, ncg, nst;               // RS = output rows per strip
};

template <bool UP, int PX0>
__global__ void __launch_bounds__(128, 4) fir_resample2(Res2P p) {
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
    const long long id = (long long)blockIdx.x * 128 + threadIdx.x;
    const int cg = (int)(id % p.ncg);
    const long long rr_ = id / p.ncg;
    const int st = (int)(rr_ % p.nst);
    const long long nc = rr_ / p.nst;
    if (nc >= p.NC) return;
    const float* xin = p.x + (size_t)nc * p.inH * p.inW;
    float* yout = p.y + (size_t)nc * p.outH * p.outW;
    const int oy0 = st * p.RS, nrows = min(p.RS, p.outH - oy0);
    const bool y_al = (reinterpret_cast<uintptr_t>(p.y) & 15) == 0 && (p.outW & 3) == 0;

    if (!UP) {
        const int ox0 = cg * 4, x0 = 2 * ox0;                   // x0: first input column of the aligned 8-column block
        // in[j] = input column x0 - PX0 + j, j = 0..9 (zero outside the image)
        auto load_row = [&](int r, float (&in)[11]) {
            const bool row_ok = (unsigned)r < (unsigned)p.inH;
            const float* rowp = xin + (size_t)(row_ok ? r : 0) * p.inW + x0;
            float buf[16];
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                if (b == 0 && PX0 == 0) continue;
                if (b == 3 && PX0 >= 2) continue;                // columns up to x0 - PX0 + 9
                const int cb = x0 - 4 + 4 * b;
                if (row_ok && cb >= 0 && cb + 3 < p.inW) {
                    const float4 v = __ldg(reinterpret_cast<const float4*>(rowp - 4 + 4 * b));
                    buf[4 * b] = v.x; buf[4 * b + 1] = v.y; buf[4 * b + 2] = v.z; buf[4 * b + 3] = v.w;
                } else {
#pragma unroll
                    for (int e = 0; e < 4; ++e)
                        buf[4 * b + e] = (row_ok && cb + e >= 0 && cb + e < p.inW) ? __ldg(rowp - 4 + 4 * b + e) : 0.f;
                }
            }
#pragma unroll
            for (int j = 0; j < 10; ++j) in[j] = buf[4 - PX0 + j];
            in[10] = 0.f;
        };
        float prev[4], cur[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) { prev[t] = 0.f; cur[t] = 0.f; }
        const int rbase = 2 * oy0 - p.pady0;                     // input row of relative tap row j = 0
        const int nt = 2 * nrows + 2;
        auto hsum = [&](const float (&in)[11], int ky, int t, float a) {
#pragma unroll
            for (int kx = 0; kx < 4; ++kx) a = fmaf(K[ky][kx], in[2 * t + kx], a);
            return a;
        };
        auto step = [&](int j, const float (&in)[11]) {
            if (!(j & 1)) {                                      // even: filter rows 2 (older output row) and 0 (opens a row)
#pragma unroll
                for (int t = 0; t < 4; ++t) { prev[t] = hsum(in, 2, t, prev[t]); cur[t] = hsum(in, 0, t, 0.f); }
            } else {                                             // odd: filter row 3 completes the older row, 1 continues the newer
                float o[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) { o[t] = hsum(in, 3, t, prev[t]); prev[t] = hsum(in, 1, t, cur[t]); }
                const int oy = oy0 + (j - 3) / 2;
                if (j >= 3) {
                    float* dst = yout + (size_t)oy * p.outW + ox0;
                    if (y_al && ox0 + 3 < p.outW) *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                    else {
#pragma unroll
                        for (int t = 0; t < 4; ++t) if (ox0 + t < p.outW) dst[t] = o[t];
                    }
                }
            }
        };
        float r0[11], r1[11], r2[11];
        load_row(rbase, r0);
        load_row(rbase + 1, r1);
        for (int j = 0; j < nt; j += 6) {                        // steps of 6 keep the parity of j compile-time inside the body
            if (j + 2 < nt) load_row(rbase + j + 2, r2);
            step(j, r0);
            if (j + 1 < nt) { if (j + 3 < nt) load_row(rbase + j + 3, r0); step(j + 1, r1); }
            if (j + 2 < nt) { if (j + 4 < nt) load_row(rbase + j + 4, r1); step(j + 2, r2); }
            if (j + 3 < nt) { if (j + 5 < nt) load_row(rbase + j + 5, r2); step(j + 3, r0); }
            if (j + 4 < nt) { if (j + 6 < nt) load_row(rbase + j + 6, r0); step(j + 4, r1); }
            if (j + 5 < nt) { if (j + 7 < nt) load_row(rbase + j + 7, r1); step(j + 5, r2); }
        }
    } else {
        const int ox0 = cg * 8, h0 = ox0 >> 1;                   // h0: first input column of the aligned 4-column block
        // in[e + 2] = input column h0 + e, e = -2..5 (zero outside the image)
        auto load_row = [&](int r, float (&in)[8]) {
            const bool row_ok = (unsigned)r < (unsigned)p.inH;
            const float* rowp = xin + (size_t)(row_ok ? r : 0) * p.inW + h0;
            if (row_ok && h0 >= 2 && h0 - 1 < p.inW) { const float2 v = __ldg(reinterpret_cast<const float2*>(rowp - 2)); in[0] = v.x; in[1] = v.y; }
            else {                                               // (h0 is a multiple of 4: h0 < 2 means columns -2, -1 = padding)
#pragma unroll
                for (int e = 0; e < 2; ++e) in[e] = (row_ok && h0 - 2 + e >= 0 && h0 - 2 + e < p.inW) ? __ldg(rowp - 2 + e) : 0.f;
            }
            if (row_ok && h0 + 3 < p.inW) { const float4 v = __ldg(reinterpret_cast<const float4*>(rowp)); in[2] = v.x; in[3] = v.y; in[4] = v.z; in[5] = v.w; }
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e) in[2 + e] = (row_ok && h0 + e < p.inW) ? __ldg(rowp + e) : 0.f;
            }
            if (row_ok && h0 + 5 < p.inW) { const float2 v = __ldg(reinterpret_cast<const float2*>(rowp + 4)); in[6] = v.x; in[7] = v.y; }
            else {
#pragma unroll
                for (int e = 0; e < 2; ++e) in[6 + e] = (row_ok && h0 + 4 + e < p.inW) ? __ldg(rowp + 4 + e) : 0.f;
            }
        };
        float podd[8], pevn[8];                                   // open output rows q = 2r-1 (filter row 1 done) and q = 2r (row 0 done)
#pragma unroll
        for (int t = 0; t < 8; ++t) { podd[t] = 0.f; pevn[t] = 0.f; }
        // horizontal polyphase sum of filter row ky for output column t: taps kx with (t + kx - PX0) even
        auto hsum = [&](const float (&in)[8], int ky, int t, float a) {
#pragma unroll
            for (int kx = 0; kx < 4; ++kx)
                if (((t + kx - PX0) & 1) == 0) a = fmaf(K[ky][kx], in[(t + kx - PX0) / 2 + 2], a);   // compile-time after unrolling
            return a;
        };
        const int q0 = oy0 - p.pady0;                            // q of the strip's first output row
        const int r_lo = (q0 + 1) >> 1, r_hi = ((q0 + nrows) >> 1) + 1;      // ceil(q0/2) .. ceil((q0+nrows-1)/2)+1  (arithmetic shift = floor)
        auto store = [&](int q, const float (&o)[8]) {
            const int oy = q + p.pady0;
            if (oy < oy0 || oy >= oy0 + nrows) return;
            float* dst = yout + (size_t)oy * p.outW + ox0;
            if (y_al && ox0 + 7 < p.outW) {
                *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
                *reinterpret_cast<float4*>(dst + 4) = make_float4(o[4], o[5], o[6], o[7]);
            } else {
#pragma unroll
                for (int t = 0; t < 8; ++t) if (ox0 + t < p.outW) dst[t] = o[t];
            }
        };
        auto step = [&](int r, const float (&in)[8]) {
            float o3[8], o2[8];
#pragma unroll
            for (int t = 0; t < 8; ++t) {
                o3[t] = hsum(in, 3, t, podd[t]);                 // q = 2r-3: filter row 1 came from input row r-1
                o2[t] = hsum(in, 2, t, pevn[t]);                 // q = 2r-2: filter row 0 came from input row r-1
                podd[t] = hsum(in, 1, t, 0.f);                   // opens q = 2r-1
                pevn[t] = hsum(in, 0, t, 0.f);                   // opens q = 2r
            }
            store(2 * r - 3, o3);
            store(2 * r - 2, o2);
        };
        float a0[8], a1[8], a2[8];
        load_row(r_lo, a0);
        load_row(r_lo + 1, a1);
        for (int r = r_lo; r <= r_hi; r += 3) {
            if (r + 2 <= r_hi) load_row(r + 2, a2);
            step(r, a0);
            if (r + 1 <= r_hi) { if (r + 3 <= r_hi) load_row(r + 3, a0); step(r + 1, a1); }
            if (r + 2 <= r_hi) { if (r + 4 <= r_hi) load_row(r + 4, a1); step(r + 2, a2); }
        }
    }
}

template <bool UP>
int launch_resample2(Res2P p, cudaStream_t st) {
    p.RS = p.outH >= 128 ? 32 : (p.outH >= 32 ? 16 : 8);
    p.ncg = UP ? (p.outW + 7) / 8 : (p.outW + 3) / 4;
    p.nst = (p.outH + p.RS - 1) / p.RS;
    const long long threads = (long long)p.ncg * p.nst * p.NC;
    const long long blocks = (threads + 127) / 128;
    if (blocks > 0x7fffffffLL) { gg::set_error("upfirdn2d: grid too large"); return GG_EINVAL; }
    switch (p.padx0) {
        case 0: fir_resample2<UP, 0><<<(unsigned)blocks, 128, 0, st>>>(p); break;
        case 1: fir_resample2<UP, 1><<<(unsigned)blocks, 128, 0, st>>>(p); break;
        case 2: fir_resample2<UP, 2><<<(unsigned)blocks, 128, 0, st>>>(p); break;
        default: fir_resample2<UP, 3><<<(unsigned)blocks, 128, 0, st>>>(p); break;
    }
    return gg::check_launch("upfirdn2d(fir_resample2)");
}

// padx0 must be one of the four instantiations; phase-major input rows must be 16-byte aligned (plain rows that are not take the
// scalar loads of fir_march: the odd-width maps behind an up-sampling convolution, 17^2 .. 1025^2)
bool stream_ok(const StreamP& p, bool in_pm) {
    if (p.padx0 < 0 || p.padx0 > 3) return false;
    return !in_pm || (p.ipW % 4 == 0 && (reinterpret_cast<uintptr_t>(p.x) & 15) == 0);
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

    // 4x4 filter, isotropic factors.  The marching kernels serve every size from 8 output columns up (column groups of neighbouring
    // strips / planes fill the warps of narrow images; the cfg 5 sweep had the generic kernel at half the speed of the reference's
    // SIMT plugin on the 16^2 and 32^2 maps); the shared-memory tile kernels below them want >= 48 columns.
    const bool f4m = (fH == 4 && fW == 4) && upx == upy && downx == downy && outW >= 8;
    const bool f4 = f4m && outW >= 48;
    if (f4m && upx == 1 && downx == 1) {   // unit rate: the register-streaming kernel (also serves the phase-major layouts)
        StreamP sp{x, f, y, N, C, inH, inW, padx0, pady0, flip, gain, outH, outW, 0, 0, 0, 0, 0, 0, 0, 0, 0};
        if (stream_ok(sp, false)) return launch_stream<0>(sp, st);
        if (f4) return launch_tile<1, 1, 0, 4, 4>(p, st);
    }
    // true 2x resampling: the marching kernels need 16-byte aligned input rows, 0 <= pad0 <= 3 and (up2) no pad0 beyond the filter reach
    const bool march_ok = f4m && padx0 >= 0 && padx0 <= 3 && pady0 >= 0 && pady0 <= 3 && inW % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
    if (march_ok && upx == 1 && downx == 2) {
        Res2P rp{x, f, y, N * C, inH, inW, padx0, pady0, flip, gain, outH, outW, 0, 0, 0};
        return launch_resample2<false>(rp, st);
    }
    if (march_ok && upx == 2 && downx == 1) {
        Res2P rp{x, f, y, N * C, inH, inW, padx0, pady0, flip, gain, outH, outW, 0, 0, 0};
        return launch_resample2<true>(rp, st);
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
    StreamP p{x, f, y, N, C, inH, inW, padx0, pady0, flip, gain, outH, outW, in_pmH, in_pmW, out_pmH, out_pmW, 0, 0, 0, 0, 0};
    cudaStream_t st = (cudaStream_t)stream;
    GG_REQUIRE(stream_ok(p, in_pm != 0), "fir4_pm: needs 0 <= padx0 <= 3 and, for a phase-major input, 16-byte aligned rows (width %% 4 == 0)");
    if (in_pm) return launch_stream<1>(p, st);
    if (out_pm) return launch_stream<2>(p, st);
    return launch_stream<0>(p, st);
}
