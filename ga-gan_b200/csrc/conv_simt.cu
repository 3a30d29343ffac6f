// Exact-fp32 (FFMA) implicit-GEMM convolution kernels for sm_100a.
//
// These serve the shapes the tcgen05 path does not take (4x4/8x8 maps, 3- or 513-channel inputs,
// strided / transposed layers until their tensor-core variants land) and are the on-GPU fp32
// cross-check of the tensor-core kernels.  They replace the ATen/cuDNN calls behind
// conv2d_gradfix.conv2d / conv_transpose2d / cudnn_convolution_backward_weight
// (torch_utils/ops/conv2d_gradfix.py:37-58,138-148,175-191) with the modulated_conv2d per-sample
// scales (training/networks.py:642,648-651) folded into the operand loads and the epilogue.
//
// conv_fwd_simt : GEMM  M = N*OH*OW pixels, N = O, K = I*KH*KW.  128x64 CTA tile, 8x4 register tile,
//                 register-prefetched double-buffered shared memory, gather addressing covers
//                 stride, zero padding, transposed (fractionally strided) convolution and weight flip.
// wgrad_simt    : GEMM  M = B (grad channels), N = A*KH*KW, K = N*HB*WB pixels; 64x64 CTA tile,
//                 split-K across CTAs with fp32 atomics when the M x N grid alone cannot fill 148 SMs.
#include "common.cuh"

namespace {

struct ConvP {
    const float* x; const float* w; float* y;
    const float* in_scale; const float* out_scale;
    int N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x, transposed, flip_w;
    int64_t M;   // N*OH*OW
    int K;       // I*KH*KW
};

constexpr int BM = 128, BN = 64, BK = 8;

__global__ void __launch_bounds__(256) conv_fwd_simt(ConvP p) {
    __shared__ __align__(16) float As[2][BK][BM];
    __shared__ __align__(16) float Bs[2][BK][BN];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    // A-operand gather: this thread always fetches pixel (m0 + tid%128), k-rows tid/128 + 2j
    const int a_mm = tid & 127, a_k0 = tid >> 7;
    const int64_t a_m = m0 + a_mm;
    const bool a_ok = a_m < p.M;
    int a_n = 0, a_oy = 0, a_ox = 0;
    if (a_ok) { a_ox = (int)(a_m % p.OW); int64_t r = a_m / p.OW; a_oy = (int)(r % p.OH); a_n = (int)(r / p.OH); }
    const float* xn = p.x + (int64_t)a_n * p.I * p.H * p.W;
    const float* sn = p.in_scale ? p.in_scale + (int64_t)a_n * p.I : nullptr;
    // B-operand: weight element (o = n0 + tid%64, k-row tid/64 + 4j)
    const int b_nn = tid & 63, b_k0 = tid >> 6;
    const int b_o = n0 + b_nn;
    const int khw = p.KH * p.KW;

    float ra[4], rb[2];
    auto load = [&](int kbase) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int k = kbase + a_k0 + 2 * j;
            float v = 0.f;
            if (a_ok && k < p.K) {
                int i = k / khw, t = k - i * khw, ky = t / p.KW, kx = t - ky * p.KW;
                int iy, ix; bool ok;
                if (!p.transposed) {
                    iy = a_oy * p.stride - p.pad_y + ky; ix = a_ox * p.stride - p.pad_x + kx;
                    ok = iy >= 0 && iy < p.H && ix >= 0 && ix < p.W;
                } else {
                    int ty = a_oy + p.pad_y - ky, tx = a_ox + p.pad_x - kx;
                    ok = ty >= 0 && tx >= 0 && (ty % p.stride) == 0 && (tx % p.stride) == 0;
                    iy = ty / p.stride; ix = tx / p.stride;
                    ok = ok && iy < p.H && ix < p.W;
                }
                if (ok) {
                    v = __ldg(xn + ((int64_t)i * p.H + iy) * p.W + ix);
                    if (sn) v *= __ldg(sn + i);
                }
            }
            ra[j] = v;
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            int k = kbase + b_k0 + 4 * j;
            float v = 0.f;
            if (b_o < p.O && k < p.K) {
                int i = k / khw, t = k - i * khw, ky = t / p.KW, kx = t - ky * p.KW;
                if (p.flip_w) { ky = p.KH - 1 - ky; kx = p.KW - 1 - kx; }
                int64_t idx = p.transposed ? (((int64_t)i * p.O + b_o) * p.KH + ky) * p.KW + kx
                                           : (((int64_t)b_o * p.I + i) * p.KH + ky) * p.KW + kx;
                v = __ldg(p.w + idx);
            }
            rb[j] = v;
        }
    };
    auto stash = [&](int buf) {
#pragma unroll
        for (int j = 0; j < 4; ++j) As[buf][a_k0 + 2 * j][a_mm] = ra[j];
#pragma unroll
        for (int j = 0; j < 2; ++j) Bs[buf][b_k0 + 4 * j][b_nn] = rb[j];
    };

    const int tm = tid & 15, tn = tid >> 4;   // 16 x 16 threads; thread tile: m = tm*4 + 64*h + jj, n = tn*4 + c
    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    load(0);
    stash(0);
    __syncthreads();
    for (int kc = 0; kc < nk; ++kc) {
        const int cur = kc & 1;
        if (kc + 1 < nk) load((kc + 1) * BK);
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][tm * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + tm * 4]);
            float4 b = *reinterpret_cast<const float4*>(&Bs[cur][kk][tn * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int a = 0; a < 8; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[a][c] = fmaf(av[a], bv[c], acc[a][c]);
        }
        if (kc + 1 < nk) stash(cur ^ 1);
        __syncthreads();
    }

    // epilogue: out_scale, NCHW store (4 consecutive pixels per thread and channel)
    const int64_t plane = (int64_t)p.OH * p.OW;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const int64_t mb = m0 + h * 64 + tm * 4;
        if (mb >= p.M) continue;
        const int64_t nimg = mb / plane, rem = mb - nimg * plane;
        const bool row4 = (mb + 3 < p.M) && (rem + 3 < plane) && ((plane & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.y) & 15) == 0);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int o = n0 + tn * 4 + c;
            if (o >= p.O) continue;
            if (row4) {
                float s = p.out_scale ? __ldg(p.out_scale + nimg * p.O + o) : 1.f;
                float4 v = make_float4(acc[h * 4 + 0][c] * s, acc[h * 4 + 1][c] * s, acc[h * 4 + 2][c] * s, acc[h * 4 + 3][c] * s);
                *reinterpret_cast<float4*>(p.y + (nimg * p.O + o) * plane + rem) = v;
            } else {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    int64_t m = mb + jj;
                    if (m >= p.M) continue;
                    int64_t ni = m / plane, rr = m - ni * plane;
                    float s = p.out_scale ? __ldg(p.out_scale + ni * p.O + o) : 1.f;
                    p.y[(ni * p.O + o) * plane + rr] = acc[h * 4 + jj][c] * s;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
struct WgradP {
    const float* a; const float* b; float* dw;
    const float* a_scale; const float* b_scale;
    int N, A, HA, WA, B, HB, WB, KH, KW, stride, pad_y, pad_x, flip_w, out_layout;
    int J;          // A*KH*KW
    int64_t KP;     // N*HB*WB
    int64_t kchunk; // pixels per split
    int atomic;
};

constexpr int WM = 64, WN = 64, WK = 16, WPITCH = 68;

__global__ void __launch_bounds__(256) wgrad_simt(WgradP p) {
    __shared__ __align__(16) float Bt[WK][WPITCH];   // [pixel][grad channel]
    __shared__ __align__(16) float At[WK][WPITCH];   // [pixel][input channel x tap]
    const int tid = threadIdx.x;
    const int b0 = blockIdx.x * WM, j0 = blockIdx.y * WN;
    const int64_t kbeg = (int64_t)blockIdx.z * p.kchunk;
    const int64_t kend = min(p.KP, kbeg + p.kchunk);
    const int lk = tid & 15, lc = tid >> 4;   // loader: pixel lk, columns lc + 16*j
    const int tb = tid & 15, tj = tid >> 4;   // compute: rows tb*4.., cols tj*4..
    const int khw = p.KH * p.KW;
    const int64_t planeB = (int64_t)p.HB * p.WB, planeA = (int64_t)p.HA * p.WA;

    // decode this thread's 4 (a, ky, kx) columns once
    int ca[4], cky[4], ckx[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        int col = j0 + lc + 16 * j;
        if (col < p.J) { ca[j] = col / khw; int t = col - ca[j] * khw; cky[j] = t / p.KW; ckx[j] = t - cky[j] * p.KW; }
        else { ca[j] = -1; cky[j] = 0; ckx[j] = 0; }
    }

    float acc[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[r][c] = 0.f;

    for (int64_t k0 = kbeg; k0 < kend; k0 += WK) {
        const int64_t pix = k0 + lk;
        const bool pok = pix < kend;
        int n = 0, oy = 0, ox = 0;
        if (pok) { ox = (int)(pix % p.WB); int64_t r = pix / p.WB; oy = (int)(r % p.HB); n = (int)(r / p.HB); }
        float vb[4], va[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int bc = b0 + lc + 16 * j;
            float v = 0.f;
            if (pok && bc < p.B) {
                v = __ldg(p.b + ((int64_t)n * p.B + bc) * planeB + (int64_t)oy * p.WB + ox);
                if (p.b_scale) v *= __ldg(p.b_scale + (int64_t)n * p.B + bc);
            }
            vb[j] = v;
            float u = 0.f;
            if (pok && ca[j] >= 0) {
                int iy = oy * p.stride - p.pad_y + cky[j], ix = ox * p.stride - p.pad_x + ckx[j];
                if (iy >= 0 && iy < p.HA && ix >= 0 && ix < p.WA) {
                    u = __ldg(p.a + ((int64_t)n * p.A + ca[j]) * planeA + (int64_t)iy * p.WA + ix);
                    if (p.a_scale) u *= __ldg(p.a_scale + (int64_t)n * p.A + ca[j]);
                }
            }
            va[j] = u;
        }
        __syncthreads();   // previous tile fully consumed
#pragma unroll
        for (int j = 0; j < 4; ++j) { Bt[lk][lc + 16 * j] = vb[j]; At[lk][lc + 16 * j] = va[j]; }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < WK; ++kk) {
            float4 b4 = *reinterpret_cast<const float4*>(&Bt[kk][tb * 4]);
            float4 a4 = *reinterpret_cast<const float4*>(&At[kk][tj * 4]);
            float bv[4] = {b4.x, b4.y, b4.z, b4.w}, av[4] = {a4.x, a4.y, a4.z, a4.w};
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[r][c] = fmaf(bv[r], av[c], acc[r][c]);
        }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        int bc = b0 + tb * 4 + r;
        if (bc >= p.B) continue;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int col = j0 + tj * 4 + c;
            if (col >= p.J) continue;
            int a = col / khw, t = col - a * khw, ky = t / p.KW, kx = t - ky * p.KW;
            if (p.flip_w) { ky = p.KH - 1 - ky; kx = p.KW - 1 - kx; }
            int64_t idx = p.out_layout ? (((int64_t)a * p.B + bc) * p.KH + ky) * p.KW + kx
                                       : (((int64_t)bc * p.A + a) * p.KH + ky) * p.KW + kx;
            if (p.atomic) atomicAdd(p.dw + idx, acc[r][c]); else p.dw[idx] = acc[r][c];
        }
    }
}

}  // namespace

namespace gg {

int conv2d_simt(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW,
                int stride, int pad_y, int pad_x, int transposed, int flip_w, const float* in_scale, const float* out_scale,
                cudaStream_t st) {
    ConvP p{x, w, y, in_scale, out_scale, N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x, transposed, flip_w,
            (int64_t)N * OH * OW, I * KH * KW};
    if (p.M == 0 || O == 0) return GG_OK;
    if (p.K == 0) { cudaMemsetAsync(y, 0, sizeof(float) * (size_t)p.M * O, st); return GG_OK; }
    dim3 grid((unsigned)((p.M + BM - 1) / BM), (unsigned)((O + BN - 1) / BN));
    conv_fwd_simt<<<grid, 256, 0, st>>>(p);
    return check_launch("conv2d(simt)");
}

int conv2d_wgrad_simt(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int KH,
                      int KW, int stride, int pad_y, int pad_x, int flip_w, int out_layout, const float* a_scale,
                      const float* b_scale, cudaStream_t st) {
    WgradP p{a, b, dw, a_scale, b_scale, N, A, HA, WA, B, HB, WB, KH, KW, stride, pad_y, pad_x, flip_w, out_layout,
             A * KH * KW, (int64_t)N * HB * WB, 0, 0};
    const size_t bytes = sizeof(float) * (size_t)A * B * KH * KW;
    if (bytes == 0) return GG_OK;
    if (p.KP == 0) { cudaMemsetAsync(dw, 0, bytes, st); return GG_OK; }
    int gm = (B + WM - 1) / WM, gn = (p.J + WN - 1) / WN;
    int64_t want = (4LL * GG_NUM_SMS + (int64_t)gm * gn - 1) / ((int64_t)gm * gn);   // ~4 waves of CTAs
    int64_t maxsplit = (p.KP + 255) / 256;
    int64_t splits = want < 1 ? 1 : (want > maxsplit ? maxsplit : want);
    if (splits > 65535) splits = 65535;
    p.kchunk = ((p.KP + splits - 1) / splits + WK - 1) / WK * WK;
    splits = (p.KP + p.kchunk - 1) / p.kchunk;
    p.atomic = splits > 1;
    if (p.atomic) {
        cudaError_t e = cudaMemsetAsync(dw, 0, bytes, st);
        if (e != cudaSuccess) { set_error("wgrad memset: %s", cudaGetErrorString(e)); return GG_ECUDA; }
    }
    dim3 grid(gm, gn, (unsigned)splits);
    wgrad_simt<<<grid, 256, 0, st>>>(p);
    return check_launch("conv2d_wgrad(simt)");
}

}  // namespace gg
