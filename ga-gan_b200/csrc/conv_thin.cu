// 1x1 convolutions with a thin channel side (<= 4 channels in or out): ToRGB (C -> 3), fromRGB (3 -> C), their data
// gradients and weight gradients.  Arithmetic intensity ~1.4 FLOP/B: these are HBM-streaming kernels, not GEMMs
// (SURVEY.md section 8(d)); exact fp32 FFMA, 128-bit loads/stores along the pixel axis, every input element read once.
//
// Replaces the ATen conv calls of conv2d_gradfix.py:141-146,178-188 for the 1x1 layers of training/networks.py:957-963
// (ToRGBLayer, modulated: `in_scale` = styles) and :1254-1258 (fromrgb Conv2dLayer).
//
//   fwd   y[n,o,p]  = os[n,o] * sum_i W(o,i) * is[n,i] * x[n,i,p]          W(o,i) = w[o,i] or w[i,o] (transposed layout)
//   wgrad dw[b,a]   = sum_{n,p} gs[n,b] G[n,b,p] * xs[n,a] X[n,a,p]
#include "common.cuh"

namespace {

constexpr int THIN = 4;          // channel bound of the thin side
constexpr int MAXC = 512;        // channel bound of the wide side (weights of one image live in shared memory)

struct ThinP {
    const float* x; const float* w; float* y; const float* is; const float* os;
    int N, I, O, w_io;
    int64_t P;                   // pixels per plane
};

// thin OUTPUT (O <= 4): one thread = 4 pixels x all outputs, loop over the wide input channels
__global__ void __launch_bounds__(256) conv1x1_thin_out(ThinP p) {
    __shared__ float sw[THIN][MAXC];
    const int n = blockIdx.y;
    for (int idx = threadIdx.x; idx < p.O * p.I; idx += blockDim.x) {
        const int o = idx / p.I, i = idx - o * p.I;
        float v = __ldg(p.w + (p.w_io ? (int64_t)i * p.O + o : (int64_t)o * p.I + i));
        if (p.is) v *= __ldg(p.is + (int64_t)n * p.I + i);
        if (p.os) v *= __ldg(p.os + (int64_t)n * p.O + o);
        sw[o][i] = v;
    }
    __syncthreads();
    const int64_t P4 = p.P >> 2;
    const float4* xn = reinterpret_cast<const float4*>(p.x + (int64_t)n * p.I * p.P);
    float4* yn = reinterpret_cast<float4*>(p.y + (int64_t)n * p.O * p.P);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < P4; q += (int64_t)gridDim.x * blockDim.x) {
        float4 acc[THIN];
#pragma unroll
        for (int o = 0; o < THIN; ++o) acc[o] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int i = 0; i < p.I; ++i) {
            const float4 v = __ldg(xn + (int64_t)i * P4 + q);
#pragma unroll
            for (int o = 0; o < THIN; ++o) {
                const float wv = sw[o][i];
                acc[o].x = fmaf(wv, v.x, acc[o].x); acc[o].y = fmaf(wv, v.y, acc[o].y);
                acc[o].z = fmaf(wv, v.z, acc[o].z); acc[o].w = fmaf(wv, v.w, acc[o].w);
            }
        }
#pragma unroll
        for (int o = 0; o < THIN; ++o)
            if (o < p.O) yn[(int64_t)o * P4 + q] = acc[o];
    }
}

// thin INPUT (I <= 4): one thread = 4 pixels, reads <= 4 planes once, streams out the wide output channels
__global__ void __launch_bounds__(256) conv1x1_thin_in(ThinP p) {
    __shared__ float sw[MAXC][THIN];
    const int n = blockIdx.y;
    for (int idx = threadIdx.x; idx < p.O * THIN; idx += blockDim.x) {
        const int o = idx / THIN, i = idx - o * THIN;
        float v = 0.f;
        if (i < p.I) {
            v = __ldg(p.w + (p.w_io ? (int64_t)i * p.O + o : (int64_t)o * p.I + i));
            if (p.is) v *= __ldg(p.is + (int64_t)n * p.I + i);
            if (p.os) v *= __ldg(p.os + (int64_t)n * p.O + o);
        }
        sw[o][i] = v;
    }
    __syncthreads();
    const int64_t P4 = p.P >> 2;
    const float4* xn = reinterpret_cast<const float4*>(p.x + (int64_t)n * p.I * p.P);
    float4* yn = reinterpret_cast<float4*>(p.y + (int64_t)n * p.O * p.P);
    for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q < P4; q += (int64_t)gridDim.x * blockDim.x) {
        float4 v[THIN];
#pragma unroll
        for (int i = 0; i < THIN; ++i) v[i] = (i < p.I) ? __ldg(xn + (int64_t)i * P4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
        for (int o = 0; o < p.O; ++o) {
            const float4 wv = *reinterpret_cast<const float4*>(&sw[o][0]);
            float4 r;
            r.x = fmaf(wv.w, v[3].x, fmaf(wv.z, v[2].x, fmaf(wv.y, v[1].x, wv.x * v[0].x)));
            r.y = fmaf(wv.w, v[3].y, fmaf(wv.z, v[2].y, fmaf(wv.y, v[1].y, wv.x * v[0].y)));
            r.z = fmaf(wv.w, v[3].z, fmaf(wv.z, v[2].z, fmaf(wv.y, v[1].z, wv.x * v[0].z)));
            r.w = fmaf(wv.w, v[3].w, fmaf(wv.z, v[2].w, fmaf(wv.y, v[1].w, wv.x * v[0].w)));
            yn[(int64_t)o * P4 + q] = r;
        }
    }
}

// weight gradient with a thin side: T = the thin tensor (<= 4 channels), Wd = the wide one.
//   part[c][t] = sum_p Wd[n,c,p] * T[n,t,p];   thin_is_a = 1: T is the conv input (a), Wd the output gradient (b)
struct ThinWgP {
    const float* T; const float* Wd; float* dw; const float* ts; const float* ws;
    int N, CT, CW, thin_is_a, out_layout;
    int64_t P;
    int chunks;                  // pixel chunks per image (grid.x)
};

constexpr int WCH = 16;          // wide channels per pass: THIN x WCH partial sums live in registers

__global__ void __launch_bounds__(256) wgrad1x1_thin(ThinWgP p) {
    __shared__ float red[8][WCH][THIN];
    const int n = blockIdx.y;
    const int64_t P4 = p.P >> 2;
    const int64_t per = (P4 + p.chunks - 1) / p.chunks;
    const int64_t q0 = (int64_t)blockIdx.x * per, q1 = min(P4, q0 + per);
    const float4* tn = reinterpret_cast<const float4*>(p.T + (int64_t)n * p.CT * p.P);
    const float4* wn = reinterpret_cast<const float4*>(p.Wd + (int64_t)n * p.CW * p.P);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // The thin planes of a pixel stay in registers while WCH wide channels stream by; the wide tensor is read exactly once,
    // the thin planes CW/WCH times (from L2).
    {
        const int c0 = blockIdx.z * WCH;             // one pass of WCH wide channels per CTA (grid.z)
        float acc[WCH][THIN];
#pragma unroll
        for (int c = 0; c < WCH; ++c)
#pragma unroll
            for (int t = 0; t < THIN; ++t) acc[c][t] = 0.f;
        for (int64_t q = q0 + threadIdx.x; q < q1; q += blockDim.x) {
            float4 v[THIN];
#pragma unroll
            for (int t = 0; t < THIN; ++t) v[t] = (t < p.CT) ? __ldg(tn + (int64_t)t * P4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int c = 0; c < WCH; ++c) {
                if (c0 + c < p.CW) {
                    const float4 g = __ldg(wn + (int64_t)(c0 + c) * P4 + q);
#pragma unroll
                    for (int t = 0; t < THIN; ++t)
                        acc[c][t] = fmaf(g.x, v[t].x, fmaf(g.y, v[t].y, fmaf(g.z, v[t].z, fmaf(g.w, v[t].w, acc[c][t]))));
                }
            }
        }
#pragma unroll
        for (int c = 0; c < WCH; ++c)
#pragma unroll
            for (int t = 0; t < THIN; ++t) {
                const float s = gg::warp_sum(acc[c][t]);
                if (lane == 0) red[warp][c][t] = s;
            }
        __syncthreads();
        if (threadIdx.x < WCH * THIN) {
            const int c = threadIdx.x / THIN, t = threadIdx.x - c * THIN;
            if (c0 + c < p.CW && t < p.CT) {
                float s = 0.f;
                for (int w8 = 0; w8 < 8; ++w8) s += red[w8][c][t];
                if (p.ws) s *= __ldg(p.ws + (int64_t)n * p.CW + c0 + c);
                if (p.ts) s *= __ldg(p.ts + (int64_t)n * p.CT + t);
                // dw is [B,A] (out_layout 0) or [A,B] (1); a = conv input channel, b = gradient channel
                const int a = p.thin_is_a ? t : c0 + c, b = p.thin_is_a ? c0 + c : t;
                const int A = p.thin_is_a ? p.CT : p.CW, B = p.thin_is_a ? p.CW : p.CT;
                atomicAdd(p.dw + (p.out_layout ? (int64_t)a * B + b : (int64_t)b * A + a), s);
            }
        }
    }
}

}  // namespace

namespace gg {

bool conv1x1_thin_eligible(const float* x, const float* y, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW,
                           int stride, int pad_y, int pad_x) {
    if (KH != 1 || KW != 1 || stride != 1 || pad_y != 0 || pad_x != 0 || OH != H || OW != W) return false;
    if (!((I <= THIN && O <= MAXC) || (O <= THIN && I <= MAXC))) return false;
    if (((int64_t)H * W) % 4 != 0 || N < 1 || N > 65535) return false;
    return ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
}

int conv1x1_thin(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int w_io, const float* in_scale,
                 const float* out_scale, cudaStream_t st) {
    ThinP p{x, w, y, in_scale, out_scale, N, I, O, w_io, (int64_t)H * W};
    const int64_t P4 = p.P >> 2;
    int gx = (int)((P4 + 255) / 256);
    const int cap = (GG_NUM_SMS * 8 + N - 1) / N;        // ~8 CTAs per SM in total, grid-stride over the pixels
    if (gx > cap) gx = cap;
    if (gx < 1) gx = 1;
    dim3 grid(gx, N);
    if (O <= THIN) conv1x1_thin_out<<<grid, 256, 0, st>>>(p);
    else           conv1x1_thin_in<<<grid, 256, 0, st>>>(p);
    return check_launch("conv2d(1x1 thin)");
}

bool wgrad1x1_thin_eligible(const float* a, const float* b, int N, int A, int HA, int WA, int B, int HB, int WB, int KH, int KW,
                            int stride, int pad_y, int pad_x) {
    if (KH != 1 || KW != 1 || stride != 1 || pad_y != 0 || pad_x != 0 || HA != HB || WA != WB) return false;
    if (!(A <= THIN || B <= THIN)) return false;
    if (((int64_t)HA * WA) % 4 != 0 || N < 1 || N > 65535) return false;
    return ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0;
}

int wgrad1x1_thin(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int out_layout,
                  const float* a_scale, const float* b_scale, cudaStream_t st) {
    const bool thin_is_a = A <= THIN;
    ThinWgP p{thin_is_a ? a : b, thin_is_a ? b : a, dw, thin_is_a ? a_scale : b_scale, thin_is_a ? b_scale : a_scale,
              N, thin_is_a ? A : B, thin_is_a ? B : A, thin_is_a ? 1 : 0, out_layout, (int64_t)HA * WA, 1};
    const int64_t P4 = p.P >> 2;
    int chunks = (int)((P4 + 4095) / 4096);              // >= 16 float4 per thread and pass
    const int cap = (GG_NUM_SMS * 8 + N - 1) / N;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    p.chunks = chunks;
    GG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * (size_t)A * B, st));
    dim3 grid(chunks, N, (p.CW + WCH - 1) / WCH);
    wgrad1x1_thin<<<grid, 256, 0, st>>>(p);
    return check_launch("conv2d_wgrad(1x1 thin)");
}

}  // namespace gg
