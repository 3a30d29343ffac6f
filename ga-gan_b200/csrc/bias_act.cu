// bias_act for sm_100a: y = clamp(act(x + b) * gain) and its first/second-order gradient kernels.
//
// Replaces torch_utils/ops/bias_act.cu:23-147 + bias_act.cpp:32-90 of the reference.  HBM-bound
// (8 B/elem forward, 12 B/elem backward), so the design is: 128-bit loads/stores, one integer
// division per FOUR elements (the reference does a div+mod per element), the bias hoisted per
// warp when a warp's 128-element span lies in one row, and the bias gradient (the reference's
// separate `dx.sum()` pass, bias_act.py:211-212) fused in as a warp-shuffle reduction + one
// atomicAdd per warp per row.
#include "common.cuh"

namespace {

struct Params {
    const float* x;
    const float* b;
    const float* xref;
    const float* yref;
    const float* dy;
    float* y;
    float* dbias;
    float alpha, gain, clamp;
    int64_t sizeX;
    int sizeB;
    int64_t stepB;
    const float* noise;      // optional [Nn, stepB] plane(s) added before the activation (forward only): the per-pixel noise of
    int64_t noise_bs;        // SynthesisLayer (networks.py:904-917), broadcast over channels; batch stride 0 = one shared plane
};

// One element.  A = the reference's cuda_idx, G = grad order.  Mirrors the case table of
// bias_act.cu:56-142 (same strict / non-strict comparisons) with accurate libm calls.
template <int A, int G>
__device__ __forceinline__ float eval(float x, float b, float xref, float yref, float dy, const Params& p) {
    const float gain = p.gain, alpha = p.alpha, clamp = p.clamp;
    const float expRange = 80.f, halfExpRange = 40.f;
    const float seluScale = 1.0507009873554804934193349852946f;
    const float seluAlpha = 1.6732632423543772848170429916717f;
    float yy = (gain != 0.f) ? yref / gain : 0.f;
    float y = 0.f;
    if (G == 0) x += b; else xref += b;

    if (A == 1) { if (G == 0 || G == 1) y = x; }
    if (A == 2) { if (G == 0) y = (x > 0.f) ? x : 0.f; if (G == 1) y = (yy > 0.f) ? x : 0.f; }
    if (A == 3) { if (G == 0) y = (x > 0.f) ? x : x * alpha; if (G == 1) y = (yy > 0.f) ? x : x * alpha; }
    if (A == 4) {
        if (G == 0) y = tanhf(x);
        if (G == 1) y = x * (1.f - yy * yy);
        if (G == 2) y = x * (1.f - yy * yy) * (-2.f * yy);
    }
    if (A == 5) {
        if (G == 0) y = (x < -expRange) ? 0.f : 1.f / (expf(-x) + 1.f);
        if (G == 1) y = x * yy * (1.f - yy);
        if (G == 2) y = x * yy * (1.f - yy) * (1.f - 2.f * yy);
    }
    if (A == 6) {
        if (G == 0) y = (x >= 0.f) ? x : expm1f(x);
        if (G == 1) y = (yy >= 0.f) ? x : x * (yy + 1.f);
        if (G == 2) y = (yy >= 0.f) ? 0.f : x * (yy + 1.f);
    }
    if (A == 7) {
        if (G == 0) y = (x >= 0.f) ? seluScale * x : (seluScale * seluAlpha) * expm1f(x);
        if (G == 1) y = (yy >= 0.f) ? x * seluScale : x * (yy + seluScale * seluAlpha);
        if (G == 2) y = (yy >= 0.f) ? 0.f : x * (yy + seluScale * seluAlpha);
    }
    if (A == 8) {
        if (G == 0) y = (x > expRange) ? x : log1pf(expf(x));
        if (G == 1) y = x * (1.f - expf(-yy));
        if (G == 2) { float c = expf(-yy); y = x * c * (1.f - c); }
    }
    if (A == 9) {
        if (G == 0) {
            y = (x < -expRange) ? 0.f : x / (expf(-x) + 1.f);
        } else {
            float c = expf(xref);
            float d = c + 1.f;
            if (G == 1) y = (xref > halfExpRange) ? x : x * c * (xref + d) / (d * d);
            else        y = (xref > halfExpRange) ? 0.f : x * c * (xref * (2.f - d) + 2.f * d) / (d * d * d);
            yref = (xref < -expRange) ? 0.f : xref / (expf(-xref) + 1.f) * gain;
        }
    }
    y *= gain * dy;
    if (clamp >= 0.f) {
        if (G == 0) y = (y > -clamp && y < clamp) ? y : (y >= 0.f) ? clamp : -clamp;
        else        y = (yref > -clamp && yref < clamp) ? y : 0.f;
    }
    return y;
}

__device__ __forceinline__ float4 ld4(const float* p, int64_t i4) { return __ldg(reinterpret_cast<const float4*>(p) + i4); }

// Vector kernel: stepB % 4 == 0, sizeX % 4 == 0, all pointers 16-byte aligned.  Each warp owns
// ITER consecutive 128-element spans, so that when UNIFORM (stepB % 128 == 0) the bias channel
// is warp-uniform per span.
template <int A, int G, bool UNIFORM>
__global__ void __launch_bounds__(256) bias_act_vec4(Params p) {
    constexpr int ITER = 8;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t n4 = p.sizeX >> 2;
    const int64_t base4 = warp * (32 * ITER) + lane;
    float acc = 0.f;     // partial dbias for channel acc_c (UNIFORM only)
    int acc_c = -1;

    // sizeX <= INT_MAX (checked by the entry point): all element / plane arithmetic fits 32-bit unsigned, whose divisions cost a
    // fraction of the 64-bit ones (two of them per 128-element span were a measurable share of this memory-bound kernel)
    const uint32_t stepB = (uint32_t)p.stepB, sizeB = (uint32_t)p.sizeB;
    float4 vx[ITER], vr[ITER], vy[ITER], vd[ITER], vn[ITER];
#pragma unroll
    for (int it = 0; it < ITER; ++it) {           // issue all loads first (memory-level parallelism)
        int64_t i4 = base4 + it * 32;
        bool ok = i4 < n4;
        vx[it] = ok ? ld4(p.x, i4) : make_float4(0, 0, 0, 0);
        if (G == 0) {
            vn[it] = make_float4(0, 0, 0, 0);
            if (p.noise && ok) {                        // element e = 4*i4 sits at pixel e % stepB of sample e / (stepB*sizeB)
                const uint32_t e = (uint32_t)(i4 << 2), plane = e / stepB;
                vn[it] = __ldg(reinterpret_cast<const float4*>(p.noise + (int64_t)(plane / sizeB) * p.noise_bs + (e - plane * stepB)));
            }
        }
        if (G > 0 && A == 9) vr[it] = (ok && p.xref) ? ld4(p.xref, i4) : make_float4(0, 0, 0, 0);
        if (G > 0) vy[it] = (ok && p.yref) ? ld4(p.yref, i4) : make_float4(0, 0, 0, 0);
        if (G == 2) vd[it] = (ok && p.dy) ? ld4(p.dy, i4) : make_float4(1, 1, 1, 1);
    }
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        int64_t i4 = base4 + it * 32;
        bool ok = i4 < n4;
        int c = 0;
        float b = 0.f;
        if (p.b || p.dbias) {
            if (UNIFORM) {
                const uint32_t span0 = (uint32_t)((warp * (32 * ITER) + it * 32) << 2);  // first element of this warp-span
                c = (int)((span0 / stepB) % sizeB);
            } else {
                c = ok ? (int)(((uint32_t)(i4 << 2) / stepB) % sizeB) : 0;
            }
            if (p.b) b = __ldg(p.b + c);
        }
        float4 r = (G > 0 && A == 9) ? vr[it] : make_float4(0, 0, 0, 0);
        float4 yr = (G > 0) ? vy[it] : make_float4(0, 0, 0, 0);
        float4 d = (G == 2) ? vd[it] : make_float4(1, 1, 1, 1);
        const float4 nz = (G == 0) ? vn[it] : make_float4(0, 0, 0, 0);
        float4 o;
        o.x = eval<A, G>(vx[it].x, b + nz.x, r.x, yr.x, d.x, p);
        o.y = eval<A, G>(vx[it].y, b + nz.y, r.y, yr.y, d.y, p);
        o.z = eval<A, G>(vx[it].z, b + nz.z, r.z, yr.z, d.z, p);
        o.w = eval<A, G>(vx[it].w, b + nz.w, r.w, yr.w, d.w, p);
        if (ok) reinterpret_cast<float4*>(p.y)[i4] = o;
        if (p.dbias) {
            float s = ok ? (o.x + o.y) + (o.z + o.w) : 0.f;
            if (UNIFORM) {
                if (c != acc_c) {  // warp-uniform branch
                    if (acc_c >= 0) { float t = gg::warp_sum(acc); if (lane == 0) atomicAdd(p.dbias + acc_c, t); }
                    acc_c = c; acc = 0.f;
                }
                acc += s;
            } else if (ok) {
                atomicAdd(p.dbias + c, s);
            }
        }
    }
    if (UNIFORM && p.dbias && acc_c >= 0) {
        float t = gg::warp_sum(acc);
        if (lane == 0) atomicAdd(p.dbias + acc_c, t);
    }
}

// Generic scalar kernel: any stepB (incl. 1 = bias along the innermost dim), any alignment.
template <int A, int G>
__global__ void __launch_bounds__(256) bias_act_scalar(Params p) {
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (; i < p.sizeX; i += stride) {
        int c = (p.b || p.dbias || p.noise) ? (int)((i / p.stepB) % p.sizeB) : 0;
        float b = p.b ? __ldg(p.b + c) : 0.f;
        if (G == 0 && p.noise) { const int64_t plane = i / p.stepB; b += __ldg(p.noise + (plane / p.sizeB) * p.noise_bs + (i - plane * p.stepB)); }
        float xr = (G > 0 && p.xref) ? __ldg(p.xref + i) : 0.f;
        float yr = (G > 0 && p.yref) ? __ldg(p.yref + i) : 0.f;
        float d = (G == 2 && p.dy) ? __ldg(p.dy + i) : 1.f;
        float o = eval<A, G>(__ldg(p.x + i), b, xr, yr, d, p);
        p.y[i] = o;
        if (p.dbias) atomicAdd(p.dbias + c, o);
    }
}

template <int A, int G>
int launch(const Params& p, cudaStream_t st) {
    auto al16 = [](const void* q) { return q == nullptr || (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    bool vec = (p.sizeX % 4 == 0) && (p.stepB % 4 == 0) && al16(p.noise) && (p.noise_bs % 4 == 0) && al16(p.x) && al16(p.xref) && al16(p.yref) && al16(p.dy) && al16(p.y);
    if (vec) {
        const int64_t per_block = 8 /*warps*/ * 32 * 8 /*ITER*/ * 4;
        int64_t grid = (p.sizeX + per_block - 1) / per_block;
        if (p.stepB % 128 == 0) bias_act_vec4<A, G, true><<<(unsigned)grid, 256, 0, st>>>(p);
        else                    bias_act_vec4<A, G, false><<<(unsigned)grid, 256, 0, st>>>(p);
    } else {
        int64_t grid = (p.sizeX + 255) / 256;
        if (grid > GG_NUM_SMS * 16) grid = GG_NUM_SMS * 16;
        bias_act_scalar<A, G><<<(unsigned)grid, 256, 0, st>>>(p);
    }
    return gg::check_launch("bias_act");
}

template <int A>
int launch_g(const Params& p, int grad, cudaStream_t st) {
    if (grad == 0) return launch<A, 0>(p, st);
    if (grad == 1) return launch<A, 1>(p, st);
    return launch<A, 2>(p, st);
}

}  // namespace

static int bias_act_impl(const float* x, const float* b, const float* xref, const float* yref, const float* dy,
                         float* y, float* dbias, int grad, int act, float alpha, float gain, float clamp,
                         int64_t sizeX, int sizeB, int64_t stepB, const float* noise, int64_t noise_bs, gg_stream_t stream) {
    GG_REQUIRE(x != nullptr && y != nullptr, "bias_act: x and y must be non-null");
    GG_REQUIRE(sizeX >= 0 && sizeX <= 0x7fffffffLL, "bias_act: x is too large");  // bias_act.cpp:40
    GG_REQUIRE(grad >= 0 && grad <= 2, "bias_act: grad must be 0, 1 or 2");
    GG_REQUIRE(act >= 1 && act <= 9, "bias_act: no CUDA kernel found for the specified activation func");
    GG_REQUIRE(b == nullptr || (sizeB >= 1 && stepB >= 1), "bias_act: b has wrong number of elements");
    GG_REQUIRE(dbias == nullptr || (sizeB >= 1 && stepB >= 1), "bias_act: dbias needs sizeB/stepB");
    if (sizeX == 0) return GG_OK;
    GG_REQUIRE(noise == nullptr || (grad == 0 && sizeB >= 1 && stepB >= 1 && noise_bs >= 0), "bias_act: noise needs grad == 0 and sizeB/stepB");
    if (b == nullptr && dbias == nullptr && noise == nullptr) { sizeB = 1; stepB = (int64_t)1 << 30; }   // no per-channel work: any layout vectorises
    Params p{x, b, xref, yref, dy, y, dbias, alpha, gain, clamp, sizeX, sizeB < 1 ? 1 : sizeB, stepB < 1 ? 1 : stepB, noise, noise_bs};
    cudaStream_t st = (cudaStream_t)stream;
    switch (act) {
        case 1: return launch_g<1>(p, grad, st);
        case 2: return launch_g<2>(p, grad, st);
        case 3: return launch_g<3>(p, grad, st);
        case 4: return launch_g<4>(p, grad, st);
        case 5: return launch_g<5>(p, grad, st);
        case 6: return launch_g<6>(p, grad, st);
        case 7: return launch_g<7>(p, grad, st);
        case 8: return launch_g<8>(p, grad, st);
        default: return launch_g<9>(p, grad, st);
    }
}

extern "C" GG_API int gg_bias_act_f32(const float* x, const float* b, const float* xref, const float* yref, const float* dy,
                               float* y, float* dbias, int grad, int act, float alpha, float gain, float clamp,
                               int64_t sizeX, int sizeB, int64_t stepB, gg_stream_t stream) {
    return bias_act_impl(x, b, xref, yref, dy, y, dbias, grad, act, alpha, gain, clamp, sizeX, sizeB, stepB, nullptr, 0, stream);
}

extern "C" GG_API int gg_bias_act_noise_f32(const float* x, const float* b, const float* noise, int64_t noise_batch_stride, float* y, int act,
                                     float alpha, float gain, float clamp, int64_t sizeX, int sizeB, int64_t stepB, gg_stream_t stream) {
    return bias_act_impl(x, b, nullptr, nullptr, nullptr, y, nullptr, 0, act, alpha, gain, clamp, sizeX, sizeB, stepB, noise, noise_batch_stride, stream);
}
