// Per-(sample, channel) dot products over the pixel axis:  out[n,c] = sum_p a[n,c,p] * b[n,c,p].
//
// These are the style / demodulation-coefficient gradients of modulated_conv2d when the per-sample scales are folded into
// the convolution kernels (training/networks.py:642 `x * styles`, :648-651 `fma(x, dcoefs, noise)` and their autograd):
//   d styles[n,i] = sum_p x[n,i,p] * dx[n,i,p] / styles[n,i]        d dcoefs[n,o] = sum_p dy[n,o,p] * y[n,o,p] / dcoefs[n,o]
// HBM-streaming: two 128-bit loads per 8 FLOPs, one read of each tensor, fp32 atomics for the per-row partial sums.
#include "common.cuh"

namespace {

struct DotP {
    const float* a; const float* b; float* out;
    int64_t rows, P;
    int chunks, vec;
};

__global__ void __launch_bounds__(256) chan_dot_kernel(DotP p) {
    __shared__ float red[8];
    const int64_t row = blockIdx.y;
    const float* ar = p.a + row * p.P;
    const float* br = p.b + row * p.P;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, acc3 = 0.f;
    if (p.vec) {
        const int64_t P4 = p.P >> 2;
        const int64_t per = (P4 + p.chunks - 1) / p.chunks;
        const int64_t q0 = (int64_t)blockIdx.x * per, q1 = min(P4, q0 + per);
        const float4* a4 = reinterpret_cast<const float4*>(ar);
        const float4* b4 = reinterpret_cast<const float4*>(br);
        int64_t q = q0 + threadIdx.x;
        for (; q + 3 * 256 < q1; q += 4 * 256) {               // 8 loads in flight per thread
            const float4 x0 = __ldg(a4 + q), x1 = __ldg(a4 + q + 256), x2 = __ldg(a4 + q + 512), x3 = __ldg(a4 + q + 768);
            const float4 y0 = __ldg(b4 + q), y1 = __ldg(b4 + q + 256), y2 = __ldg(b4 + q + 512), y3 = __ldg(b4 + q + 768);
            acc0 = fmaf(x0.x, y0.x, fmaf(x0.y, y0.y, fmaf(x0.z, y0.z, fmaf(x0.w, y0.w, acc0))));
            acc1 = fmaf(x1.x, y1.x, fmaf(x1.y, y1.y, fmaf(x1.z, y1.z, fmaf(x1.w, y1.w, acc1))));
            acc2 = fmaf(x2.x, y2.x, fmaf(x2.y, y2.y, fmaf(x2.z, y2.z, fmaf(x2.w, y2.w, acc2))));
            acc3 = fmaf(x3.x, y3.x, fmaf(x3.y, y3.y, fmaf(x3.z, y3.z, fmaf(x3.w, y3.w, acc3))));
        }
        for (; q < q1; q += 256) {
            const float4 x0 = __ldg(a4 + q), y0 = __ldg(b4 + q);
            acc0 = fmaf(x0.x, y0.x, fmaf(x0.y, y0.y, fmaf(x0.z, y0.z, fmaf(x0.w, y0.w, acc0))));
        }
    } else {
        const int64_t per = (p.P + p.chunks - 1) / p.chunks;
        const int64_t q0 = (int64_t)blockIdx.x * per, q1 = min(p.P, q0 + per);
        for (int64_t q = q0 + threadIdx.x; q < q1; q += 256) acc0 = fmaf(__ldg(ar + q), __ldg(br + q), acc0);
    }
    float s = gg::warp_sum((acc0 + acc1) + (acc2 + acc3));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        if (p.chunks > 1) atomicAdd(p.out + row, t); else p.out[row] = t;
    }
}


// out[n,c] = sum_p ds[n,c,p] * (pre(y[n,c,p]) - bias[c] - noise[n,p]),  pre = the inverse of y = act(.) * gain for the invertible
// activations (linear, lrelu with alpha != 0).  This is the demodulation-coefficient gradient of a convolution whose bias / noise /
// activation epilogue was fused into the kernel (gg_conv2d_act_f32): the pre-activation tensor is never materialised, it is
// reconstructed from the saved output.
struct PreP {
    const float* ds; const float* y; const float* bias; const float* noise; float* out;
    int64_t rows, P, noise_bs;
    int C, chunks, act;
    float inv_gain, inv_gain_alpha;
};

__global__ void __launch_bounds__(256) chan_dot_preact_kernel(PreP p) {
    __shared__ float red[8];
    const int64_t row = blockIdx.y;
    const int n = (int)(row / p.C), c = (int)(row - (int64_t)n * p.C);
    const float4* d4 = reinterpret_cast<const float4*>(p.ds + row * p.P);
    const float4* y4 = reinterpret_cast<const float4*>(p.y + row * p.P);
    const float4* n4 = p.noise ? reinterpret_cast<const float4*>(p.noise + (int64_t)n * p.noise_bs) : nullptr;
    const float b = p.bias ? __ldg(p.bias + c) : 0.f;
    const int64_t P4 = p.P >> 2;
    const int64_t per = (P4 + p.chunks - 1) / p.chunks;
    const int64_t q0 = (int64_t)blockIdx.x * per, q1 = min(P4, q0 + per);
    float acc = 0.f;
    auto pre = [&](float v) { return (p.act == 3 && !(v > 0.f)) ? v * p.inv_gain_alpha : v * p.inv_gain; };
    for (int64_t q = q0 + threadIdx.x; q < q1; q += 256) {
        const float4 d = __ldg(d4 + q), y = __ldg(y4 + q);
        const float4 z = n4 ? __ldg(n4 + q) : make_float4(0.f, 0.f, 0.f, 0.f);
        acc = fmaf(d.x, pre(y.x) - (b + z.x), fmaf(d.y, pre(y.y) - (b + z.y), fmaf(d.z, pre(y.z) - (b + z.z), fmaf(d.w, pre(y.w) - (b + z.w), acc))));
    }
    float s = gg::warp_sum(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) t += red[w];
        if (p.chunks > 1) atomicAdd(p.out + row, t); else p.out[row] = t;
    }
}

}  // namespace

extern "C" GG_API int gg_chan_dot_preact_f32(const float* ds, const float* y, const float* bias, const float* noise, int64_t noise_batch_stride,
                                      float* out, int N, int C, int64_t P, int act, float alpha, float gain, gg_stream_t stream) {
    GG_REQUIRE(ds && y && out, "chan_dot_preact: null pointer");
    GG_REQUIRE((act == 1 || (act == 3 && alpha != 0.f)) && gain != 0.f, "chan_dot_preact: the activation must be invertible (linear, or lrelu with alpha != 0)");
    GG_REQUIRE(N >= 0 && C >= 1 && P >= 4 && P % 4 == 0 && (int64_t)N * C * P <= 0x7fffffffLL, "chan_dot_preact: planes must be non-empty multiples of 4 elements");
    GG_REQUIRE(((reinterpret_cast<uintptr_t>(ds) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(noise)) & 15) == 0 &&
               (noise == nullptr || noise_batch_stride == 0 || noise_batch_stride == P), "chan_dot_preact: operands must be 16-byte aligned; noise is [P] or [N,P]");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t rows = (int64_t)N * C;
    if (rows == 0) return GG_OK;
    PreP p{ds, y, bias, noise, out, rows, P, noise_batch_stride, C, 1, act, 1.f / gain, 1.f / (gain * (act == 3 ? alpha : 1.f))};
    int64_t chunks = (6LL * GG_NUM_SMS + rows - 1) / rows;
    const int64_t maxc = (P + 4095) / 4096;
    if (chunks > maxc) chunks = maxc;
    if (chunks < 1) chunks = 1;
    p.chunks = (int)chunks;
    GG_REQUIRE(rows <= 65535, "chan_dot_preact: too many (sample, channel) rows");
    if (p.chunks > 1) GG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows, st));
    dim3 grid((unsigned)p.chunks, (unsigned)rows);
    chan_dot_preact_kernel<<<grid, 256, 0, st>>>(p);
    return gg::check_launch("chan_dot_preact");
}

extern "C" GG_API int gg_chan_dot_f32(const float* a, const float* b, float* out, int64_t rows, int64_t P, gg_stream_t stream) {
    GG_REQUIRE(a && b && out, "chan_dot: null pointer");
    GG_REQUIRE(rows >= 0 && P >= 0 && rows <= 0x7fffffffLL && rows * P <= 0x7fffffffLL * 4, "chan_dot: tensor is too large");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows == 0) return GG_OK;
    if (P == 0) { GG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows, st)); return GG_OK; }
    DotP p{a, b, out, rows, P, 1, 0};
    p.vec = (P % 4 == 0 && ((reinterpret_cast<uintptr_t>(a) | reinterpret_cast<uintptr_t>(b)) & 15) == 0) ? 1 : 0;
    // enough CTAs to fill the machine a few times over, but at least 4 x 256 x 4 elements per CTA
    int64_t chunks = (6LL * GG_NUM_SMS + rows - 1) / rows;
    const int64_t maxc = (P + 4095) / 4096;
    if (chunks > maxc) chunks = maxc;
    if (chunks < 1) chunks = 1;
    if (chunks > 65535) chunks = 65535;
    p.chunks = (int)chunks;
    GG_REQUIRE(rows <= 65535LL * 1024, "chan_dot: too many rows");
    if (p.chunks > 1) GG_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)rows, st));
    // grid.y is limited to 65535 rows per launch
    for (int64_t r0 = 0; r0 < rows; r0 += 65535) {
        DotP q = p;
        q.a = a + r0 * P; q.b = b + r0 * P; q.out = out + r0;
        q.rows = rows - r0 < 65535 ? rows - r0 : 65535;
        dim3 grid((unsigned)p.chunks, (unsigned)q.rows);
        chan_dot_kernel<<<grid, 256, 0, st>>>(q);
        int rc = gg::check_launch("chan_dot");
        if (rc != GG_OK) return rc;
    }
    return GG_OK;
}
