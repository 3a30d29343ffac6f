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

// ------------------------------------------------------------------------------------------------
// y[r, p] = s[r] * x[r, p]: a per-(sample, channel) scale broadcast over the pixels of its plane.  The second-order gradients of
// modulated_conv2d (path-length regularisation) need g[n,c] * t[n,c,:,:] as a tensor in two places per layer; torch's broadcasting
// multiply takes its non-vectorised kernel for that stride pattern.  Flat grid-stride over 128-bit groups; the row of a group is one
// integer division (P % 4 == 0, so a group never straddles two rows).
namespace {
__global__ void __launch_bounds__(256) scale_rows_vec4(const float4* __restrict__ x, const float* __restrict__ s, float4* __restrict__ y,
                                                       int64_t total4, int64_t P4) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += stride) {
        const float f = __ldg(s + i / P4);
        const float4 v = __ldg(x + i);
        y[i] = make_float4(f * v.x, f * v.y, f * v.z, f * v.w);
    }
}
__global__ void __launch_bounds__(256) scale_rows_scalar(const float* __restrict__ x, const float* __restrict__ s, float* __restrict__ y,
                                                         int64_t total, int64_t P) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) y[i] = __ldg(s + i / P) * __ldg(x + i);
}
}  // namespace

extern "C" GG_API int gg_scale_rows_f32(const float* x, const float* s, float* y, int64_t rows, int64_t P, gg_stream_t stream) {
    GG_REQUIRE(x && s && y, "scale_rows: null pointer");
    GG_REQUIRE(rows >= 0 && P >= 0 && (rows == 0 || P <= 0x7fffffffLL * 4 / rows), "scale_rows: tensor is too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = rows * P;
    if (total == 0) return GG_OK;
    const bool vec = P % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const int64_t work = vec ? total / 4 : total;
    int64_t blocks = (work + 255) / 256;
    if (blocks > (int64_t)GG_NUM_SMS * 16) blocks = (int64_t)GG_NUM_SMS * 16;
    if (vec) scale_rows_vec4<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(x), s, reinterpret_cast<float4*>(y), total / 4, P / 4);
    else scale_rows_scalar<<<(unsigned)blocks, 256, 0, st>>>(x, s, y, total, P);
    return gg::check_launch("scale_rows");
}

// y[r, p] = s[r] * x[r, p] + z[(r / C) * zbs + p]: fma(x, dcoefs[N,C,1,1], noise) of the reference's non-fused modulated_conv2d
// (training/networks.py:648, torch_utils/ops/fma.py:15-16) for callers that keep that formulation (this build's own modulated_conv2d
// carries both terms inside the convolution kernel).  z is one [P] plane shared by the batch (zbs == 0) or one per sample (zbs == P).
namespace {
// one row (= one (sample, channel) plane) per blockIdx.x, its columns spread over blockIdx.y: no per-element index division
template <typename V>
__global__ void __launch_bounds__(256) fma_rows_kernel(const V* __restrict__ x, const float* __restrict__ s, const V* __restrict__ z,
                                                       V* __restrict__ y, int PV, int C, int zbsV) {
    const int r = blockIdx.x;
    const float f = __ldg(s + r);
    const V* xr = x + (int64_t)r * PV;
    const V* zr = z + (int64_t)(r / C) * zbsV;
    V* yr = y + (int64_t)r * PV;
    for (int64_t p = (int64_t)blockIdx.y * blockDim.x + threadIdx.x; p < PV; p += (int64_t)gridDim.y * blockDim.x) {
        if constexpr (sizeof(V) == 16) {
            const float4 v = __ldg(xr + p), w = __ldg(zr + p);
            yr[p] = make_float4(fmaf(f, v.x, w.x), fmaf(f, v.y, w.y), fmaf(f, v.z, w.z), fmaf(f, v.w, w.w));
        } else {
            yr[p] = fmaf(f, __ldg(xr + p), __ldg(zr + p));
        }
    }
}
}  // namespace

extern "C" GG_API int gg_fma_rows_f32(const float* x, const float* s, const float* z, int64_t z_batch_stride, float* y, int64_t rows, int64_t C,
                                      int64_t P, gg_stream_t stream) {
    GG_REQUIRE(x && s && z && y, "fma_rows: null pointer");
    GG_REQUIRE(rows >= 0 && P >= 0 && C >= 1 && rows % C == 0 && rows <= 0x7fffffffLL && P <= 0x7fffffffLL &&
               (rows == 0 || P <= 0x7fffffffLL * 4 / rows), "fma_rows: bad extents or tensor too large");
    GG_REQUIRE(z_batch_stride == 0 || z_batch_stride == P, "fma_rows: z is one [P] plane (stride 0) or one per sample (stride P)");
    cudaStream_t st = (cudaStream_t)stream;
    if (rows * P == 0) return GG_OK;
    const bool vec = P % 4 == 0 && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(z)) & 15) == 0;
    const int64_t PV = vec ? P / 4 : P;
    const int threads = (int)(PV >= 256 ? 256 : (PV + 31) / 32 * 32);
    int64_t chunks = (PV + threads - 1) / threads;
    if (chunks > 4096) chunks = 4096;
    const dim3 grid((unsigned)rows, (unsigned)chunks);
    if (vec) fma_rows_kernel<float4><<<grid, threads, 0, st>>>(reinterpret_cast<const float4*>(x), s, reinterpret_cast<const float4*>(z),
                                                               reinterpret_cast<float4*>(y), (int)PV, (int)C, (int)(z_batch_stride / 4));
    else fma_rows_kernel<float><<<grid, threads, 0, st>>>(x, s, z, y, (int)PV, (int)C, (int)z_batch_stride);
    return gg::check_launch("fma_rows");
}

// y[r, p] = s1[r] * x1[r, p] + s2[r] * x2[r, p]: the two gradient contributions that meet in front of a convolution in the second-order
// pass (a * gg_dx + g_da * x), combined in one pass so that ONE convolution / weight-gradient launch serves both.
namespace {
__global__ void __launch_bounds__(256) axpby_rows_vec4(const float4* __restrict__ x1, const float* __restrict__ s1, const float4* __restrict__ x2,
                                                       const float* __restrict__ s2, float4* __restrict__ y, int64_t total4, int64_t P4) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total4; i += stride) {
        const int64_t r = i / P4;
        const float f1 = __ldg(s1 + r), f2 = __ldg(s2 + r);
        const float4 u = __ldg(x1 + i), v = __ldg(x2 + i);
        y[i] = make_float4(fmaf(f1, u.x, f2 * v.x), fmaf(f1, u.y, f2 * v.y), fmaf(f1, u.z, f2 * v.z), fmaf(f1, u.w, f2 * v.w));
    }
}
__global__ void __launch_bounds__(256) axpby_rows_scalar(const float* __restrict__ x1, const float* __restrict__ s1, const float* __restrict__ x2,
                                                         const float* __restrict__ s2, float* __restrict__ y, int64_t total, int64_t P) {
    const int64_t stride = (int64_t)gridDim.x * 256;
    for (int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x; i < total; i += stride) {
        const int64_t r = i / P;
        y[i] = fmaf(__ldg(s1 + r), __ldg(x1 + i), __ldg(s2 + r) * __ldg(x2 + i));
    }
}
}  // namespace

extern "C" GG_API int gg_axpby_rows_f32(const float* x1, const float* s1, const float* x2, const float* s2, float* y, int64_t rows, int64_t P,
                                        gg_stream_t stream) {
    GG_REQUIRE(x1 && s1 && x2 && s2 && y, "axpby_rows: null pointer");
    GG_REQUIRE(rows >= 0 && P >= 0 && (rows == 0 || P <= 0x7fffffffLL * 4 / rows), "axpby_rows: tensor is too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t total = rows * P;
    if (total == 0) return GG_OK;
    const bool vec = P % 4 == 0 && ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(y)) & 15) == 0;
    const int64_t work = vec ? total / 4 : total;
    int64_t blocks = (work + 255) / 256;
    if (blocks > (int64_t)GG_NUM_SMS * 16) blocks = (int64_t)GG_NUM_SMS * 16;
    if (vec) axpby_rows_vec4<<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(x1), s1, reinterpret_cast<const float4*>(x2), s2,
                                                              reinterpret_cast<float4*>(y), total / 4, P / 4);
    else axpby_rows_scalar<<<(unsigned)blocks, 256, 0, st>>>(x1, s1, x2, s2, y, total, P);
    return gg::check_launch("axpby_rows");
}
