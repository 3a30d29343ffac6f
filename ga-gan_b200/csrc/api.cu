// C-ABI glue of libgagan_b200.so: error state, argument validation and kernel-family dispatch.
#include "tc_common.cuh"
#include <string.h>

namespace gg {

static thread_local char t_err[512] = "";
std::atomic<int64_t> g_launches{0};
int g_march_enabled = 1;

// 64 bytes of host-mapped, portable pinned memory: the wait-loop watchdog of the tcgen05 kernels writes its record here before it
// traps, so the cause of a faulted context can still be read (tc_common.cuh)
unsigned long long* watchdog_host_record() {
    static std::atomic<unsigned long long*> rec{nullptr};
    static std::atomic<int> tried{0};
    if (tried.exchange(1) == 0) {
        void* h = nullptr;
        if (cudaHostAlloc(&h, 64, cudaHostAllocMapped | cudaHostAllocPortable) == cudaSuccess && h != nullptr) {
            memset(h, 0, 64);
            rec.store(static_cast<unsigned long long*>(h));
        } else {
            (void)cudaGetLastError();
        }
    }
    return rec.load();
}

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(t_err, sizeof(t_err), fmt, ap);
    va_end(ap);
}

// conv_simt.cu
int conv2d_simt(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, int, int, int,
                const float*, const float*, cudaStream_t);
int conv2d_wgrad_simt(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, int, int, int,
                      const float*, const float*, cudaStream_t);
// conv_thin.cu (HBM-streaming 1x1 kernels with <= 4 channels on one side: ToRGB / fromRGB)
bool conv1x1_thin_eligible(const float* x, const float* y, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW,
                           int stride, int pad_y, int pad_x);
int conv1x1_thin(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int w_io, const float* in_scale,
                 const float* out_scale, cudaStream_t st);
bool wgrad1x1_thin_eligible(const float* a, const float* b, int N, int A, int HA, int WA, int B, int HB, int WB, int KH, int KW,
                            int stride, int pad_y, int pad_x);
int wgrad1x1_thin(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int out_layout,
                  const float* a_scale, const float* b_scale, cudaStream_t st);
// conv_tc.cu (tcgen05 / TMEM / TMA path)
bool conv2d_tc_eligible(int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int stride, int pad_y, int pad_x,
                        int transposed);
int conv2d_tc(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int pad_y,
              int pad_x, int flip_w, int w_is_IO, const float* in_scale, const float* out_scale, int nprod, const ggtc::ConvEpilogue* epi,
              cudaStream_t st);
// conv_march.cu (row-marching tcgen05 path for <= 64 output channels)
bool conv2d_march_eligible(const float* x, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int stride, int pad_y, int pad_x);
int conv2d_march(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int K, int OH, int OW, int pad_y, int pad_x,
                 int flip_w, int w_is_IO, const float* in_scale, const float* out_scale, int nprod, const ggtc::ConvEpilogue* epi, cudaStream_t st);
bool wgrad_tc_eligible(int N, int A, int HA, int WA, int B, int HB, int WB, int KH, int KW, int stride, int pad_y, int pad_x);
int wgrad_tc(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int KH, int KW,
             int pad_y, int pad_x, int flip_w, int out_layout, const float* a_scale, const float* b_scale, int nprod, int pm_dim,
             unsigned pm_dead, cudaStream_t st);

}  // namespace gg

extern "C" GG_API const char* gg_last_error(void) { return gg::t_err; }
extern "C" GG_API int gg_version(void) { return 100; }
extern "C" GG_API int64_t gg_launch_count(void) { return gg::g_launches.load(); }
extern "C" GG_API int gg_watchdog_report(char* buf, int len) {
    unsigned long long* r = gg::watchdog_host_record();
    if (buf == nullptr || len <= 0) return 0;
    buf[0] = 0;
    if (r == nullptr) return 0;
    // layout written by mbar_watchdog (tc_common.cuh), 32-bit words: [0] = 0x5744 << 16 | kernel family, [2] = CTA, [3] = thread, [4] = barrier
    const volatile unsigned int* w = reinterpret_cast<const volatile unsigned int*>(r);
    if ((w[0] >> 16) != 0x5744u) return 0;
    static const char* const names[] = {"?", "conv_tc_kernel", "conv_march_kernel", "wgrad_tc_kernel", "wgrad_tma_kernel"};
    const unsigned tag = (w[0] & 0xffffu) < 5 ? (w[0] & 0xffffu) : 0;
    snprintf(buf, (size_t)len, "watchdog: %s CTA %u thread %u waited more than 8 G cycles on the mbarrier at shared address 0x%x", names[tag],
             w[2], w[3], w[4]);
    return 1;
}
extern "C" GG_API int gg_set_conv_kernel_family(int family) { const int old = gg::g_march_enabled; gg::g_march_enabled = (family != 0); return old; }

extern "C" GG_API int gg_device_ok(void) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    int major = 0, minor = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return (major == 10 && minor == 0) ? 1 : 0;
}

// `epi` (may be null): bias / noise / activation to apply to the result.  *epi_fused tells the caller whether the convolution kernel
// did it in its epilogue; if not, the caller runs the bias_act kernel over y.
static int conv2d_impl(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW,
                       int OH, int OW, int stride, int pad_y, int pad_x, int transposed, int flip_w,
                       const float* in_scale, const float* out_scale, int prec, int* used_prec, const ggtc::ConvEpilogue* epi,
                       bool* epi_fused, gg_stream_t stream) {
    if (epi_fused) *epi_fused = false;
    GG_REQUIRE(x && w && y, "conv2d: null pointer");
    GG_REQUIRE(N >= 0 && I >= 1 && H >= 1 && W >= 1 && O >= 1 && KH >= 1 && KW >= 1, "conv2d: bad shape");
    GG_REQUIRE(stride >= 1 && pad_y >= 0 && pad_x >= 0, "conv2d: bad stride/padding");
    GG_REQUIRE(OH >= 1 && OW >= 1, "conv2d: output must be at least 1x1");
    if (stride != 1) {   // stride 1: the output extent is free (positions that read outside x see zeros)
        if (!transposed) {
            GG_REQUIRE(OH == (H + 2 * pad_y - KH) / stride + 1 && OW == (W + 2 * pad_x - KW) / stride + 1, "conv2d: output size mismatch");
        } else {
            int bh = (H - 1) * stride - 2 * pad_y + KH, bw = (W - 1) * stride - 2 * pad_x + KW;
            GG_REQUIRE(OH >= bh && OH < bh + stride && OW >= bw && OW < bw + stride, "conv_transpose2d: output size mismatch");
        }
    }
    GG_REQUIRE((int64_t)N * I * H * W <= 0x7fffffffLL && (int64_t)N * O * OH * OW <= 0x7fffffffLL, "conv2d: tensor is too large");
    GG_REQUIRE(prec == GG_PREC_AUTO || prec == GG_PREC_AUTO_FAST || prec == GG_PREC_FP32_SIMT || prec == GG_PREC_TF32X1 || prec == GG_PREC_TF32X3,
               "conv2d: unknown precision mode %d", prec);
    cudaStream_t st = (cudaStream_t)stream;
    bool tc_ok = gg::conv2d_tc_eligible(N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x, transposed);
    // the TMA descriptor of x needs a 16-byte aligned base: an unaligned view falls back to the exact FFMA kernel under AUTO
    const bool is_auto = (prec == GG_PREC_AUTO || prec == GG_PREC_AUTO_FAST);
    const int auto_tc = (prec == GG_PREC_AUTO_FAST) ? GG_PREC_TF32X1 : GG_PREC_TF32X3;
    if (is_auto && (reinterpret_cast<uintptr_t>(x) & 15) != 0) tc_ok = false;
    if ((prec == GG_PREC_TF32X1 || prec == GG_PREC_TF32X3) && !tc_ok) {
        gg::set_error("conv2d: shape N=%d I=%d H=%d W=%d O=%d k=%dx%d stride=%d transposed=%d is not served by the tcgen05 path",
                      N, I, H, W, O, KH, KW, stride, transposed);
        return GG_EUNSUPPORTED;
    }
    const bool thin_ok = N > 0 && gg::conv1x1_thin_eligible(x, y, N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x);
    int use = is_auto ? ((tc_ok && !thin_ok) ? auto_tc : GG_PREC_FP32_SIMT) : prec;
    if (used_prec) *used_prec = use;
    if (use == GG_PREC_FP32_SIMT && thin_ok)
        return gg::conv1x1_thin(x, w, y, N, I, H, W, O, transposed, in_scale, out_scale, st);
    if (use == GG_PREC_FP32_SIMT)
        return gg::conv2d_simt(x, w, y, N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x, transposed, flip_w, in_scale,
                               out_scale, st);
    // stride-1 conv_transpose2d == correlation with the flipped kernel and padding K-1-p; both are
    // handled inside the tensor-core path through its weight-packing step.
    const int tpy = transposed ? KH - 1 - pad_y : pad_y, tpx = transposed ? KW - 1 - pad_x : pad_x;
    const bool fuse = epi && epi->act >= 1 && epi->act <= 3;          // linear / relu / lrelu ride in the store loop of the tcgen05 kernels
    if (epi_fused) *epi_fused = fuse;
    if (gg::g_march_enabled && gg::conv2d_march_eligible(x, N, I, H, W, O, KH, KW, OH, OW, stride, tpy, tpx))
        return gg::conv2d_march(x, w, y, N, I, H, W, O, KH, OH, OW, tpy, tpx, transposed ? !flip_w : flip_w, transposed, in_scale, out_scale, use,
                                fuse ? epi : nullptr, st);
    return gg::conv2d_tc(x, w, y, N, I, H, W, O, KH, KW, OH, OW, transposed ? KH - 1 - pad_y : pad_y, transposed ? KW - 1 - pad_x : pad_x,
                         transposed ? !flip_w : flip_w, transposed, in_scale, out_scale, use, fuse ? epi : nullptr, st);
}

extern "C" GG_API int gg_conv2d_f32(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW,
                             int OH, int OW, int stride, int pad_y, int pad_x, int transposed, int flip_w,
                             const float* in_scale, const float* out_scale, int prec, int* used_prec, gg_stream_t stream) {
    return conv2d_impl(x, w, y, N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x, transposed, flip_w, in_scale, out_scale, prec, used_prec,
                       nullptr, nullptr, stream);
}

extern "C" GG_API int gg_conv2d_act_f32(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW,
                                 int OH, int OW, int stride, int pad_y, int pad_x, int transposed, int flip_w,
                                 const float* in_scale, const float* out_scale, const float* bias, const float* noise,
                                 int64_t noise_batch_stride, int act, float alpha, float gain, float clamp, int prec, int* used_prec,
                                 int* fused, gg_stream_t stream) {
    GG_REQUIRE(act >= 1 && act <= 9, "conv2d_act: act must be the reference's cuda_idx 1..9");
    GG_REQUIRE(noise == nullptr || noise_batch_stride == 0 || noise_batch_stride == (int64_t)OH * OW, "conv2d_act: noise must be [OH*OW] or [N, OH*OW]");
    ggtc::ConvEpilogue epi{bias, noise, (long long)noise_batch_stride, act, alpha, gain, clamp};
    bool did = false;
    int rc = conv2d_impl(x, w, y, N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x, transposed, flip_w, in_scale, out_scale, prec, used_prec,
                         &epi, &did, stream);
    if (fused) *fused = did ? 1 : 0;
    if (rc != GG_OK || did || N == 0) return rc;
    // the kernel family that served this shape (FFMA / thin 1x1) has no epilogue: the same arithmetic as a second launch, in place
    if (noise == nullptr)
        return gg_bias_act_f32(y, bias, nullptr, nullptr, nullptr, y, nullptr, 0, act, alpha, gain, clamp, (int64_t)N * O * OH * OW, O, (int64_t)OH * OW, stream);
    return gg_bias_act_noise_f32(y, bias, noise, noise_batch_stride, y, act, alpha, gain, clamp, (int64_t)N * O * OH * OW, O, (int64_t)OH * OW, stream);
}

extern "C" GG_API int gg_conv2d_wgrad_f32(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB,
                                   int KH, int KW, int stride, int pad_y, int pad_x, int flip_w, int out_layout,
                                   const float* a_scale, const float* b_scale, int prec, int* used_prec, gg_stream_t stream) {
    return gg_conv2d_wgrad_pm_f32(a, b, dw, N, A, HA, WA, B, HB, WB, KH, KW, stride, pad_y, pad_x, flip_w, out_layout, a_scale, b_scale,
                                  prec, used_prec, 0, 0u, stream);
}

extern "C" GG_API int gg_conv2d_wgrad_pm_f32(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB,
                                      int KH, int KW, int stride, int pad_y, int pad_x, int flip_w, int out_layout,
                                      const float* a_scale, const float* b_scale, int prec, int* used_prec, int pm_dim,
                                      unsigned pm_dead, gg_stream_t stream) {
    GG_REQUIRE(pm_dim >= 0 && pm_dim <= 2 && pm_dead <= 0xffffu, "conv2d_wgrad: bad phase-major hint");
    GG_REQUIRE(a && b && dw, "conv2d_wgrad: null pointer");
    GG_REQUIRE(N >= 0 && A >= 1 && B >= 1 && HA >= 1 && WA >= 1 && HB >= 1 && WB >= 1 && KH >= 1 && KW >= 1, "conv2d_wgrad: bad shape");
    GG_REQUIRE(stride >= 1 && pad_y >= 0 && pad_x >= 0, "conv2d_wgrad: bad stride/padding");
    GG_REQUIRE((int64_t)N * A * HA * WA <= 0x7fffffffLL && (int64_t)N * B * HB * WB <= 0x7fffffffLL, "conv2d_wgrad: tensor is too large");
    GG_REQUIRE(prec == GG_PREC_AUTO || prec == GG_PREC_AUTO_FAST || prec == GG_PREC_FP32_SIMT || prec == GG_PREC_TF32X1 || prec == GG_PREC_TF32X3,
               "conv2d_wgrad: unknown precision mode %d", prec);
    cudaStream_t st = (cudaStream_t)stream;
    bool tc_ok = gg::wgrad_tc_eligible(N, A, HA, WA, B, HB, WB, KH, KW, stride, pad_y, pad_x);
    if ((prec == GG_PREC_TF32X1 || prec == GG_PREC_TF32X3) && !tc_ok) {
        gg::set_error("conv2d_wgrad: shape is not served by the tcgen05 path");
        return GG_EUNSUPPORTED;
    }
    const bool thin_ok = N > 0 && gg::wgrad1x1_thin_eligible(a, b, N, A, HA, WA, B, HB, WB, KH, KW, stride, pad_y, pad_x);
    const bool is_auto = (prec == GG_PREC_AUTO || prec == GG_PREC_AUTO_FAST);
    int use = is_auto ? ((tc_ok && !thin_ok) ? (prec == GG_PREC_AUTO_FAST ? GG_PREC_TF32X1 : GG_PREC_TF32X3) : GG_PREC_FP32_SIMT) : prec;
    if (used_prec) *used_prec = use;
    if (use == GG_PREC_FP32_SIMT && thin_ok)
        return gg::wgrad1x1_thin(a, b, dw, N, A, HA, WA, B, out_layout, a_scale, b_scale, st);
    if (use == GG_PREC_FP32_SIMT)
        return gg::conv2d_wgrad_simt(a, b, dw, N, A, HA, WA, B, HB, WB, KH, KW, stride, pad_y, pad_x, flip_w, out_layout, a_scale,
                                     b_scale, st);
    return gg::wgrad_tc(a, b, dw, N, A, HA, WA, B, HB, WB, KH, KW, pad_y, pad_x, flip_w, out_layout, a_scale, b_scale, use, pm_dim, pm_dead, st);
}
