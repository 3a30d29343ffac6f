/*
 * gagan_b200.h -- C ABI of the B200-native StyleGAN2 conv hot path (libgagan_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry
 * point names the reference interface it replaces (paths relative to
 * /root/reference/DissimilarDomains).  All device pointers are fp32, dense, NCHW unless
 * stated; all kernels are enqueued on `stream` (the caller's current CUDA stream, as the
 * reference does with at::cuda::getCurrentCUDAStream(), upfirdn2d.cpp:92 / bias_act.cpp:88)
 * on the CURRENT device and never synchronise.  Inputs are borrowed for the duration of
 * the enqueued work; outputs are caller-allocated (torch's caching allocator in the
 * Python host).  The library keeps no per-call state and is re-entrant (autograd worker
 * threads call it concurrently during backward).
 *
 * Return value: 0 on success, negative GG_E* on failure; gg_last_error() returns the
 * message of the last failure on the calling thread (the reference raises through
 * TORCH_CHECK; the Python host turns a non-zero code into RuntimeError).
 * There is no CPU fallback anywhere behind this header.
 */
#ifndef GAGAN_B200_H_
#define GAGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* gg_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define GG_API __attribute__((visibility("default")))
#else
#define GG_API
#endif

enum {
    GG_OK = 0,
    GG_EINVAL = -1,      /* bad argument (the reference's TORCH_CHECK failures) */
    GG_ECUDA = -2,       /* CUDA runtime / launch error */
    GG_EUNSUPPORTED = -3 /* legal in the reference, not served by this build */
};

/* Conv arithmetic modes (the reference computes fp32 with TF32 off, training_loop.py:207-208). */
enum {
    GG_PREC_FP32_SIMT = 0, /* exact fp32 FFMA kernel (small / odd shapes, and the on-GPU cross-check) */
    GG_PREC_TF32X1 = 1,    /* tcgen05 kind::tf32, one product per MAC (fast mode, ~3e-4 per layer) */
    GG_PREC_TF32X3 = 3,    /* tcgen05 kind::tf32, hi/lo split, 3 products per MAC (~4e-7, the parity mode) */
    GG_PREC_AUTO = -1,     /* TF32X3 on tensor cores when the shape is eligible, else FP32_SIMT */
    GG_PREC_AUTO_FAST = -2 /* like AUTO but TF32X1 on the tensor cores: the fast mode, NOT fp32-faithful (bench.py fast_mode) */
};

GG_API const char* gg_last_error(void);
GG_API int gg_version(void);
/* 1 if the current device is sm_100 (the only target of this library), else 0. */
GG_API int gg_device_ok(void);

/* ------------------------------------------------------------------------------------------
 * bias_act -- replaces `_plugin.bias_act(x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp)`
 * (torch_utils/ops/bias_act.cpp:32-90, kernel bias_act.cu:23-147).
 *
 *   grad=0: y = clamp(act(x + b) * gain)
 *   grad=1: y = d/dx of the above applied to incoming gradient `x`, using saved `yref` (or `xref`
 *           for swish); zero where |yref| >= clamp.
 *   grad=2: second-order term, additionally multiplied by `dy` (has_2nd_grad activations only).
 *
 * `x`,`xref`,`yref`,`dy`,`y` hold sizeX elements each (xref/yref/dy/b may be NULL = absent, the
 * reference's empty tensor).  Bias index of element i is (i / stepB) % sizeB, exactly
 * bias_act.cu:44 (stepB = x.stride(dim)).  act = the reference's cuda_idx 1..9 (linear, relu,
 * lrelu, tanh, sigmoid, elu, selu, softplus, swish).  clamp < 0 disables clamping.
 *
 * Extension over the reference: if `dbias` is non-NULL (sizeB floats, grad>=1) the kernel also
 * accumulates dbias[c] += sum of y over all elements with bias index c -- the reference does this
 * as a second full pass `dx.sum(...)` (bias_act.py:211-212).  dbias must be zero-filled by the
 * caller (or hold a value to accumulate onto).
 */
GG_API int gg_bias_act_f32(const float* x, const float* b, const float* xref, const float* yref, const float* dy,
                    float* y, float* dbias, int grad, int act, float alpha, float gain, float clamp,
                    int64_t sizeX, int sizeB, int64_t stepB, gg_stream_t stream);

/* Forward bias_act with the SynthesisLayer's per-pixel noise folded in (training/networks.py:904-921 adds
 * `noise [N,1,H,W]` or `[H,W]` to the conv output with a separate full-tensor pass before bias_act):
 *   y[i] = clamp(act(x[i] + b[c(i)] + noise[n(i) * noise_batch_stride + i % stepB]) * gain),   n(i) = i / (stepB * sizeB)
 * x is dense NCHW with stepB = H*W and sizeB = C; noise_batch_stride = 0 broadcasts one plane over the batch. */
GG_API int gg_bias_act_noise_f32(const float* x, const float* b, const float* noise, int64_t noise_batch_stride, float* y, int act,
                          float alpha, float gain, float clamp, int64_t sizeX, int sizeB, int64_t stepB, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * upfirdn2d -- replaces `_plugin.upfirdn2d(x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1,
 * flip, gain)` (torch_utils/ops/upfirdn2d.cpp:16-94, kernels upfirdn2d.cu:29-200).
 *
 * x: [N,C,inH,inW] dense NCHW; f: [fH,fW] dense fp32; y: [N,C,outH,outW] dense NCHW with
 *   outW = (inW*upx + padx0 + padx1 - fW + downx) / downx   (upfirdn2d.cpp:32-33; checked here).
 * Negative padding crops.  flip=0 convolves (filter flipped), flip=1 correlates.
 */
GG_API int gg_upfirdn2d_f32(const float* x, const float* f, float* y, int N, int C, int inH, int inW, int fH, int fW,
                     int upx, int upy, int downx, int downy, int padx0, int padx1, int pady0, int pady1,
                     int flip, float gain, int outH, int outW, gg_stream_t stream);

/* 4x4 FIR at unit rate (up = down = 1) with an optional PHASE-MAJOR side -- the filter passes of the stride-2 layers
 * (conv2d_resample.py:119-122 FIR before the stride-2 conv, :139 FIR after the stride-2 transposed conv) fused with the
 * space-to-depth / depth-to-space re-layout that lets those convolutions run as stride-1 tensor-core GEMMs:
 *
 *   t_pm[n, (py,px,c), Y, X]  <->  t[n, c, 2Y+py, 2X+px]
 *
 *   y[n,c,oy,ox] = gain * sum_{ky,kx} F[ky][kx] * x[n,c, oy+ky-pady0, ox+kx-padx0]      (x == 0 outside inH x inW)
 *   F = f (flip=1) or f flipped in both axes (flip=0), exactly as gg_upfirdn2d_f32.
 *
 * in_pm=1 : x is [N,4C,in_pmH,in_pmW] phase-major, (inH,inW) is the valid logical extent (<= 2*in_pmH x 2*in_pmW).
 * out_pm=1: y is [N,4C,out_pmH,out_pmW] phase-major; logical positions outside (outH,outW) are written as zero.
 * Otherwise the side is plain [N,C,H,W].  At most one side is phase-major.
 */
GG_API int gg_fir4_pm_f32(const float* x, const float* f, float* y, int N, int C, int inH, int inW, int padx0, int pady0,
                   int flip, float gain, int outH, int outW, int in_pm, int in_pmH, int in_pmW, int out_pm, int out_pmH,
                   int out_pmW, gg_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * conv2d -- replaces the ATen/cuDNN calls behind `conv2d_gradfix.conv2d / conv_transpose2d`
 * (torch_utils/ops/conv2d_gradfix.py:37-58,138-148) and, with the optional per-sample scales, the
 * multiply passes around them in `modulated_conv2d` (training/networks.py:641-653):
 *
 *   y[n,o,:,:] = out_scale[n,o] * sum_i  corr_or_convT( in_scale[n,i] * x[n,i,:,:], w[o,i,:,:] )
 *
 * transposed=0: correlation with `stride`, zero padding (pad_y, pad_x); w is [O,I,KH,KW].
 * transposed=1: conv_transpose2d with `stride`, padding (pad_y,pad_x); w is [I,O,KH,KW]
 *               (torch layout); OH = (H-1)*stride - 2*pad_y + KH + output_padding (caller passes OH/OW).
 * stride == 1:  OH/OW are free -- (pad_y,pad_x) is the top/left padding and every position that reads outside
 *               x sees zeros, so a caller may crop or extend the output (this is what makes the op closed
 *               under differentiation without copies; conv2d_gradfix.conv2d_s1 in the Python host).
 * flip_w=1 uses w flipped in both spatial axes (conv2d_resample.py:35-36).
 * in_scale [N,I] / out_scale [N,O] may be NULL (plain conv, the discriminator).
 * groups == 1 only (grouped convs are split by the host).
 * prec: GG_PREC_*; AUTO picks the tcgen05 path when eligible.  `used_prec` (nullable) reports the choice.
 */
GG_API int gg_conv2d_f32(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW,
                  int OH, int OW, int stride, int pad_y, int pad_x, int transposed, int flip_w,
                  const float* in_scale, const float* out_scale, int prec, int* used_prec, gg_stream_t stream);

/* Weight gradient of the op above (replaces aten::cudnn_convolution_backward_weight,
 * conv2d_gradfix.py:175-191).  For the correlation y = conv(a, w):
 *   dw[o,i,ky,kx] = sum_{n,oy,ox} (b_scale[n,o]*b[n,o,oy,ox]) * (a_scale[n,i]*a[n,i, oy*stride-pad_y+ky, ox*stride-pad_x+kx])
 * a: [N,A,HA,WA] (the conv input), b: [N,B,HB,WB] (gradient w.r.t. the conv output), dw: [B,A,KH,KW].
 * out_layout=1 writes dw transposed as [A,B,KH,KW] (the conv_transpose2d weight layout, roles swapped
 * by the host).  flip_w=1 writes the spatially flipped gradient.  dw is overwritten.  Positions of `a` outside
 * its extent count as zeros (b may be smaller or larger than the natural correlation output).
 */
GG_API int gg_conv2d_wgrad_f32(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB,
                        int KH, int KW, int stride, int pad_y, int pad_x, int flip_w, int out_layout,
                        const float* a_scale, const float* b_scale, int prec, int* used_prec, gg_stream_t stream);

/* Same weight gradient with a structural hint for the phase-major stride-2 forms of conv2d_resample.py:119-142 (this repo's
 * torch_utils/ops/conv2d_resample.py: phase_major_weight_down / _up): dimension `pm_dim - 1` of dw (pm_dim = 1 or 2; 0 = no
 * hint) consists of 4 equal groups, one per sub-pixel phase (py,px), and bit (group*4 + ky*2 + kx) of `pm_dead` marks the taps of
 * that group that are zero BY CONSTRUCTION in the weight (7 of the 16 for a 3x3 kernel).  Those entries of dw may be left zero
 * instead of being computed: the host discards them.  KH == KW == 2 only; ignored otherwise. */
GG_API int gg_conv2d_wgrad_pm_f32(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB,
                           int KH, int KW, int stride, int pad_y, int pad_x, int flip_w, int out_layout,
                           const float* a_scale, const float* b_scale, int prec, int* used_prec, int pm_dim, unsigned pm_dead,
                           gg_stream_t stream);

/* out[r] = sum_p a[r,p] * b[r,p] for `rows` rows of P contiguous floats (row = one (sample, channel) plane): the style and
 * demodulation-coefficient gradients of modulated_conv2d once the per-sample scales live inside the conv kernels
 * (training/networks.py:642,648-651 and their autograd reductions).  out is overwritten. */
GG_API int gg_chan_dot_f32(const float* a, const float* b, float* out, int64_t rows, int64_t P, gg_stream_t stream);

/* gg_conv2d_f32 with the bias_act pass that follows it in the networks fused into the kernel's store loop:
 *     y = clamp(act(out_scale * conv(in_scale * x, w) + bias[o] + noise[pixel]) * gain)
 * i.e. `bias_act.bias_act(modulated_conv2d(..., noise=noise), b, act=..., gain=..., clamp=...)` of SynthesisLayer.forward
 * (training/networks.py:904-921) and `bias_act(conv2d_resample(...), b, ...)` of Conv2dLayer.forward (:752-760) as ONE launch
 * when the convolution is the last operator of the layer.  `bias` [O], `noise` ([OH*OW] with noise_batch_stride 0, or [N, OH*OW]
 * with noise_batch_stride OH*OW) may be NULL; act = the reference's cuda_idx.  linear / relu / lrelu are fused by the tcgen05
 * kernels (*fused = 1); for every other activation or kernel family the same result is produced by a second launch of the
 * bias_act kernel in place (*fused = 0).  stride-1 semantics and all other arguments as gg_conv2d_f32. */
GG_API int gg_conv2d_act_f32(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW,
                      int OH, int OW, int stride, int pad_y, int pad_x, int transposed, int flip_w,
                      const float* in_scale, const float* out_scale, const float* bias, const float* noise,
                      int64_t noise_batch_stride, int act, float alpha, float gain, float clamp, int prec, int* used_prec,
                      int* fused, gg_stream_t stream);

/* out[n,c] = sum_p ds[n,c,p] * (pre(y[n,c,p]) - bias[c] - noise[n,p]) with pre() the inverse of y = act(.) * gain (linear, or lrelu
 * with alpha != 0): the gradient of the [N,C] output scale (demodulation coefficients) of gg_conv2d_act_f32, whose pre-activation
 * tensor is never materialised.  ds, y: [N,C,P] dense, 16-byte aligned, P % 4 == 0. */
GG_API int gg_chan_dot_preact_f32(const float* ds, const float* y, const float* bias, const float* noise, int64_t noise_batch_stride,
                           float* out, int N, int C, int64_t P, int act, float alpha, float gain, gg_stream_t stream);

/* y[r,p] = s[r] * x[r,p] for `rows` rows of P contiguous floats (row = one (sample, channel) plane): the per-sample channel scale as a
 * tensor, needed by the SECOND-order gradients of modulated_conv2d (training/loss.py:96-109 path-length regularisation: the gradient of
 * a style / demodulation gradient w.r.t. the activations).  x and y may alias. */
GG_API int gg_scale_rows_f32(const float* x, const float* s, float* y, int64_t rows, int64_t P, gg_stream_t stream);

/* y[r,p] = s[r] * x[r,p] + z[(r / C) * z_batch_stride + p]  (rows = N*C planes of P contiguous floats): replaces the torch.addcmul
 * behind torch_utils/ops/fma.py:15-16,23 for the one shape the reference calls it with (training/networks.py:648:
 * fma(x [N,O,H,W], dcoefs [N,O,1,1], noise [N,1,H,W] or [H,W])).  z_batch_stride is 0 (one plane for the batch) or P. */
GG_API int gg_fma_rows_f32(const float* x, const float* s, const float* z, int64_t z_batch_stride, float* y, int64_t rows, int64_t C,
                           int64_t P, gg_stream_t stream);

/* y[r,p] = s1[r] * x1[r,p] + s2[r] * x2[r,p]: two such scaled tensors summed in one pass -- the two gradient contributions that meet in
 * front of a convolution in the second-order pass (through d/dx and through the style gradient d/da), so that one launch serves both. */
GG_API int gg_axpby_rows_f32(const float* x1, const float* s1, const float* x2, const float* s2, float* y, int64_t rows, int64_t P,
                      gg_stream_t stream);

/* Number of kernels this library has launched since load (all streams); bench.py reports the delta. */
GG_API int64_t gg_launch_count(void);

/* The tcgen05 kernels bound every mbarrier wait (~4 s of SM clocks); on expiry the waiting thread records what it was waiting for in
 * host-mapped memory and traps (the launch fails with a CUDA error instead of hanging the GPU).  Writes a one-line description of the
 * first expiry of this process into buf and returns 1, or returns 0 if no watchdog has fired. */
GG_API int gg_watchdog_report(char* buf, int len);

/* Which tcgen05 kernel family serves the 3x3 stride-1 layers with <= 64 output channels: 1 (default) = the row-marching kernel
 * (conv_march.cu: filter rows in the MMA N dimension), 0 = the tile kernel that serves every other shape (conv_tc.cu).  Both
 * compute the same fp32-faithful 3xTF32 result; the switch exists for A/B measurements (tools/microbench.py) and for the parity
 * tests that run every shape through both.  Process-wide; returns the previous setting. */
GG_API int gg_set_conv_kernel_family(int family);

#ifdef __cplusplus
}
#endif
#endif /* GAGAN_B200_H_ */
