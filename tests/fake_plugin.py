"""A torch (CPU, any dtype) stand-in for the conv2d_plugin object, for testing the AUTOGRAD ALGEBRA of
ga-gan_b200/torch_utils/ops/conv2d_gradfix.py without a GPU: same call signatures and semantics as custom_ops._Plugin
(include/gagan_b200.h: gg_conv2d_f32 with stride-1 free-extent outputs, gg_conv2d_wgrad_f32, gg_chan_dot_f32, gg_scale_rows_f32,
gg_axpby_rows_f32), written with plain torch ops.  Test infrastructure only: nothing in the product imports it."""
import torch
import torch.nn.functional as F


def _bc(s):
    return s[:, :, None, None]


class FakePlugin:
    last_conv_prec = 0
    last_wgrad_prec = 0

    def __init__(self, permute_seed=None):
        """permute_seed: sum the contracted channels of every convolution in a (seeded) permuted order -- the same mathematics with
        another fp32 rounding, i.e. what a second correct implementation (the GPU kernels) looks like to a test.  Used to calibrate
        the bounds of GPU tests that cannot be run where they are written (leaky-ReLU sign flips, ill-conditioned R1 gradients)."""
        self.permute_seed = permute_seed

    def conv2d(self, x, w, stride=1, padding=(0, 0), transposed=False, output_padding=(0, 0), flip_w=False, in_scale=None, out_scale=None,
               prec=None, out_hw=None, flop_scale=1.0, epilogue=None):
        assert epilogue is None
        if self.permute_seed is not None:
            perm = torch.randperm(int(x.shape[1]), generator=torch.Generator().manual_seed(self.permute_seed + int(x.shape[1])))
            x, w = x[:, perm].contiguous(), (w[perm] if transposed else w[:, perm]).contiguous()
            in_scale = in_scale[:, perm] if in_scale is not None else None
        if stride != 1:                      # the generic strided route (conv2d_gradfix._conv2d_gradfix): no scales, no free extents
            assert in_scale is None and out_scale is None and out_hw is None
            v = w.flip([2, 3]) if flip_w else w
            return F.conv_transpose2d(x, v, stride=stride, padding=padding, output_padding=output_padding) if transposed \
                else F.conv2d(x, v, stride=stride, padding=padding)
        kh, kw = int(w.shape[2]), int(w.shape[3])
        if transposed:                       # conv_transpose2d(x, w[I,O], padding=p) == correlation with w^T flipped, padding k-1-p
            v, fl, py, px = w.transpose(0, 1), not flip_w, kh - 1 - padding[0], kw - 1 - padding[1]
        else:
            v, fl, py, px = w, flip_w, padding[0], padding[1]
        if fl:
            v = v.flip([2, 3])
        xs = x * _bc(in_scale) if in_scale is not None else x
        H, W = int(x.shape[2]), int(x.shape[3])
        OH, OW = (int(out_hw[0]), int(out_hw[1])) if out_hw is not None else (H + 2 * py - kh + 1, W + 2 * px - kw + 1)
        pb, pr = OH + kh - 1 - py - H, OW + kw - 1 - px - W          # rows / columns of zeros (or cropping, if negative) below / right
        y = F.conv2d(F.pad(xs, (px, pr, py, pb)), v)
        assert tuple(y.shape[2:]) == (OH, OW)
        return y * _bc(out_scale) if out_scale is not None else y

    def conv2d_wgrad(self, a, b, kernel_size, stride=1, padding=(0, 0), flip_w=False, out_layout=0, a_scale=None, b_scale=None, prec=None,
                     flop_scale=1.0, pm=None):
        kh, kw = kernel_size
        s_ = int(stride)
        a_s = a * _bc(a_scale) if a_scale is not None else a
        b_s = b * _bc(b_scale) if b_scale is not None else b
        HA, WA, HB, WB = int(a.shape[2]), int(a.shape[3]), int(b.shape[2]), int(b.shape[3])
        # dw[o,i,ky,kx] = sum b[n,o,oy,ox] * a[n,i, oy*stride - pad + ky, ox*stride - pad + kx]   (include/gagan_b200.h)
        need_h, need_w = (HB - 1) * s_ + kh, (WB - 1) * s_ + kw
        ap = F.pad(a_s, (padding[1], max(need_w - padding[1] - WA, 0), padding[0], max(need_h - padding[0] - HA, 0)))
        if s_ == 1:
            ap = F.pad(a_s, (padding[1], WB + kw - 1 - padding[1] - WA, padding[0], HB + kh - 1 - padding[0] - HA))
        dw = torch.stack([torch.stack([torch.einsum('noyx,niyx->oi', b_s, ap[:, :, ky:ky + (HB - 1) * s_ + 1:s_, kx:kx + (WB - 1) * s_ + 1:s_])
                                       for kx in range(kw)], dim=-1) for ky in range(kh)], dim=-2)              # [B, A, kh, kw]
        if flip_w:
            dw = dw.flip([2, 3])
        return dw.transpose(0, 1).contiguous() if out_layout else dw

    def chan_dot(self, a, b):
        return (a * b).flatten(2).sum(-1)

    def scale_rows(self, x, s):
        return x * _bc(s)

    def fma_rows(self, x, s, z):
        N, C, H, W = x.shape
        return x * _bc(s) + (z.reshape(1, 1, H, W) if z.numel() == H * W else z.reshape(N, 1, H, W))

    def axpby_rows(self, x1, s1, x2, s2):
        return x1 * _bc(s1) + x2 * _bc(s2)

    # ---- upfirdn2d_plugin / fir4_pm (include/gagan_b200.h: gg_upfirdn2d_f32, gg_fir4_pm_f32)

    @staticmethod
    def _pad_crop(u, x0, x1, y0, y1):
        u = F.pad(u, (max(x0, 0), max(x1, 0), max(y0, 0), max(y1, 0)))
        return u[:, :, max(-y0, 0): u.shape[2] - max(-y1, 0), max(-x0, 0): u.shape[3] - max(-x1, 0)]

    @staticmethod
    def _fir(u, f, flip, gain):
        ff = (f * gain).to(u.dtype)
        if not flip:
            ff = ff.flip([0, 1])
        N, C, H, W = u.shape
        return F.conv2d(u.reshape(N * C, 1, H, W), ff[None, None]).reshape(N, C, H - ff.shape[0] + 1, W - ff.shape[1] + 1)

    def upfirdn2d(self, x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        N, C, H, W = x.shape
        u = F.pad(x.reshape(N, C, H, 1, W, 1), (0, upx - 1, 0, 0, 0, upy - 1)).reshape(N, C, H * upy, W * upx)
        y = self._fir(self._pad_crop(u, padx0, padx1, pady0, pady1), f, flip, gain)
        return y[:, :, ::downy, ::downx].clone()            # (a fresh tensor, like the kernel's output: callers modify it in place)

    def fir4_pm(self, x, f, padx0, pady0, flip, gain, in_hw, out_hw, in_pm=None, out_pm=None):
        from torch_utils.ops import upfirdn2d as U
        t = U.depth_to_space(x)[:, :, :in_hw[0], :in_hw[1]] if in_pm is not None else x
        H, W = int(in_hw[0]), int(in_hw[1])
        y = self._fir(self._pad_crop(t, padx0, out_hw[1] + 3 - W - padx0, pady0, out_hw[0] + 3 - H - pady0), f, flip, gain)
        assert tuple(y.shape[2:]) == (int(out_hw[0]), int(out_hw[1]))
        return (U.space_to_depth(y, out_pm[0], out_pm[1]) if out_pm is not None else y).clone()

    # ---- bias_act_plugin (include/gagan_b200.h: gg_bias_act_f32 incl. the fused bias gradient, gg_bias_act_noise_f32)
    _ACTS = {1: 'linear', 2: 'relu', 3: 'lrelu', 4: 'tanh', 5: 'sigmoid', 6: 'elu', 7: 'selu', 8: 'softplus', 9: 'swish'}

    def bias_act(self, x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp, dbias=None):
        from oracle import ops_ref as R                     # (tests may use the oracle; the product never does)
        name = self._ACTS[int(act)]
        shape = [-1 if i == dim else 1 for i in range(x.ndim)]
        bb = b.reshape(shape) if b.numel() else None
        if grad == 0:
            return R.bias_act(x, b if b.numel() else None, dim=dim, act=name, alpha=alpha, gain=gain, clamp=(clamp if clamp >= 0 else None))
        xr = None
        if xref.numel():
            xr = xref + bb if bb is not None else xref       # bias_act.cu: the bias joins xref when grad > 0
        y = R.bias_act_grad_formula(grad, name, x, xr, yref if yref.numel() else torch.zeros_like(x), dy if dy.numel() else None,
                                    alpha, gain, clamp)
        if dbias is not None:
            dbias += y.sum([i for i in range(x.ndim) if i != dim])
        return y

    def bias_act_noise(self, x, b, noise, act, alpha, gain, clamp):
        null = torch.empty([0], dtype=x.dtype)
        n = noise.reshape(1, 1, *noise.shape) if noise.ndim == 2 else noise
        return self.bias_act(x + n, b, null, null, null, 0, 1, act, alpha, gain, clamp)
