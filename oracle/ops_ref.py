"""CPU oracle for the StyleGAN2 conv hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Restates, in plain torch CPU ops (dtype-generic: fp32 for parity, fp64 for
identities), the algorithm of the reference's `impl='ref'` operators:

    upfirdn2d          DissimilarDomains/torch_utils/ops/upfirdn2d.py:43-125,179-219,292-404
    bias_act           DissimilarDomains/torch_utils/ops/bias_act.py:23-60,127-157
                       (+ the explicit gradient formulas of bias_act.cu:56-142)
    conv2d_resample    DissimilarDomains/torch_utils/ops/conv2d_resample.py:29-154
    fma                DissimilarDomains/torch_utils/ops/fma.py:15-58
    modulated_conv2d   DissimilarDomains/training/networks.py:591-668

Third-party arithmetic behind the reference that is NOT under /root/reference:
the dense contraction and the activations are PyTorch/ATen (oneDNN on CPU,
cuDNN on GPU); nothing pins a version (upstream asks for "PyTorch 1.7.1",
DissimilarDomains/Dockerfile:9 uses nvcr.io/nvidia/pytorch:20.12-py3); this
container has torch 2.11.0.  The oracle therefore calls the same ATen entry
points (`torch.nn.functional.conv2d/conv_transpose2d`) for the contraction and
restates everything around them.

Pinning: the reference ships no golden vectors or tests for this path
(SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF THE REFERENCE
ITSELF, executed in the authoring container by `tests/golden/make_golden.py`
(which imports /root/reference live, compares every function here against the
reference's own `impl='ref'` code on seeded inputs, and writes the fixtures
under tests/golden/).  `tests/test_oracle_golden.py` re-checks the oracle
against those committed fixtures on every run.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this package.  The product (ga-gan_b200/) never does.
"""
import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# Argument parsing -- integer work, must be bit-exact with the reference.


def parse_scaling(scaling):
    """upfirdn2d.py:43-50"""
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple)) and all(isinstance(v, int) for v in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def parse_padding(padding):
    """upfirdn2d.py:53-62"""
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple)) and all(isinstance(v, int) for v in padding)
    if len(padding) == 2:
        px, py = padding
        padding = [px, px, py, py]
    px0, px1, py0, py1 = padding
    return px0, px1, py0, py1


def filter_size(f):
    """upfirdn2d.py:65-76 -> (fw, fh)"""
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    return int(f.shape[-1]), int(f.shape[0])


def upfirdn2d_out_size(in_size, up, pad0, pad1, fsize, down):
    """upfirdn2d.cpp:32-33 (C integer division of a value that is >= 0 whenever the op is legal)."""
    return (in_size * up + pad0 + pad1 - fsize + down) // down


def setup_filter(f, normalize=True, flip_filter=False, gain=1, separable=None):
    """upfirdn2d.py:81-125"""
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2] and f.numel() > 0
    if f.ndim == 0:
        f = f[None]
    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = torch.outer(f, f)
    assert f.ndim == (1 if separable else 2)
    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    return f


# ----------------------------------------------------------------------------
# upfirdn2d


def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:179-219 restated as an explicit tap loop.

    y[n,c,oy,ox] = sum_{ky,kx} K[ky,kx] * xz[n,c, oy*downy + ky, ox*downx + kx]
    where xz is x zero-stuffed by `up` and padded/cropped by `padding`, and
    K = f*gain^(ndim/2), flipped unless flip_filter (the reference then runs a
    correlation, `:208-215`).  Separable f = two 1-D passes (`:213-215`).
    Differentiable by plain autograd to any order.
    """
    assert isinstance(x, torch.Tensor) and x.ndim == 4
    if f is None:
        f = torch.ones([1, 1], dtype=torch.float32)
    assert f.ndim in [1, 2] and f.dtype == torch.float32
    N, C, H, W = x.shape
    upx, upy = parse_scaling(up)
    downx, downy = parse_scaling(down)
    px0, px1, py0, py1 = parse_padding(padding)

    xz = x.new_zeros([N, C, H * upy, W * upx])
    xz[:, :, ::upy, ::upx] = x
    xz = F.pad(xz, [max(px0, 0), max(px1, 0), max(py0, 0), max(py1, 0)])
    xz = xz[:, :, max(-py0, 0): xz.shape[2] - max(-py1, 0), max(-px0, 0): xz.shape[3] - max(-px1, 0)]

    K = (f * (gain ** (f.ndim / 2))).to(x.dtype)
    if not flip_filter:
        K = K.flip(list(range(K.ndim)))

    def corr(t, k2):  # valid correlation with a [kh,kw] kernel, tap by tap
        kh, kw = k2.shape
        oh, ow = t.shape[2] - kh + 1, t.shape[3] - kw + 1
        acc = None
        for ky in range(kh):
            for kx in range(kw):
                term = t[:, :, ky:ky + oh, kx:kx + ow] * k2[ky, kx]
                acc = term if acc is None else acc + term
        return acc

    if K.ndim == 2:
        y = corr(xz, K)
    else:
        y = corr(xz, K[None, :])
        y = corr(y, K[:, None])
    return y[:, :, ::downy, ::downx]


def filter2d(x, f, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:292-324"""
    px0, px1, py0, py1 = parse_padding(padding)
    fw, fh = filter_size(f)
    p = [px0 + fw // 2, px1 + (fw - 1) // 2, py0 + fh // 2, py1 + (fh - 1) // 2]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:329-364"""
    upx, upy = parse_scaling(up)
    px0, px1, py0, py1 = parse_padding(padding)
    fw, fh = filter_size(f)
    p = [px0 + (fw + upx - 1) // 2, px1 + (fw - upx) // 2, py0 + (fh + upy - 1) // 2, py1 + (fh - upy) // 2]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1):
    """upfirdn2d.py:369-404"""
    downx, downy = parse_scaling(down)
    px0, px1, py0, py1 = parse_padding(padding)
    fw, fh = filter_size(f)
    p = [px0 + (fw - downx + 1) // 2, px1 + (fw - downx) // 2, py0 + (fh - downy + 1) // 2, py1 + (fh - downy) // 2]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain)


def upfirdn2d_backward_args(x_shape, dy_shape, f, up, down, padding):
    """upfirdn2d.py:264-283: the (up, down, padding) of the op that maps dy -> dx (flip is negated, gain kept)."""
    upx, upy = parse_scaling(up)
    downx, downy = parse_scaling(down)
    px0, px1, py0, py1 = parse_padding(padding)
    _, _, ih, iw = x_shape
    _, _, oh, ow = dy_shape
    fw, fh = filter_size(f)
    p = [fw - px0 - 1, iw * upx - ow * downx + px0 - upx + 1, fh - py0 - 1, ih * upy - oh * downy + py0 - upy + 1]
    return dict(up=[downx, downy], down=[upx, upy], padding=p)


# ----------------------------------------------------------------------------
# bias_act

_SELU_SCALE = 1.0507009873554804934193349852946
_SELU_ALPHA = 1.6732632423543772848170429916717

# name -> (def_alpha, def_gain, cuda_idx, ref, has_2nd_grad)    bias_act.py:23-60
ACTIVATIONS = {
    'linear':   (0,   1,          1, '',  False),
    'relu':     (0,   np.sqrt(2), 2, 'y', False),
    'lrelu':    (0.2, np.sqrt(2), 3, 'y', False),
    'tanh':     (0,   1,          4, 'y', True),
    'sigmoid':  (0,   1,          5, 'y', True),
    'elu':      (0,   1,          6, 'y', True),
    'selu':     (0,   1,          7, 'y', True),
    'softplus': (0,   1,          8, 'y', True),
    'swish':    (0,   np.sqrt(2), 9, 'x', True),
}


def _act(x, act, alpha):
    if act == 'linear':
        return x
    if act == 'relu':
        return torch.relu(x)
    if act == 'lrelu':
        return F.leaky_relu(x, alpha)
    if act == 'tanh':
        return torch.tanh(x)
    if act == 'sigmoid':
        return torch.sigmoid(x)
    if act == 'elu':
        return F.elu(x)
    if act == 'selu':
        return F.selu(x)
    if act == 'softplus':
        return F.softplus(x)
    if act == 'swish':
        return torch.sigmoid(x) * x
    raise KeyError(act)


def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """bias_act.py:127-157: y = clamp(act(x + b) * gain)."""
    def_alpha, def_gain = ACTIVATIONS[act][0], ACTIVATIONS[act][1]
    alpha = float(alpha if alpha is not None else def_alpha)
    gain = float(gain if gain is not None else def_gain)
    clamp = float(clamp if clamp is not None else -1)
    if b is not None:
        assert b.ndim == 1 and 0 <= dim < x.ndim and b.shape[0] == x.shape[dim]
        x = x + b.reshape([-1 if i == dim else 1 for i in range(x.ndim)])
    x = _act(x, act, alpha)
    if gain != 1:
        x = x * gain
    if clamp >= 0:
        x = x.clamp(-clamp, clamp)
    return x


def bias_act_grad_formula(grad, act, x_in, xref_plus_b, yref, dy, alpha, gain, clamp):
    """The native kernel's explicit gradient formulas, bias_act.cu:40-142.

    grad=1: returns d/dx of the forward applied to `x_in` (= incoming dy).
    grad=2: returns the second-order term (only has_2nd_grad activations).
    `yref` is the saved forward OUTPUT (post gain/clamp); `xref_plus_b` the saved
    pre-activation (only swish needs it); `dy` multiplies the result (grad=2).
    Used to check the CUDA grad kernels independently of autograd.
    """
    x = x_in
    yy = yref / gain if gain != 0 else torch.zeros_like(x)
    one = 1.0
    if act == 'linear':
        y = x
    elif act == 'relu':
        y = torch.where(yy > 0, x, torch.zeros_like(x))
    elif act == 'lrelu':
        y = torch.where(yy > 0, x, x * alpha)
    elif act == 'tanh':
        y = x * (one - yy * yy) if grad == 1 else x * (one - yy * yy) * (-2.0 * yy)
    elif act == 'sigmoid':
        y = x * yy * (one - yy) if grad == 1 else x * yy * (one - yy) * (one - 2.0 * yy)
    elif act == 'elu':
        y = torch.where(yy >= 0, x if grad == 1 else torch.zeros_like(x), x * (yy + one))
    elif act == 'selu':
        sa = _SELU_SCALE * _SELU_ALPHA
        y = torch.where(yy >= 0, x * _SELU_SCALE if grad == 1 else torch.zeros_like(x), x * (yy + sa))
    elif act == 'softplus':
        c = torch.exp(-yy)
        y = x * (one - c) if grad == 1 else x * c * (one - c)
    elif act == 'swish':
        xr = xref_plus_b
        c = torch.exp(xr)
        d = c + one
        if grad == 1:
            y = torch.where(xr > 40, x, x * c * (xr + d) / (d * d))
        else:
            y = torch.where(xr > 40, torch.zeros_like(x), x * c * (xr * (2.0 - d) + 2.0 * d) / (d * d * d))
        yref = torch.where(xr < -80, torch.zeros_like(xr), xr / (torch.exp(-xr) + one) * gain)
    else:
        raise KeyError(act)
    y = y * gain * (dy if dy is not None else 1.0)
    if clamp >= 0:
        y = torch.where((yref > -clamp) & (yref < clamp), y, torch.zeros_like(y))
    return y


# ----------------------------------------------------------------------------
# fma


def fma(a, b, c):
    """fma.py:15-16: a*b+c with broadcasting (torch.addcmul)."""
    return torch.addcmul(c, a, b)


# ----------------------------------------------------------------------------
# conv2d_resample


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_resample.py:29-54 (the channels_last 1x1 workaround `:40-50` is fp16-only and layout-only)."""
    if not flip_weight:
        w = w.flip([2, 3])
    op = F.conv_transpose2d if transpose else F.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def conv2d_resample_plan(w_shape, f, up, down, padding, groups=1):
    """The integer bookkeeping of conv2d_resample.py:86-154 as data: which branch, with which pads.

    Returns a dict describing the exact op sequence; compared bit-for-bit with the
    product's planner in the tests.
    """
    out_channels, in_channels_per_group, kh, kw = [int(v) for v in w_shape]
    fw, fh = filter_size(f)
    px0, px1, py0, py1 = parse_padding(padding)
    if up > 1:
        px0 += (fw + up - 1) // 2
        px1 += (fw - up) // 2
        py0 += (fh + up - 1) // 2
        py1 += (fh - up) // 2
    if down > 1:
        px0 += (fw - down + 1) // 2
        px1 += (fw - down) // 2
        py0 += (fh - down + 1) // 2
        py1 += (fh - down) // 2
    if kw == 1 and kh == 1 and (down > 1 and up == 1):
        return dict(branch='down_1x1', fir_pad=[px0, px1, py0, py1])
    if kw == 1 and kh == 1 and (up > 1 and down == 1):
        return dict(branch='up_1x1', fir_pad=[px0, px1, py0, py1])
    if down > 1 and up == 1:
        return dict(branch='down', fir_pad=[px0, px1, py0, py1])
    if up > 1:
        px0 -= kw - 1
        px1 -= kw - up
        py0 -= kh - 1
        py1 -= kh - up
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        return dict(branch='up', conv_pad=[pyt, pxt], fir_pad=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt])
    if up == 1 and down == 1 and px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return dict(branch='plain', conv_pad=[py0, px0])
    return dict(branch='generic', fir_pad=[px0, px1, py0, py1])


def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    """conv2d_resample.py:59-154"""
    assert x.ndim == 4 and w.ndim == 4 and w.dtype == x.dtype
    out_channels, in_channels_per_group, kh, kw = [int(v) for v in w.shape]
    plan = conv2d_resample_plan(w.shape, f, up, down, padding, groups)
    br = plan['branch']
    if br == 'down_1x1':
        x = upfirdn2d(x, f, down=down, padding=plan['fir_pad'], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
    if br == 'up_1x1':
        x = _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d(x, f, up=up, padding=plan['fir_pad'], gain=up ** 2, flip_filter=flip_filter)
    if br == 'down':
        x = upfirdn2d(x, f, padding=plan['fir_pad'], flip_filter=flip_filter)
        return _conv2d_wrapper(x, w, stride=down, groups=groups, flip_weight=flip_weight)
    if br == 'up':
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, out_channels // groups, in_channels_per_group, kh, kw)
            w = w.transpose(1, 2)
            w = w.reshape(groups * in_channels_per_group, out_channels // groups, kh, kw)
        x = _conv2d_wrapper(x, w, stride=up, padding=plan['conv_pad'], groups=groups, transpose=True,
                            flip_weight=(not flip_weight))
        x = upfirdn2d(x, f, padding=plan['fir_pad'], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
        return x
    if br == 'plain':
        return _conv2d_wrapper(x, w, padding=plan['conv_pad'], groups=groups, flip_weight=flip_weight)
    x = upfirdn2d(x, (f if up > 1 else None), up=up, padding=plan['fir_pad'], gain=up ** 2, flip_filter=flip_filter)
    x = _conv2d_wrapper(x, w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d(x, f, down=down, flip_filter=flip_filter)
    return x


# ----------------------------------------------------------------------------
# modulated_conv2d


def modulated_conv2d(x, weight, styles, noise=None, up=1, down=1, padding=0, resample_filter=None,
                     demodulate=True, flip_weight=True, fused_modconv=True):
    """networks.py:591-668 (fp32 path; the fp16 pre-normalisation `:622-627` is out of scope)."""
    N = x.shape[0]
    O, I, kh, kw = weight.shape
    assert x.shape[1] == I and tuple(styles.shape) == (N, I)
    w = None
    dcoefs = None
    if demodulate or fused_modconv:
        w = weight.unsqueeze(0) * styles.reshape(N, 1, -1, 1, 1)
    if demodulate:
        dcoefs = (w.square().sum(dim=[2, 3, 4]) + 1e-8).rsqrt()
    if demodulate and fused_modconv:
        w = w * dcoefs.reshape(N, -1, 1, 1, 1)
    if not fused_modconv:
        x = x * styles.reshape(N, -1, 1, 1)
        x = conv2d_resample(x, weight, f=resample_filter, up=up, down=down, padding=padding, flip_weight=flip_weight)
        if demodulate and noise is not None:
            x = fma(x, dcoefs.reshape(N, -1, 1, 1), noise)
        elif demodulate:
            x = x * dcoefs.reshape(N, -1, 1, 1)
        elif noise is not None:
            x = x + noise
        return x
    x = x.reshape(1, -1, *x.shape[2:])
    w = w.reshape(-1, I, kh, kw)
    x = conv2d_resample(x, w, f=resample_filter, up=up, down=down, padding=padding, groups=N, flip_weight=flip_weight)
    x = x.reshape(N, -1, *x.shape[2:])
    if noise is not None:
        x = x + noise
    return x


# ----------------------------------------------------------------------------
# Error metrics used by every parity test (BASELINE.json north_star).


def max_rel_err(a, b):
    """max|a-b| / max|b|  -- the north star's "max relative error"."""
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    denom = b.abs().max().item()
    if denom == 0:
        return float((a - b).abs().max().item())
    return float((a - b).abs().max().item() / denom)
