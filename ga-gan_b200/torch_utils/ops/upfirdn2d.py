"""Pad / upsample / FIR-filter / downsample 2-D images on sm_100a.

Same API as the reference's `torch_utils/ops/upfirdn2d.py`: `setup_filter` (:81-125), `upfirdn2d`
(:130-174), `filter2d` (:292-324), `upsample2d` (:329-364), `downsample2d` (:369-404), the parsing
helpers that `conv2d_resample` imports (`_parse_padding` :53-62, `_get_filter_size` :65-76,
`_parse_scaling` :43-50) and the same autograd structure (`Upfirdn2dCuda` :237-283: the backward
of upfirdn2d is upfirdn2d with up<->down swapped, mirrored padding and the flipped filter, so
gradients of any order stay inside this one op).  The body is the C-ABI kernel family
`gg_upfirdn2d_f32` (include/gagan_b200.h); fp32 kernels (float16 images: `_util.fp16_storage`); no `impl='ref'`, no CPU fallback.
"""
import numpy as np
import torch

from .. import custom_ops
from ..._util import check_dims, fp16_storage

# ----------------------------------------------------------------------------

_plugin = None


def _init():
    """Bind the native library (upfirdn2d.py:28-40).  Raises instead of falling back."""
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('upfirdn2d_plugin', sources=['upfirdn2d.cu'])
    return True


def _parse_scaling(scaling):
    if isinstance(scaling, int):
        scaling = [scaling, scaling]
    assert isinstance(scaling, (list, tuple))
    assert all(isinstance(x, int) for x in scaling)
    sx, sy = scaling
    assert sx >= 1 and sy >= 1
    return sx, sy


def _parse_padding(padding):
    if isinstance(padding, int):
        padding = [padding, padding]
    assert isinstance(padding, (list, tuple))
    assert all(isinstance(x, int) for x in padding)
    if len(padding) == 2:
        padx, pady = padding
        padding = [padx, padx, pady, pady]
    padx0, padx1, pady0, pady1 = padding
    return padx0, padx1, pady0, pady1


def _get_filter_size(f):
    if f is None:
        return 1, 1
    assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
    fw = int(f.shape[-1])
    fh = int(f.shape[0])
    check_dims(f, [fh, fw][:f.ndim], 'filter')
    assert fw >= 1 and fh >= 1
    return fw, fh


# ----------------------------------------------------------------------------

def setup_filter(f, device=torch.device('cpu'), normalize=True, flip_filter=False, gain=1, separable=None):
    r"""Convenience function to setup 2D FIR filter for `upfirdn2d()` (upfirdn2d.py:81-125).

    Returns a float32 tensor `[filter_height, filter_width]` (non-separable) or `[filter_taps]`
    (separable; chosen automatically for 1-D filters with >= 8 taps).
    """
    if f is None:
        f = 1
    f = torch.as_tensor(f, dtype=torch.float32)
    assert f.ndim in [0, 1, 2]
    assert f.numel() > 0
    if f.ndim == 0:
        f = f[np.newaxis]

    if separable is None:
        separable = (f.ndim == 1 and f.numel() >= 8)
    if f.ndim == 1 and not separable:
        f = f.ger(f)
    assert f.ndim == (1 if separable else 2)

    if normalize:
        f = f / f.sum()
    if flip_filter:
        f = f.flip(list(range(f.ndim)))
    f = f * (gain ** (f.ndim / 2))
    f = f.to(device=device)
    return f


# ----------------------------------------------------------------------------

def _check_input(x):
    if x.device.type != 'cuda':
        raise RuntimeError('upfirdn2d: the B200 build has no CPU path; x must be a CUDA tensor')
    if x.dtype != torch.float32:                        # (float16 never arrives here: fp16_storage)
        raise RuntimeError('upfirdn2d: this build serves fp32 kernels (float16 images through them); other dtypes are out of scope')
    _init()


@fp16_storage('x')
def upfirdn2d(x, f, up=1, down=1, padding=0, flip_filter=False, gain=1, impl='cuda'):
    r"""Pad, upsample, filter, and downsample a batch of 2D images (upfirdn2d.py:130-174).

    1. zero-stuff by `up`; 2. pad (negative = crop) by `padding`; 3. convolve with `f`
    (`flip_filter=True` = correlate); 4. keep every `down`-th pixel.  Gradients of arbitrary order.
    float16 images: fp32 arithmetic, float16 result (`_util.fp16_storage`).
    """
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl != 'cuda':
        raise RuntimeError("upfirdn2d: impl='ref' is not shipped in the B200 build (see oracle/ops_ref.py)")
    _check_input(x)
    return _upfirdn2d_cuda(up=up, down=down, padding=padding, flip_filter=flip_filter, gain=gain).apply(x, f)


# ----------------------------------------------------------------------------

_upfirdn2d_cuda_cache = dict()


def _upfirdn2d_cuda(up=1, down=1, padding=0, flip_filter=False, gain=1):
    """Autograd Function factory (upfirdn2d.py:227-287), cached per argument tuple."""
    upx, upy = _parse_scaling(up)
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)

    key = (upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
    if key in _upfirdn2d_cuda_cache:
        return _upfirdn2d_cuda_cache[key]

    class Upfirdn2dCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):  # pylint: disable=arguments-differ
            assert isinstance(x, torch.Tensor) and x.ndim == 4
            if f is None:
                f = torch.ones([1, 1], dtype=torch.float32, device=x.device)
            assert isinstance(f, torch.Tensor) and f.ndim in [1, 2]
            y = x
            if f.ndim == 2:
                y = _plugin.upfirdn2d(y, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip_filter, gain)
            else:  # separable: two 1-D passes with sqrt(gain) each (upfirdn2d.py:251-259)
                y = _plugin.upfirdn2d(y, f.unsqueeze(0), upx, 1, downx, 1, padx0, padx1, 0, 0, flip_filter, np.sqrt(gain))
                y = _plugin.upfirdn2d(y, f.unsqueeze(1), 1, upy, 1, downy, 0, 0, pady0, pady1, flip_filter, np.sqrt(gain))
            ctx.save_for_backward(f)
            ctx.x_shape = x.shape
            return y

        @staticmethod
        def backward(ctx, dy):  # pylint: disable=arguments-differ
            f, = ctx.saved_tensors
            _, _, ih, iw = ctx.x_shape
            _, _, oh, ow = dy.shape
            fw, fh = _get_filter_size(f)
            p = [
                fw - padx0 - 1,
                iw * upx - ow * downx + padx0 - upx + 1,
                fh - pady0 - 1,
                ih * upy - oh * downy + pady0 - upy + 1,
            ]
            dx = None
            df = None
            if ctx.needs_input_grad[0]:
                dx = _upfirdn2d_cuda(up=[downx, downy], down=[upx, upy], padding=p, flip_filter=(not flip_filter),
                                     gain=gain).apply(dy, f)
            assert not ctx.needs_input_grad[1]
            return dx, df

    _upfirdn2d_cuda_cache[key] = Upfirdn2dCuda
    return Upfirdn2dCuda


# ----------------------------------------------------------------------------
# Unit-rate FIR fused with the phase-major re-layout of conv2d_resample's stride-2 paths.
#
#   t_pm[n, (py,px,c), Y, X]  <->  t[n, c, 2Y+py, 2X+px]
#
# fir_to_pm   : plain x -> upfirdn2d(x, f, padding) -> phase-major [N,4C,ys,xs] (zero outside the FIR output)
# fir_from_pm : phase-major z (valid logical extent `valid_hw`) -> upfirdn2d(..., padding) -> plain
# Each is the other's gradient (same rule as upfirdn2d.py:264-283: flipped filter, padding p), so gradients of any order
# exist.  Only 4x4 filters; other filters take the unfused route in conv2d_resample.

_fir_pm_cache = dict()


def space_to_depth(x, ys, xs):
    """[N,C,H,W] -> [N,4C,ys,xs] phase-major, zero-padded (or cropped) to 2ys x 2xs first."""
    N, C, H, W = x.shape
    x = torch.nn.functional.pad(x, (0, 2 * xs - W, 0, 2 * ys - H))
    return x.reshape(N, C, ys, 2, xs, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 4 * C, ys, xs)


def depth_to_space(z):
    """[N,4O,ys,xs] phase-major -> [N,O,2ys,2xs]."""
    N, C4, ys, xs = z.shape
    return z.reshape(N, 2, 2, C4 // 4, ys, xs).permute(0, 3, 4, 1, 5, 2).reshape(N, C4 // 4, 2 * ys, 2 * xs)


def _fused_fir_ok(f, px0, plain_w, pm_w):
    # the fused kernel wants a 2-D 4x4 filter, 0 <= padx0 <= 3 (also for its gradient, whose padx0 is 3 - padx0) and
    # 16-byte aligned rows on both sides (128-bit loads / stores)
    return f is not None and f.ndim == 2 and tuple(f.shape) == (4, 4) and 0 <= px0 <= 3 and plain_w % 4 == 0 and pm_w % 4 == 0


def fir_to_pm(x, f, padding, flip_filter, gain, ys, xs):
    _check_input(x)
    px0, px1, py0, py1 = _parse_padding(padding)
    if not _fused_fir_ok(f, px0, int(x.shape[3]), int(xs)):      # odd widths / other filters: the same result from the unfused ops
        return space_to_depth(upfirdn2d(x, f, padding=padding, flip_filter=flip_filter, gain=gain), ys, xs)
    oh, ow = x.shape[2] + py0 + py1 - 3, x.shape[3] + px0 + px1 - 3
    return _fir_pm(True, px0, py0, bool(flip_filter), float(gain), (int(x.shape[2]), int(x.shape[3])), (oh, ow), (int(ys), int(xs))).apply(x, f)


def fir_from_pm(z, f, padding, flip_filter, gain, valid_hw):
    _check_input(z)
    px0, px1, py0, py1 = _parse_padding(padding)
    vh, vw = int(valid_hw[0]), int(valid_hw[1])
    fw, fh = _get_filter_size(f)
    oh, ow = vh + py0 + py1 - fh + 1, vw + px0 + px1 - fw + 1
    if not _fused_fir_ok(f, px0, ow, int(z.shape[3])):
        full = depth_to_space(z)[:, :, :vh, :vw]
        return upfirdn2d(full, f, padding=padding, flip_filter=flip_filter, gain=gain)
    return _fir_pm(False, px0, py0, bool(flip_filter), float(gain), (vh, vw), (oh, ow), (int(z.shape[2]), int(z.shape[3]))).apply(z, f)


def _fir_pm(to_pm, px0, py0, flip, gain, in_hw, out_hw, pm_hw):
    """to_pm: plain [N,C,*in_hw] -> phase-major [N,4C,*pm_hw] with valid extent out_hw;
    else: phase-major [N,4C,*pm_hw] with valid extent in_hw -> plain [N,C,*out_hw]."""
    key = (to_pm, px0, py0, flip, gain, in_hw, out_hw, pm_hw)
    if key in _fir_pm_cache:
        return _fir_pm_cache[key]
    assert out_hw[0] >= 1 and out_hw[1] >= 1

    class FirPM(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, f):
            ctx.save_for_backward(f)
            if to_pm:
                return _plugin.fir4_pm(x, f, px0, py0, flip, gain, in_hw, out_hw, out_pm=pm_hw)
            return _plugin.fir4_pm(x, f, px0, py0, flip, gain, in_hw, out_hw, in_pm=pm_hw)

        @staticmethod
        def backward(ctx, dy):
            f, = ctx.saved_tensors
            dx = None
            if ctx.needs_input_grad[0]:
                # adjoint of a unit-rate FIR: the same FIR with the filter flipped and padding (3 - p0) on the leading side;
                # its input extent is this op's output extent and vice versa (upfirdn2d.py:270-275 with up = down = 1)
                dx = _fir_pm(not to_pm, 3 - px0, 3 - py0, not flip, gain, out_hw, in_hw, pm_hw).apply(dy, f)
            return dx, None

    _fir_pm_cache[key] = FirPM
    return FirPM


# ----------------------------------------------------------------------------

def filter2d(x, f, padding=0, flip_filter=False, gain=1, impl='cuda'):
    r"""Filter a batch of 2D images; output shape == input shape (upfirdn2d.py:292-324)."""
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [
        padx0 + fw // 2,
        padx1 + (fw - 1) // 2,
        pady0 + fh // 2,
        pady1 + (fh - 1) // 2,
    ]
    return upfirdn2d(x, f, padding=p, flip_filter=flip_filter, gain=gain, impl=impl)


def upsample2d(x, f, up=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    r"""Upsample a batch of 2D images; output shape == input shape * up (upfirdn2d.py:329-364)."""
    upx, upy = _parse_scaling(up)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [
        padx0 + (fw + upx - 1) // 2,
        padx1 + (fw - upx) // 2,
        pady0 + (fh + upy - 1) // 2,
        pady1 + (fh - upy) // 2,
    ]
    return upfirdn2d(x, f, up=up, padding=p, flip_filter=flip_filter, gain=gain * upx * upy, impl=impl)


def downsample2d(x, f, down=2, padding=0, flip_filter=False, gain=1, impl='cuda'):
    r"""Downsample a batch of 2D images; output shape == input shape / down (upfirdn2d.py:369-404)."""
    downx, downy = _parse_scaling(down)
    padx0, padx1, pady0, pady1 = _parse_padding(padding)
    fw, fh = _get_filter_size(f)
    p = [
        padx0 + (fw - downx + 1) // 2,
        padx1 + (fw - downx) // 2,
        pady0 + (fh - downy + 1) // 2,
        pady1 + (fh - downy) // 2,
    ]
    return upfirdn2d(x, f, down=down, padding=p, flip_filter=flip_filter, gain=gain * 1, impl=impl)

# ----------------------------------------------------------------------------
