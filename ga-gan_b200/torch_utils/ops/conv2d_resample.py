"""2D convolution with optional up/downsampling.

API of the reference's `torch_utils/ops/conv2d_resample.py:59-154`.  `plan()` restates the
integer pad bookkeeping and branch selection (:86-154) as data so that it can be checked
bit-for-bit against the oracle; `conv2d_resample()` executes the plan on this build's
`upfirdn2d` and `conv2d_gradfix` ops.
"""
import torch

from .. import misc
from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _parse_padding
from .upfirdn2d import _get_filter_size


def _get_weight_shape(w):
    shape = [int(sz) for sz in w.shape]
    misc.assert_shape(w, shape)
    return shape


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_resample.py:29-54.  `flip_weight=False` means true convolution: flip the kernel first.
    (The channels_last 1x1 cuDNN workaround of :40-50 is fp16-era and has no equivalent here.)"""
    if not flip_weight:
        w = w.flip([2, 3])
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


def plan(w_shape, f, up, down, padding):
    """Branch + pads chosen by conv2d_resample.py:86-154 for a weight of shape `w_shape`."""
    _, _, kh, kw = [int(v) for v in w_shape]
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = _parse_padding(padding)
    if up > 1:
        px0, px1 = px0 + (fw + up - 1) // 2, px1 + (fw - up) // 2
        py0, py1 = py0 + (fh + up - 1) // 2, py1 + (fh - up) // 2
    if down > 1:
        px0, px1 = px0 + (fw - down + 1) // 2, px1 + (fw - down) // 2
        py0, py1 = py0 + (fh - down + 1) // 2, py1 + (fh - down) // 2
    one_by_one = (kw == 1 and kh == 1)
    if one_by_one and down > 1 and up == 1:
        return dict(branch='down_1x1', fir_pad=[px0, px1, py0, py1])
    if one_by_one and up > 1 and down == 1:
        return dict(branch='up_1x1', fir_pad=[px0, px1, py0, py1])
    if down > 1 and up == 1:
        return dict(branch='down', fir_pad=[px0, px1, py0, py1])
    if up > 1:
        px0, px1, py0, py1 = px0 - (kw - 1), px1 - (kw - up), py0 - (kh - 1), py1 - (kh - up)
        pxt = max(min(-px0, -px1), 0)
        pyt = max(min(-py0, -py1), 0)
        return dict(branch='up', conv_pad=[pyt, pxt], fir_pad=[px0 + pxt, px1 + pxt, py0 + pyt, py1 + pyt])
    if px0 == px1 and py0 == py1 and px0 >= 0 and py0 >= 0:
        return dict(branch='plain', conv_pad=[py0, px0])
    return dict(branch='generic', fir_pad=[px0, px1, py0, py1])


@misc.profiled_function
def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False):
    r"""2D convolution with optional up/downsampling; padding is applied once, up front.

    x `[N, I, H, W]`, w `[O, I//groups, kh, kw]`, f from `upfirdn2d.setup_filter()` or None.
    `flip_weight=True` = correlation (what `F.conv2d` does), False = convolution.
    """
    assert isinstance(x, torch.Tensor) and (x.ndim == 4)
    assert isinstance(w, torch.Tensor) and (w.ndim == 4) and (w.dtype == x.dtype)
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and (up >= 1)
    assert isinstance(down, int) and (down >= 1)
    assert isinstance(groups, int) and (groups >= 1)
    out_channels, in_channels_per_group, kh, kw = _get_weight_shape(w)
    pl = plan(w.shape, f, up, down, padding)
    branch = pl['branch']

    if branch == 'down_1x1':      # FIR-decimate first, then the 1x1 conv on the small image
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=pl['fir_pad'], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)

    if branch == 'up_1x1':        # 1x1 conv on the small image, then zero-stuff + FIR
        x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x=x, f=f, up=up, padding=pl['fir_pad'], gain=up ** 2, flip_filter=flip_filter)

    if branch == 'down':          # FIR at full resolution, then a strided conv
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=pl['fir_pad'], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, stride=down, groups=groups, flip_weight=flip_weight)

    if branch == 'up':            # transposed strided conv, then FIR (then optional decimation)
        if groups == 1:
            w = w.transpose(0, 1)
        else:
            w = w.reshape(groups, out_channels // groups, in_channels_per_group, kh, kw).transpose(1, 2)
            w = w.reshape(groups * in_channels_per_group, out_channels // groups, kh, kw)
        x = _conv2d_wrapper(x=x, w=w, stride=up, padding=pl['conv_pad'], groups=groups, transpose=True,
                            flip_weight=(not flip_weight))
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=pl['fir_pad'], gain=up ** 2, flip_filter=flip_filter)
        if down > 1:
            x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
        return x

    if branch == 'plain':
        return _conv2d_wrapper(x=x, w=w, padding=pl['conv_pad'], groups=groups, flip_weight=flip_weight)

    # generic: resample with upfirdn2d around an unpadded conv
    x = upfirdn2d.upfirdn2d(x=x, f=(f if up > 1 else None), up=up, padding=pl['fir_pad'], gain=up ** 2, flip_filter=flip_filter)
    x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
    if down > 1:
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, flip_filter=flip_filter)
    return x
