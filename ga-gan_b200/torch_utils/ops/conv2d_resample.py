"""2D convolution with optional up/downsampling.

API of the reference's `torch_utils/ops/conv2d_resample.py:59-154`.  `plan()` restates the
integer pad bookkeeping and branch selection (:86-154) as data so that it can be checked
bit-for-bit against the oracle; `conv2d_resample()` executes the plan on this build's
`upfirdn2d` and `conv2d_gradfix` ops.
"""
import torch

from ..._util import check_dims, scoped, fp16_storage
from . import conv2d_gradfix
from . import upfirdn2d
from .upfirdn2d import _parse_padding
from .upfirdn2d import _get_filter_size


def _get_weight_shape(w):
    shape = [int(sz) for sz in w.shape]
    check_dims(w, shape, 'weight')
    return shape


def _conv2d_wrapper(x, w, stride=1, padding=0, groups=1, transpose=False, flip_weight=True):
    """conv2d_resample.py:29-54.  `flip_weight=False` means true convolution: flip the kernel first.
    (The channels_last 1x1 cuDNN workaround of :40-50 is fp16-era and has no equivalent here.)"""
    if not flip_weight:
        w = w.flip([2, 3])
    op = conv2d_gradfix.conv_transpose2d if transpose else conv2d_gradfix.conv2d
    return op(x, w, stride=stride, padding=padding, groups=groups)


# ----------------------------------------------------------------------------
# Phase-major ("space-to-depth") form of the stride-2 layers.
#
#   x_pm[n, (py,px,c), Y, X] = x[n, c, 2Y+py, 2X+px]
#
# A 3x3 stride-2 correlation is a 2x2 stride-1 correlation over the 4C phase channels, a 3x3 stride-2 transposed
# convolution is a 2x2 stride-1 correlation producing the 4O output phases (weights below; 7 of the 16 (phase, tap)
# blocks are structurally zero and are skipped by the tcgen05 kernel, so the MAC count equals the reference's).  That
# way every dense contraction of the networks -- forward, data gradient and weight gradient -- runs through the ONE
# stride-1 tensor-core kernel (conv2d_gradfix.conv2d_s1), with the same results as conv2d_resample.py:119-142.



class _Live(float):
    """Fraction of structurally non-zero (phase, tap) weight blocks of a phase-major weight, carrying the structure itself:
    `pm = (pm_dim, pm_dead)` as defined by gg_conv2d_wgrad_pm_f32 (include/gagan_b200.h) so that the weight-gradient kernel
    can skip the blocks that are zero by construction."""
    pm = None


def _pm_live(kind, kh, kw):
    dead = 0
    for py in range(2):
        for px in range(2):
            for a in range(2):
                for b in range(2):
                    ky, kx = (2 * a + py, 2 * b + px) if kind == 'down' else (py + 2 * (1 - a), px + 2 * (1 - b))
                    if ky >= kh or kx >= kw:
                        dead |= 1 << ((py * 2 + px) * 4 + a * 2 + b)
    live = _Live((16 - bin(dead).count('1')) / 16.0)       # 9/16 for a 3x3 kernel
    live.pm = (2 if kind == 'down' else 1, dead)       # down: dim 1 of W2 [O,4I,2,2] is phase-grouped; up: dim 0 of W2 [4O,I,2,2]
    return live


def phase_major_weight_down(w):
    """[O,I,kh,kw] (kh,kw <= 4) -> [O,4I,2,2]:  W2[o,(py,px,i),a,b] = w[o,i,2a+py,2b+px]  (0 beyond the kernel)."""
    O, I, kh, kw = w.shape
    wp = torch.nn.functional.pad(w, (0, 4 - kw, 0, 4 - kh)).reshape(O, I, 2, 2, 2, 2)      # [O,I,a,py,b,px]
    return wp.permute(0, 3, 5, 1, 2, 4).reshape(O, 4 * I, 2, 2)


def phase_major_weight_up(w):
    """[O,I,kh,kw] -> [4O,I,2,2]:  W2[(py,px,o),i,a,b] = w[o,i,py+2(1-a),px+2(1-b)]  (conv_transpose2d, stride 2, pad 0)."""
    O, I, kh, kw = w.shape
    wp = torch.nn.functional.pad(w, (0, 4 - kw, 0, 4 - kh)).reshape(O, I, 2, 2, 2, 2).flip([2, 4])
    return wp.permute(3, 5, 0, 1, 2, 4).reshape(4 * O, I, 2, 2)


space_to_depth = upfirdn2d.space_to_depth
depth_to_space = upfirdn2d.depth_to_space


def _round_up(v, m):
    return (v + m - 1) // m * m


def down2_phase_major(x, w, f, fir_pad, flip_weight, flip_filter, conv_s1, fir_to_pm, epilogue=None):
    """conv2d_resample.py:119-122 (FIR, then stride-2 conv) with the conv in phase-major form.
    `conv_s1(x, w, padding, out_hw, live)` and `fir_to_pm(x, f, padding, flip_filter, gain, ys, xs)` are injected (the CPU
    algebra test passes torch stand-ins); on the device the FIR writes the phase-major tensor directly."""
    kh, kw = int(w.shape[2]), int(w.shape[3])
    fw, fh = _get_filter_size(f)
    px0, px1, py0, py1 = fir_pad
    fir_h, fir_w = x.shape[2] + py0 + py1 - fh + 1, x.shape[3] + px0 + px1 - fw + 1
    oh, ow = (fir_h - kh) // 2 + 1, (fir_w - kw) // 2 + 1
    if not flip_weight:
        w = w.flip([2, 3])
    xs = fir_to_pm(x, f, fir_pad, flip_filter, 1, oh + 1, _round_up(ow + 1, 4))   # width % 4: TMA row pitch must be 16-byte aligned
    if epilogue is not None:           # the convolution is this layer's last operator: bias / activation ride in its store loop
        return conv_s1(xs, phase_major_weight_down(w), (0, 0), (oh, ow), _pm_live('down', kh, kw), epilogue=epilogue)
    return conv_s1(xs, phase_major_weight_down(w), (0, 0), (oh, ow), _pm_live('down', kh, kw))


def up2_phase_major(x, w, f, fir_pad, flip_weight, flip_filter, conv_s1, fir_from_pm, in_scale=None, out_scale=None):
    """conv2d_resample.py:125-139 (stride-2 transposed conv with pad 0, then FIR with gain 4) in phase-major form;
    `fir_from_pm(z, f, padding, flip_filter, gain, valid_hw)` reads the phase-major conv output directly."""
    N, I, H, W = x.shape
    live = _pm_live('up', int(w.shape[2]), int(w.shape[3]))
    if flip_weight:                      # the reference hands `not flip_weight` to the transposed conv (:138)
        w = w.flip([2, 3])
    xs_w = _round_up(W + 1, 4)           # the gradient of this tensor is a TMA source in backward: keep the width aligned
    if in_scale is not None or out_scale is not None:      # per-sample scales: the 4 output phases of channel o share out_scale[:, o]
        z = conv_s1(x, phase_major_weight_up(w), (1, 1), (H + 1, xs_w), live, in_scale=in_scale,
                    out_scale=(out_scale.repeat(1, 4) if out_scale is not None else None))
    else:
        z = conv_s1(x, phase_major_weight_up(w), (1, 1), (H + 1, xs_w), live)   # [N,4O,H+1,xs_w]; valid logical extent 2H+1 x 2W+1
    return fir_from_pm(z, f, fir_pad, flip_filter, 4, (2 * H + 1, 2 * W + 1))


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


def _conv_s1(x, w, padding, out_hw, live, in_scale=None, out_scale=None, epilogue=None):
    return conv2d_gradfix.conv2d_s1(x, w, padding=padding, out_hw=out_hw, live=float(live), in_scale=in_scale, out_scale=out_scale,
                                    pm=getattr(live, 'pm', None), epilogue=epilogue)


def _bias_act_after(y, epilogue):
    """The unfused form of `epilogue` (branches whose last operator is not the convolution)."""
    from . import bias_act
    if epilogue is None:
        return y
    return bias_act.bias_act(y, epilogue.get('bias'), act=epilogue.get('act', 'linear'), alpha=epilogue.get('alpha'), gain=epilogue.get('gain'),
                             clamp=epilogue.get('clamp'), noise=epilogue.get('noise'))


def _fir_to_pm(x, f, padding, flip_filter, gain, ys, xs):
    return upfirdn2d.fir_to_pm(x, f, padding, flip_filter, gain, ys, xs)


def _fir_from_pm(z, f, padding, flip_filter, gain, valid_hw):
    return upfirdn2d.fir_from_pm(z, f, padding, flip_filter, gain, valid_hw)


@scoped
@fp16_storage('x')
def conv2d_resample(x, w, f=None, up=1, down=1, padding=0, groups=1, flip_weight=True, flip_filter=False, in_scale=None,
                    out_scale=None, epilogue=None):
    r"""2D convolution with optional up/downsampling; padding is applied once, up front.

    x `[N, I, H, W]`, w `[O, I//groups, kh, kw]`, f from `upfirdn2d.setup_filter()` or None.
    `flip_weight=True` = correlation (what `F.conv2d` does), False = convolution.

    Extensions over the reference signature (all optional): `in_scale [N,I]` / `out_scale [N,O]` = the modulation / demodulation
    factors of modulated_conv2d, and `epilogue = dict(bias, noise, act, alpha, gain, clamp)` = the `bias_act(y + noise, bias, ...)`
    call that follows the layer: the result is bias_act(conv2d_resample(...)), computed inside the convolution kernel where the
    convolution is the last operator of the branch and by the bias_act kernel otherwise.
    """
    if epilogue is not None:
        fusable = (groups == 1) and epilogue.get('act', 'linear') in conv2d_gradfix.FUSABLE_ACTS
        _, _, kh_, kw_ = _get_weight_shape(w)
        pl_ = plan(w.shape, f, up, down, padding)
        if fusable and pl_['branch'] == 'plain' and pl_['conv_pad'][0] <= kh_ - 1 and pl_['conv_pad'][1] <= kw_ - 1 \
                and isinstance(x, torch.Tensor) and x.is_cuda and x.dtype == torch.float32:
            return conv2d_gradfix.conv2d_s1(x, w, padding=pl_['conv_pad'], flip=(not flip_weight), in_scale=in_scale, out_scale=out_scale,
                                            epilogue=epilogue)
        if fusable and pl_['branch'] == 'down' and down == 2 and kh_ <= 4 and kw_ <= 4 and in_scale is None and out_scale is None:
            return down2_phase_major(x, w, f, pl_['fir_pad'], flip_weight, flip_filter, _conv_s1, _fir_to_pm, epilogue=epilogue)
        return _bias_act_after(conv2d_resample(x, w, f=f, up=up, down=down, padding=padding, groups=groups, flip_weight=flip_weight,
                                               flip_filter=flip_filter, in_scale=in_scale, out_scale=out_scale), epilogue)

    assert isinstance(x, torch.Tensor) and (x.ndim == 4)
    assert isinstance(w, torch.Tensor) and (w.ndim == 4) and (w.dtype == x.dtype)
    assert f is None or (isinstance(f, torch.Tensor) and f.ndim in [1, 2] and f.dtype == torch.float32)
    assert isinstance(up, int) and (up >= 1)
    assert isinstance(down, int) and (down >= 1)
    assert isinstance(groups, int) and (groups >= 1)
    out_channels, in_channels_per_group, kh, kw = _get_weight_shape(w)
    pl = plan(w.shape, f, up, down, padding)
    branch = pl['branch']

    # `in_scale [N,I]` / `out_scale [N,O]` (an extension used by modulated_conv2d): y = out_scale * op(in_scale * x).  The
    # stride-1 and the phase-major up path fold them into the tensor-core kernel; every other branch multiplies explicitly.
    if in_scale is not None or out_scale is not None:
        assert groups == 1
        if branch == 'plain' and pl['conv_pad'][0] <= kh - 1 and pl['conv_pad'][1] <= kw - 1:
            return conv2d_gradfix.conv2d_s1(x, w, padding=pl['conv_pad'], flip=(not flip_weight), in_scale=in_scale, out_scale=out_scale)
        if branch == 'up' and up == 2 and down == 1 and kh <= 4 and kw <= 4 and pl['conv_pad'] == [0, 0]:
            return up2_phase_major(x, w, f, pl['fir_pad'], flip_weight, flip_filter, _conv_s1, _fir_from_pm, in_scale, out_scale)
        if in_scale is not None:
            x = x * in_scale[:, :, None, None]
        y = conv2d_resample(x, w, f=f, up=up, down=down, padding=padding, groups=groups, flip_weight=flip_weight, flip_filter=flip_filter)
        return y * out_scale[:, :, None, None] if out_scale is not None else y

    if branch == 'down_1x1':      # FIR-decimate first, then the 1x1 conv on the small image
        x = upfirdn2d.upfirdn2d(x=x, f=f, down=down, padding=pl['fir_pad'], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)

    if branch == 'up_1x1':        # 1x1 conv on the small image, then zero-stuff + FIR
        x = _conv2d_wrapper(x=x, w=w, groups=groups, flip_weight=flip_weight)
        return upfirdn2d.upfirdn2d(x=x, f=f, up=up, padding=pl['fir_pad'], gain=up ** 2, flip_filter=flip_filter)

    if branch == 'down':          # FIR at full resolution, then a strided conv
        if down == 2 and groups == 1 and kh <= 4 and kw <= 4:
            return down2_phase_major(x, w, f, pl['fir_pad'], flip_weight, flip_filter, _conv_s1, _fir_to_pm)
        x = upfirdn2d.upfirdn2d(x=x, f=f, padding=pl['fir_pad'], flip_filter=flip_filter)
        return _conv2d_wrapper(x=x, w=w, stride=down, groups=groups, flip_weight=flip_weight)

    if branch == 'up':            # transposed strided conv, then FIR (then optional decimation)
        if up == 2 and down == 1 and groups == 1 and kh <= 4 and kw <= 4 and pl['conv_pad'] == [0, 0]:
            return up2_phase_major(x, w, f, pl['fir_pad'], flip_weight, flip_filter, _conv_s1, _fir_from_pm)
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
