"""The B200 operators behind a rosinality-style StyleGAN2 module (second half).

GA-GAN's `SimilarDomains/` tree carries a second StyleGAN2 implementation (gan_models/StyleGAN2/model.py, the rosinality
code base) whose hot path is the same mathematics under other names:

    reference (SimilarDomains/gan_models/StyleGAN2/...)              this build
    ---------------------------------------------------------------  ------------------------------------------------------------
    op.upfirdn2d(input, kernel, up, down, pad)                        torch_utils.ops.upfirdn2d.upfirdn2d(x, f=kernel, up, down,
      op/upfirdn2d_torch_native.py:10-59 (the only one op/__init__      padding=[p0, p1, p0, p1])  -- true convolution with the
      selects, :1-6)                                                   kernel in both (flip_filter=False), gains baked into it
    op.fused_leaky_relu(input, bias, negative_slope, scale)           bias_act(x, bias, dim, act='lrelu', alpha=negative_slope,
      op/fused_act_torch_native.py:23-37, FusedLeakyReLU :10-20        gain=scale)
    ModulatedConv2d.forward (model.py:230-275): per-sample weights    training.networks.modulated_conv2d (shared-weight form;
      scale * W * style, demodulated, grouped F.conv2d /              styles and demodulation coefficients ride inside the
      F.conv_transpose2d(stride 2) + Blur, or Blur + stride-2 conv      tcgen05 kernel; up / down through the phase-major forms)
    EqualConv2d.forward (model.py:108-117): F.conv2d                  conv2d_gradfix.conv2d
    ConvLayer = [Blur,] EqualConv2d [, FusedLeakyReLU|ScaledLeakyReLU] conv2d_resample(x, w, f, down=2, padding=k//2,
      (model.py:666-713), the discriminator's layer                     epilogue=dict(bias, act='lrelu', ...)) -- one plan, the
                                                                       activation in the convolution's store loop where possible
    StyledConv.forward (model.py:334-341): conv, NoiseInjection,      modulated_conv2d(..., noise=weight * noise,
      FusedLeakyReLU                                                    epilogue=dict(bias, act='lrelu', gain=scale))

The pad arithmetic of the two code bases agrees for the odd kernel sizes the networks use: rosinality's Blur pads
((p+1)//2 + factor - 1, p//2 + 1) with p = 4 - factor - (k - 1) around a transposed convolution and ((p+1)//2, p//2) with
p = 4 - factor + (k - 1) before a strided one (model.py:196-211, 680-685); conv2d_resample.py:94-104,132-137 with padding = k // 2
yields the same FIR pads.  `install_rosinality` checks that equality per layer instead of trusting it, and refuses (raises) layers
it cannot express -- it never falls back to the module's own torch code.

    import gagan_b200
    gagan_b200.install(None)                       # or install('<...>/DissimilarDomains')
    from gan_models.StyleGAN2 import model         # the user's SimilarDomains checkout
    gagan_b200.install_rosinality(model)           # Generator / Discriminator of that module now run on libgagan_b200.so
"""
import math
import torch

from .torch_utils.ops import bias_act as _bias_act
from .torch_utils.ops import upfirdn2d as _upfirdn2d
from .torch_utils.ops import conv2d_gradfix as _conv2d_gradfix
from .torch_utils.ops import conv2d_resample as _conv2d_resample
from .training import networks as _networks


def upfirdn2d(input, kernel, up=1, down=1, pad=(0, 0)):
    """op/upfirdn2d_torch_native.py:10-15: the same pad pair on both axes, convolution with `kernel` as given."""
    return _upfirdn2d.upfirdn2d(input, kernel.float(), up=up, down=down, padding=[int(pad[0]), int(pad[1]), int(pad[0]), int(pad[1])])


def fused_leaky_relu(input, bias, negative_slope=0.2, scale=2 ** 0.5):
    """op/fused_act_torch_native.py:23-37: leaky_relu(input + bias) * scale; the bias runs along dim 1, or along the last dim of a
    rank-3 tensor."""
    dim = input.ndim - 1 if input.ndim == 3 else 1
    return _bias_act.bias_act(input, bias, dim=dim, act='lrelu', alpha=negative_slope, gain=scale)


def _normalized_filter(blur, gain):
    """rosinality bakes the up-sampling gain into the Blur kernel (model.py:77-78); conv2d_resample applies `up ** 2` itself."""
    f = blur.kernel.float()                 # (filters are float32 whatever the module was cast to: upfirdn2d.py:81-125)
    return f if gain == 1 else f / gain


def _check_pads(what, w_shape, f, up, down, padding, blur_pad):
    pl = _conv2d_resample.plan(w_shape, f, up, down, padding)
    p0, p1 = int(blur_pad[0]), int(blur_pad[1])
    if pl['branch'] == 'up_1x1':
        # a 1x1 transposed convolution with stride `up` is zero-stuffing WITHOUT the trailing zeros that upfirdn2d's up-sampling
        # appends (conv2d_resample.py:113-116 runs the 1x1 convolution first and lets upfirdn2d stuff): one pad less on the far side
        p1 -= up - 1
    want = [p0, p1, p0, p1]
    conv_pad = pl.get('conv_pad', [0, 0])
    if pl['fir_pad'] != want or (up > 1 and conv_pad != [0, 0]):
        raise NotImplementedError(f'gagan_b200.install_rosinality: {what} pads {want} around its resampling convolution; '
                                  f'conv2d_resample(padding={padding}) gives {pl["fir_pad"]} (conv padding {conv_pad})')


def _modulated_conv2d_forward(self, input, style, is_s_code=False, offset_power=1., offsets=None, noise=None, epilogue=None):
    """ModulatedConv2d.forward (model.py:230-275) on training.networks.modulated_conv2d.  `offset_power` / `offsets` are accepted
    and unused, as in the reference class (its subclasses in offsets_model.py act on them)."""
    batch, in_channel = int(input.shape[0]), int(input.shape[1])
    if not is_s_code:
        style = self.modulation(style)
    style = style.reshape(batch, in_channel)
    weight = self.weight[0] * self.scale                                   # [O, I, k, k]
    k = int(self.kernel_size)
    kw = dict(demodulate=bool(self.demodulate), noise=noise, epilogue=epilogue)
    if self.upsample:
        f = _normalized_filter(self.blur, 4)
        _check_pads('ModulatedConv2d(upsample=True)', weight.shape, f, 2, 1, k // 2, self.blur.pad)
        # conv_transpose2d(input, weight[I,O]) without a flip == modulated_conv2d(up=2, flip_weight=False)  (conv2d_resample.py:125-139)
        return _networks.modulated_conv2d(x=input, weight=weight, styles=style, up=2, padding=k // 2, resample_filter=f, flip_weight=False, **kw)
    if self.downsample:
        f = _normalized_filter(self.blur, 1)
        _check_pads('ModulatedConv2d(downsample=True)', weight.shape, f, 1, 2, k // 2, self.blur.pad)
        return _networks.modulated_conv2d(x=input, weight=weight, styles=style, down=2, padding=k // 2, resample_filter=f, flip_weight=True, **kw)
    return _networks.modulated_conv2d(x=input, weight=weight, styles=style, padding=int(self.padding), flip_weight=True, **kw)


def _equal_conv2d_forward(self, input):
    """EqualConv2d.forward (model.py:108-117)."""
    return _conv2d_gradfix.conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride, padding=self.padding)


def _fused_leaky_relu_module_forward(self, input):
    return fused_leaky_relu(input, self.bias, self.negative_slope, self.scale)


def _styled_conv_forward(self, input, style, noise=None, is_s_code=False):
    """StyledConv.forward (model.py:334-341): modulated convolution, noise injection (`image + weight * noise`, fresh N(0,1) noise per
    call when none is given: model.py:284-289) and the fused leaky ReLU, as ONE modulated_conv2d call with its epilogue."""
    act = self.activate
    up = 2 if self.conv.upsample else 1
    if noise is None:
        noise = torch.randn([int(input.shape[0]), 1, int(input.shape[2]) * up, int(input.shape[3]) * up], device=input.device, dtype=input.dtype)
    epilogue = dict(bias=act.bias, act='lrelu', alpha=act.negative_slope, gain=act.scale, clamp=None)
    return _modulated_conv2d_forward(self.conv, input, style, is_s_code=is_s_code, noise=self.noise.weight * noise, epilogue=epilogue)


def _conv_layer_forward(m):
    Blur, EqualConv2d, FusedLeakyReLU, ScaledLeakyReLU = m.Blur, m.EqualConv2d, m.FusedLeakyReLU, m.ScaledLeakyReLU

    def forward(self, input):
        """ConvLayer (model.py:666-713) = [Blur,] EqualConv2d [, activation] as one conv2d_resample call with its epilogue."""
        mods = list(self)
        blur = mods.pop(0) if mods and isinstance(mods[0], Blur) else None
        conv = mods.pop(0) if mods and isinstance(mods[0], EqualConv2d) else None
        act = mods.pop(0) if mods else None
        if conv is None or mods or not (act is None or isinstance(act, (FusedLeakyReLU, ScaledLeakyReLU))):
            raise NotImplementedError('gagan_b200.install_rosinality: ConvLayer with an unexpected layer sequence ' + repr(self))
        w = conv.weight * conv.scale
        k = int(w.shape[2])
        if isinstance(act, FusedLeakyReLU):
            epilogue = dict(bias=act.bias, act='lrelu', alpha=act.negative_slope, gain=act.scale, clamp=None)
        elif isinstance(act, ScaledLeakyReLU):
            epilogue = dict(bias=conv.bias, act='lrelu', alpha=act.negative_slope, gain=math.sqrt(2), clamp=None)
        else:
            epilogue = dict(bias=conv.bias, act='linear', alpha=None, gain=1, clamp=None) if conv.bias is not None else None
        if isinstance(act, FusedLeakyReLU) and conv.bias is not None:
            raise NotImplementedError('gagan_b200.install_rosinality: ConvLayer with two biases')
        if blur is not None:
            if conv.stride != 2 or conv.padding != 0:
                raise NotImplementedError('gagan_b200.install_rosinality: a Blur in front of a convolution that is not stride 2 / padding 0')
            f = _normalized_filter(blur, 1)
            _check_pads('ConvLayer(downsample=True)', w.shape, f, 1, 2, k // 2, blur.pad)
            return _conv2d_resample.conv2d_resample(x=input, w=w, f=f, down=2, padding=k // 2, epilogue=epilogue)
        if conv.stride != 1:
            raise NotImplementedError('gagan_b200.install_rosinality: strided EqualConv2d without a Blur')
        return _conv2d_resample.conv2d_resample(x=input, w=w, padding=int(conv.padding), epilogue=epilogue)
    return forward


def install_rosinality(model_module, fused_layers=True):
    """Bind this build's operators into a rosinality-style StyleGAN2 module (SimilarDomains/gan_models/StyleGAN2/model.py, or any
    module with the same classes).  Idempotent; returns the module.

    Always: the module-level `upfirdn2d` / `fused_leaky_relu` names (Upsample, Downsample, Blur, EqualLinear resolve them at call
    time), `FusedLeakyReLU.forward`, `ModulatedConv2d.forward`, `EqualConv2d.forward`.
    fused_layers: additionally `StyledConv.forward` and `ConvLayer.forward` as single calls with their noise / bias / activation
    epilogue (same values, fewer passes over HBM).  With False those two keep the module's own code on the replaced pieces."""
    m = model_module
    state = getattr(m, '_gagan_b200_rosinality', None)
    if state is None:
        state = dict(styled_conv_forward=m.StyledConv.forward, conv_layer_forward=m.ConvLayer.forward,
                     fused_conv_layer_forward=_conv_layer_forward(m))
        m._gagan_b200_rosinality = state
        m.upfirdn2d = upfirdn2d
        m.fused_leaky_relu = fused_leaky_relu
        m.FusedLeakyReLU.forward = _fused_leaky_relu_module_forward
        m.ModulatedConv2d.forward = _modulated_conv2d_forward
        m.EqualConv2d.forward = _equal_conv2d_forward
    m.StyledConv.forward = _styled_conv_forward if fused_layers else state['styled_conv_forward']
    m.ConvLayer.forward = state['fused_conv_layer_forward'] if fused_layers else state['conv_layer_forward']
    state['fused_layers'] = bool(fused_layers)
    return m
