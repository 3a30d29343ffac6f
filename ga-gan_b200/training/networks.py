"""`modulated_conv2d` and the StyleGAN2 generator / discriminator module tree that calls it.

`modulated_conv2d` keeps the reference signature (DissimilarDomains/training/networks.py:591-604) and is
the hot op.  The modules below are the callers that fix its shapes; they mirror the reference classes
name-for-name and parameter-for-parameter (state_dict keys are identical, so reference checkpoints /
golden weights load with `load_state_dict`), restricted to what the fp32 configs use: c_dim >= 0,
G architecture 'skip', D architecture 'resnet', no fp16 blocks.  StyleSpace offsets
('additive' / 'multiplicative' domain modulation, networks.py:140-160,515-523) are supported because the
GA population-evaluation workload (SURVEY.md section 8(d) cfg 4) is defined over them; the low-rank weight
offsets of Affine+/AffineLight+ are a next-round row.

    FullyConnectedLayer :673-704    Conv2dLayer :709-760       MappingNetwork :765-842
    SynthesisLayer :847-922         ToRGBLayer :927-963        SynthesisBlock :968-1074
    SynthesisNetwork :1079-1132     Generator :1137-1171       DiscriminatorBlock :1176-1272
    MinibatchStdLayer :1277-1301    DiscriminatorEpilogue :1306-1367   Discriminator :1372-1437
"""
import numpy as np
import torch

from torch_utils import misc
from torch_utils.ops import conv2d_resample
from torch_utils.ops import upfirdn2d
from torch_utils.ops import bias_act
from torch_utils.ops import fma

# ----------------------------------------------------------------------------


@misc.profiled_function
def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


# ----------------------------------------------------------------------------

@misc.profiled_function
def modulated_conv2d(
        x,                      # Input tensor of shape [batch_size, in_channels, in_height, in_width].
        weight,                 # Weight tensor of shape [out_channels, in_channels, kernel_height, kernel_width].
        styles,                 # Modulation coefficients of shape [batch_size, in_channels].
        noise=None,             # Optional noise tensor to add to the output activations.
        up=1,                   # Integer upsampling factor.
        down=1,                 # Integer downsampling factor.
        padding=0,              # Padding with respect to the upsampled image.
        resample_filter=None,   # Low-pass filter from upfirdn2d.setup_filter().
        demodulate=True,        # Apply weight demodulation?
        flip_weight=True,       # False = convolution, True = correlation (matches F.conv2d).
        fused_modconv=True,     # Accepted for API compatibility; see below.
):
    """y[n,o] = dcoef[n,o] * sum_i conv(styles[n,i] * x[n,i], weight[o,i]) (+ noise).

    The reference has two formulations (networks.py:641-653 non-fused, :655-668 fused/grouped with
    materialised per-sample weights [N*O, I, kh, kw]); they agree to 4e-7 in fp32.
    This build always evaluates the shared-weight form -- per-sample weights are never materialised,
    in training or in eval mode -- and computes the demodulation coefficients as the small GEMM
        dcoef = rsqrt((styles^2) @ (sum_{kh,kw} weight^2)^T + 1e-8)           (app. B.3)
    instead of reducing an [N,O,I,kh,kw] tensor (302 MB per 512-channel layer at N=32).
    """
    batch_size = x.shape[0]
    out_channels, in_channels, kh, kw = weight.shape
    misc.assert_shape(weight, [out_channels, in_channels, kh, kw])
    misc.assert_shape(x, [batch_size, in_channels, None, None])
    misc.assert_shape(styles, [batch_size, in_channels])
    if x.dtype != torch.float32:
        raise RuntimeError('modulated_conv2d: this build serves the fp32 path only')
    _ = fused_modconv

    dcoefs = None
    if demodulate:
        wsq = weight.square().sum(dim=[2, 3])                                  # [O, I]
        dcoefs = (styles.square() @ wsq.t() + 1e-8).rsqrt()                    # [N, O]

    # styles and demodulation coefficients ride inside the convolution kernel (operand conversion / epilogue)
    x = conv2d_resample.conv2d_resample(x=x, w=weight.to(x.dtype), f=resample_filter, up=up, down=down, padding=padding,
                                        flip_weight=flip_weight, in_scale=styles.to(x.dtype), out_scale=dcoefs)
    if noise is not None:
        x = x + noise.to(x.dtype)
    return x


# ----------------------------------------------------------------------------

class FullyConnectedLayer(torch.nn.Module):
    def __init__(self, in_features, out_features, bias=True, activation='linear', lr_multiplier=1, bias_init=0):
        super().__init__()
        self.activation = activation
        self.weight = torch.nn.Parameter(torch.randn([out_features, in_features]) / lr_multiplier)
        self.bias = torch.nn.Parameter(torch.full([out_features], np.float32(bias_init))) if bias else None
        self.weight_gain = lr_multiplier / np.sqrt(in_features)
        self.bias_gain = lr_multiplier

    def forward(self, x):
        w = self.weight.to(x.dtype) * self.weight_gain
        b = self.bias
        if b is not None:
            b = b.to(x.dtype)
            if self.bias_gain != 1:
                b = b * self.bias_gain
        if self.activation == 'linear' and b is not None:
            x = torch.addmm(b.unsqueeze(0), x, w.t())
        else:
            x = x.matmul(w.t())
            x = bias_act.bias_act(x, b, act=self.activation)
        return x


class Conv2dLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, bias=True, activation='linear', up=1, down=1,
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False, trainable=True):
        super().__init__()
        assert not channels_last, 'channels_last is an fp16-path option'
        self.activation = activation
        self.up = up
        self.down = down
        self.conv_clamp = conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        weight = torch.randn([out_channels, in_channels, kernel_size, kernel_size])
        bias = torch.zeros([out_channels]) if bias else None
        if trainable:
            self.weight = torch.nn.Parameter(weight)
            self.bias = torch.nn.Parameter(bias) if bias is not None else None
        else:
            self.register_buffer('weight', weight)
            if bias is not None:
                self.register_buffer('bias', bias)
            else:
                self.bias = None

    def forward(self, x, gain=1):
        w = self.weight * self.weight_gain
        b = self.bias.to(x.dtype) if self.bias is not None else None
        flip_weight = (self.up == 1)
        if b is None and self.activation == 'linear' and self.conv_clamp is None:
            # bias-free linear layer (the resnet skip path): act(x)*gain is a rescaling of the (tiny) weight tensor
            w = w * (self.act_gain * gain)
            return conv2d_resample.conv2d_resample(x=x, w=w.to(x.dtype), f=self.resample_filter, up=self.up, down=self.down,
                                                   padding=self.padding, flip_weight=flip_weight)
        x = conv2d_resample.conv2d_resample(x=x, w=w.to(x.dtype), f=self.resample_filter, up=self.up, down=self.down,
                                            padding=self.padding, flip_weight=flip_weight)
        act_gain = self.act_gain * gain
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return bias_act.bias_act(x, b, act=self.activation, gain=act_gain, clamp=act_clamp)


class MappingNetwork(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, num_ws, num_layers=8, embed_features=None, layer_features=None,
                 activation='lrelu', lr_multiplier=0.01, w_avg_beta=0.995):
        super().__init__()
        self.z_dim = z_dim
        self.c_dim = c_dim
        self.w_dim = w_dim
        self.num_ws = num_ws
        self.num_layers = num_layers
        self.w_avg_beta = w_avg_beta
        if embed_features is None:
            embed_features = w_dim
        if c_dim == 0:
            embed_features = 0
        if layer_features is None:
            layer_features = w_dim
        features_list = [z_dim + embed_features] + [layer_features] * (num_layers - 1) + [w_dim]
        if c_dim > 0:
            self.embed = FullyConnectedLayer(c_dim, embed_features)
        for idx in range(num_layers):
            layer = FullyConnectedLayer(features_list[idx], features_list[idx + 1], activation=activation,
                                        lr_multiplier=lr_multiplier)
            setattr(self, f'fc{idx}', layer)
        if num_ws is not None and w_avg_beta is not None:
            self.register_buffer('w_avg', torch.zeros([w_dim]))

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, skip_w_avg_update=False):
        x = None
        if self.z_dim > 0:
            misc.assert_shape(z, [None, self.z_dim])
            x = normalize_2nd_moment(z.to(torch.float32))
        if self.c_dim > 0:
            misc.assert_shape(c, [None, self.c_dim])
            y = normalize_2nd_moment(self.embed(c.to(torch.float32)))
            x = torch.cat([x, y], dim=1) if x is not None else y
        for idx in range(self.num_layers):
            x = getattr(self, f'fc{idx}')(x)
        if self.w_avg_beta is not None and self.training and not skip_w_avg_update:
            self.w_avg.copy_(x.detach().mean(dim=0).lerp(self.w_avg, self.w_avg_beta))
        if self.num_ws is not None:
            x = x.unsqueeze(1).repeat([1, self.num_ws, 1])
        if truncation_psi != 1:
            assert self.w_avg_beta is not None
            if self.num_ws is None or truncation_cutoff is None:
                x = self.w_avg.lerp(x, truncation_psi)
            else:
                x[:, :truncation_cutoff] = self.w_avg.lerp(x[:, :truncation_cutoff], truncation_psi)
        return x


# ----------------------------------------------------------------------------
# StyleSpace domain modulation (the subset the GA workload needs; networks.py:140-160,474-532).


def _register_modulation(layer, in_channels, use_domain_modulation, parametrization):
    layer.use_domain_modulation = use_domain_modulation
    layer.domain_modulation_parametrization = parametrization if use_domain_modulation else None
    if use_domain_modulation:
        if parametrization not in ('additive', 'multiplicative'):
            raise NotImplementedError(f'domain modulation "{parametrization}" is not part of this round '
                                      '(S-space additive / multiplicative only)')
        layer.offset = torch.nn.Parameter(torch.zeros([1, in_channels]))
        layer.register_buffer('offset_mask', torch.ones([1, in_channels]))
        layer.register_buffer('ones', torch.ones([1, in_channels]))


def w_to_s(layer, w, weight_gain):
    styles = layer.affine(w) * weight_gain
    if layer.use_domain_modulation:
        if layer.domain_modulation_parametrization == 'multiplicative':
            styles = (layer.ones + layer.offset * layer.offset_mask) * styles
        else:
            styles = styles + layer.offset * layer.offset_mask
    return styles


class SynthesisLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=3, up=1, use_noise=True,
                 activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, channels_last=False,
                 use_domain_modulation=False, domain_modulation_parametrization='multiplicative', **_):
        super().__init__()
        assert not channels_last
        self.is_rgb = False
        self.resolution = resolution
        self.up = up
        self.use_noise = use_noise
        self.activation = activation
        self.conv_clamp = conv_clamp
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.padding = kernel_size // 2
        self.act_gain = bias_act.activation_funcs[activation].def_gain
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        if use_noise:
            self.register_buffer('noise_const', torch.randn([resolution, resolution]))
            self.noise_strength = torch.nn.Parameter(torch.zeros([]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        _register_modulation(self, in_channels, use_domain_modulation, domain_modulation_parametrization)
        self.layer_idx = None

    def forward(self, x, w, noise_mode='random', fused_modconv=True, gain=1, **_):
        assert noise_mode in ['random', 'const', 'none']
        in_resolution = self.resolution // self.up
        misc.assert_shape(x, [None, self.weight.shape[1], in_resolution, in_resolution])
        styles = w_to_s(self, w, 1.0)
        noise = None
        if self.use_noise and noise_mode == 'random':
            noise = torch.randn([x.shape[0], 1, self.resolution, self.resolution], device=x.device) * self.noise_strength
        if self.use_noise and noise_mode == 'const':
            noise = self.noise_const * self.noise_strength
        flip_weight = (self.up == 1)
        # the noise is added by the bias_act kernel (one pass instead of add + bias_act); modulated_conv2d(noise=...) itself
        # still accepts it for API compatibility
        x = modulated_conv2d(x=x, weight=self.weight, styles=styles, noise=None, up=self.up, padding=self.padding,
                             resample_filter=self.resample_filter, flip_weight=flip_weight, fused_modconv=fused_modconv)
        act_gain = self.act_gain * gain
        act_clamp = self.conv_clamp * gain if self.conv_clamp is not None else None
        return bias_act.bias_act(x, self.bias.to(x.dtype), act=self.activation, gain=act_gain, clamp=act_clamp, noise=noise)


class ToRGBLayer(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, kernel_size=1, conv_clamp=None, channels_last=False,
                 use_domain_modulation=False, domain_modulation_parametrization='multiplicative', **_):
        super().__init__()
        assert not channels_last
        self.is_rgb = True
        self.resolution = resolution
        self.conv_clamp = conv_clamp
        self.affine = FullyConnectedLayer(w_dim, in_channels, bias_init=1)
        self.weight = torch.nn.Parameter(torch.randn([out_channels, in_channels, kernel_size, kernel_size]))
        self.bias = torch.nn.Parameter(torch.zeros([out_channels]))
        self.weight_gain = 1 / np.sqrt(in_channels * (kernel_size ** 2))
        _register_modulation(self, in_channels, use_domain_modulation, domain_modulation_parametrization)
        self.layer_idx = None

    def forward(self, x, w, fused_modconv=True, **_):
        styles = w_to_s(self, w, self.weight_gain)
        x = modulated_conv2d(x=x, weight=self.weight, styles=styles, demodulate=False, fused_modconv=fused_modconv)
        return bias_act.bias_act(x, self.bias.to(x.dtype), clamp=self.conv_clamp)


class SynthesisBlock(torch.nn.Module):
    def __init__(self, in_channels, out_channels, w_dim, resolution, img_channels, is_last, architecture='skip',
                 resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False, fp16_channels_last=False, **layer_kwargs):
        assert architecture == 'skip', "only the 'skip' generator architecture is served (train.py never selects another)"
        assert not use_fp16, 'fp16 blocks are out of scope for this round'
        super().__init__()
        self.in_channels = in_channels
        self.w_dim = w_dim
        self.resolution = resolution
        self.img_channels = img_channels
        self.is_last = is_last
        self.architecture = architecture
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.num_conv = 0
        self.num_torgb = 0
        if in_channels == 0:
            self.const = torch.nn.Parameter(torch.randn([out_channels, resolution, resolution]))
        if in_channels != 0:
            self.conv0 = SynthesisLayer(in_channels, out_channels, w_dim=w_dim, resolution=resolution, up=2,
                                        resample_filter=resample_filter, conv_clamp=conv_clamp, **layer_kwargs)
            self.num_conv += 1
        self.conv1 = SynthesisLayer(out_channels, out_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                    **layer_kwargs)
        self.num_conv += 1
        self.torgb = ToRGBLayer(out_channels, img_channels, w_dim=w_dim, resolution=resolution, conv_clamp=conv_clamp,
                                **layer_kwargs)
        self.num_torgb += 1

    def forward(self, x, img, ws, force_fp32=False, fused_modconv=None, **layer_kwargs):
        misc.assert_shape(ws, [None, self.num_conv + self.num_torgb, self.w_dim])
        w_iter = iter(ws.unbind(dim=1))
        if fused_modconv is None:
            fused_modconv = not self.training
        if self.in_channels == 0:
            x = self.const.unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
        else:
            misc.assert_shape(x, [None, self.in_channels, self.resolution // 2, self.resolution // 2])
            x = self.conv0(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
            x = self.conv1(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
        if img is not None:   # upstream semantics; the fork lost this guard
            misc.assert_shape(img, [None, self.img_channels, self.resolution // 2, self.resolution // 2])
            img = upfirdn2d.upsample2d(img, self.resample_filter)
        y = self.torgb(x, next(w_iter), fused_modconv=fused_modconv, **layer_kwargs)
        img = img.add_(y) if img is not None else y
        return x, img


class SynthesisNetwork(torch.nn.Module):
    def __init__(self, w_dim, img_resolution, img_channels, channel_base=32768, channel_max=512, num_fp16_res=0,
                 **block_kwargs):
        assert img_resolution >= 4 and img_resolution & (img_resolution - 1) == 0
        assert num_fp16_res == 0, 'this build serves the --fp32 configs (num_fp16_res=0)'
        super().__init__()
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.img_channels = img_channels
        self.block_resolutions = [2 ** i for i in range(2, self.img_resolution_log2 + 1)]
        channels_dict = {res: min(channel_base // res, channel_max) for res in self.block_resolutions}
        self.num_ws = 0
        for res in self.block_resolutions:
            in_channels = channels_dict[res // 2] if res > 4 else 0
            out_channels = channels_dict[res]
            is_last = (res == self.img_resolution)
            block = SynthesisBlock(in_channels, out_channels, w_dim=w_dim, resolution=res, img_channels=img_channels,
                                   is_last=is_last, **block_kwargs)
            self.num_ws += block.num_conv
            if is_last:
                self.num_ws += block.num_torgb
            setattr(self, f'b{res}', block)

    def forward(self, ws, **block_kwargs):
        misc.assert_shape(ws, [None, self.num_ws, self.w_dim])
        ws = ws.to(torch.float32)
        block_ws = []
        w_idx = 0
        for res in self.block_resolutions:
            block = getattr(self, f'b{res}')
            block_ws.append(ws.narrow(1, w_idx, block.num_conv + block.num_torgb))
            w_idx += block.num_conv
        x = img = None
        for res, cur_ws in zip(self.block_resolutions, block_ws):
            x, img = getattr(self, f'b{res}')(x, img, cur_ws, **block_kwargs)
        return img


class Generator(torch.nn.Module):
    def __init__(self, z_dim, c_dim, w_dim, img_resolution, img_channels, mapping_kwargs={}, synthesis_kwargs={}):
        super().__init__()
        self.z_dim = z_dim
        self.c_dim = c_dim
        self.w_dim = w_dim
        self.img_resolution = img_resolution
        self.img_channels = img_channels
        self.synthesis = SynthesisNetwork(w_dim=w_dim, img_resolution=img_resolution, img_channels=img_channels,
                                          **synthesis_kwargs)
        self.num_ws = self.synthesis.num_ws
        self.mapping = MappingNetwork(z_dim=z_dim, c_dim=c_dim, w_dim=w_dim, num_ws=self.num_ws, **mapping_kwargs)
        idx = 0
        for _, module in self.named_modules():   # layer numbering for StyleSpace edits (networks.py:1161-1166)
            if isinstance(module, (SynthesisLayer, ToRGBLayer)):
                module.layer_idx = idx
                idx += 1

    def forward(self, z, c, truncation_psi=1, truncation_cutoff=None, **synthesis_kwargs):
        ws = self.mapping(z, c, truncation_psi=truncation_psi, truncation_cutoff=truncation_cutoff)
        return self.synthesis(ws, **synthesis_kwargs)


# ----------------------------------------------------------------------------

class DiscriminatorBlock(torch.nn.Module):
    def __init__(self, in_channels, tmp_channels, out_channels, resolution, img_channels, first_layer_idx,
                 architecture='resnet', activation='lrelu', resample_filter=[1, 3, 3, 1], conv_clamp=None, use_fp16=False,
                 fp16_channels_last=False, freeze_layers=0, **_):
        assert in_channels in [0, tmp_channels]
        assert architecture == 'resnet', "only the 'resnet' discriminator architecture is served"
        assert not use_fp16
        super().__init__()
        self.in_channels = in_channels
        self.resolution = resolution
        self.img_channels = img_channels
        self.first_layer_idx = first_layer_idx
        self.architecture = architecture
        self.register_buffer('resample_filter', upfirdn2d.setup_filter(resample_filter))
        self.num_layers = 0

        def trainable_gen():
            while True:
                layer_idx = self.first_layer_idx + self.num_layers
                trainable = (layer_idx >= freeze_layers)
                self.num_layers += 1
                yield trainable

        trainable_iter = trainable_gen()
        if in_channels == 0:
            self.fromrgb = Conv2dLayer(img_channels, tmp_channels, kernel_size=1, activation=activation,
                                       trainable=next(trainable_iter), conv_clamp=conv_clamp)
        self.conv0 = Conv2dLayer(tmp_channels, tmp_channels, kernel_size=3, activation=activation,
                                 trainable=next(trainable_iter), conv_clamp=conv_clamp)
        self.conv1 = Conv2dLayer(tmp_channels, out_channels, kernel_size=3, activation=activation, down=2,
                                 trainable=next(trainable_iter), resample_filter=resample_filter, conv_clamp=conv_clamp)
        self.skip = Conv2dLayer(tmp_channels, out_channels, kernel_size=1, bias=False, down=2,
                                trainable=next(trainable_iter), resample_filter=resample_filter)

    def forward(self, x, img, force_fp32=False):
        if x is not None:
            misc.assert_shape(x, [None, self.in_channels, self.resolution, self.resolution])
        if self.in_channels == 0:
            misc.assert_shape(img, [None, self.img_channels, self.resolution, self.resolution])
            y = self.fromrgb(img.to(torch.float32))
            x = x + y if x is not None else y
            img = None
        y = self.skip(x, gain=np.sqrt(0.5))
        x = self.conv0(x)
        x = self.conv1(x, gain=np.sqrt(0.5))
        x = y.add_(x)
        return x, img


class MinibatchStdLayer(torch.nn.Module):
    def __init__(self, group_size, num_channels=1):
        super().__init__()
        self.group_size = group_size
        self.num_channels = num_channels

    def forward(self, x):
        N, C, H, W = x.shape
        G = min(int(self.group_size), int(N)) if self.group_size is not None else int(N)
        F = self.num_channels
        c = C // F
        y = x.reshape(G, -1, F, c, H, W)    # [GnFcHW] split minibatch N into n groups of size G, channels into F groups
        y = y - y.mean(dim=0)               # subtract the group mean
        y = y.square().mean(dim=0)          # [nFcHW] variance over the group
        y = (y + 1e-8).sqrt()
        y = y.mean(dim=[2, 3, 4])           # [nF]
        y = y.reshape(-1, F, 1, 1)
        y = y.repeat(G, 1, H, W)            # [NFHW]
        return torch.cat([x, y], dim=1)


class DiscriminatorEpilogue(torch.nn.Module):
    def __init__(self, in_channels, cmap_dim, resolution, img_channels, architecture='resnet', mbstd_group_size=4,
                 mbstd_num_channels=1, activation='lrelu', conv_clamp=None, **_):
        assert architecture == 'resnet'
        super().__init__()
        self.in_channels = in_channels
        self.cmap_dim = cmap_dim
        self.resolution = resolution
        self.img_channels = img_channels
        self.architecture = architecture
        self.mbstd = MinibatchStdLayer(group_size=mbstd_group_size, num_channels=mbstd_num_channels) if mbstd_num_channels > 0 else None
        self.conv = Conv2dLayer(in_channels + mbstd_num_channels, in_channels, kernel_size=3, activation=activation,
                                conv_clamp=conv_clamp)
        self.fc = FullyConnectedLayer(in_channels * (resolution ** 2), in_channels, activation=activation)
        self.out = FullyConnectedLayer(in_channels, 1 if cmap_dim == 0 else cmap_dim)

    def forward(self, x, img, cmap, force_fp32=False):
        misc.assert_shape(x, [None, self.in_channels, self.resolution, self.resolution])
        x = x.to(torch.float32)
        if self.mbstd is not None:
            x = self.mbstd(x)
        x = self.conv(x)
        x = self.fc(x.flatten(1))
        x = self.out(x)
        if self.cmap_dim > 0:
            misc.assert_shape(cmap, [None, self.cmap_dim])
            x = (x * cmap).sum(dim=1, keepdim=True) * (1 / np.sqrt(self.cmap_dim))
        return x


class Discriminator(torch.nn.Module):
    def __init__(self, c_dim, img_resolution, img_channels, architecture='resnet', channel_base=32768, channel_max=512,
                 num_fp16_res=0, conv_clamp=None, cmap_dim=None, block_kwargs={}, mapping_kwargs={}, epilogue_kwargs={}):
        assert num_fp16_res == 0, 'this build serves the --fp32 configs (num_fp16_res=0)'
        super().__init__()
        self.c_dim = c_dim
        self.img_resolution = img_resolution
        self.img_resolution_log2 = int(np.log2(img_resolution))
        self.img_channels = img_channels
        self.block_resolutions = [2 ** i for i in range(self.img_resolution_log2, 2, -1)]
        channels_dict = {res: min(channel_base // res, channel_max) for res in self.block_resolutions + [4]}
        if cmap_dim is None:
            cmap_dim = channels_dict[4]
        if c_dim == 0:
            cmap_dim = 0
        common_kwargs = dict(img_channels=img_channels, architecture=architecture, conv_clamp=conv_clamp)
        cur_layer_idx = 0
        for res in self.block_resolutions:
            in_channels = channels_dict[res] if res < img_resolution else 0
            block = DiscriminatorBlock(in_channels, channels_dict[res], channels_dict[res // 2], resolution=res,
                                       first_layer_idx=cur_layer_idx, **block_kwargs, **common_kwargs)
            setattr(self, f'b{res}', block)
            cur_layer_idx += block.num_layers
        if c_dim > 0:
            self.mapping = MappingNetwork(z_dim=0, c_dim=c_dim, w_dim=cmap_dim, num_ws=None, w_avg_beta=None, **mapping_kwargs)
        self.b4 = DiscriminatorEpilogue(channels_dict[4], cmap_dim=cmap_dim, resolution=4, **epilogue_kwargs, **common_kwargs)

    def forward(self, img, c, **block_kwargs):
        x = None
        for res in self.block_resolutions:
            x, img = getattr(self, f'b{res}')(x, img, **block_kwargs)
        cmap = None
        if self.c_dim > 0:
            cmap = self.mapping(None, c)
        return self.b4(x, img, cmap)
