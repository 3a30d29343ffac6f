"""CPU oracle for the CALLERS of the hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Functional (parameter-dict) restatement of the StyleGAN2 generator,
discriminator and the four loss phases as the reference defines them, built on
`oracle.ops_ref`.  Parameter / buffer names are the reference's state_dict names
so that one weight dict drives the reference (when generating goldens), this
oracle and the product modules.

    FullyConnectedLayer      DissimilarDomains/training/networks.py:673-704
    Conv2dLayer              :709-760
    MappingNetwork           :765-842
    SynthesisLayer / ToRGB   :847-963
    SynthesisBlock / Network :968-1132   (with the upstream `img is None` guard,
                                          SURVEY.md section 0.2 -- the fork de-indented it)
    DiscriminatorBlock       :1176-1272
    MinibatchStdLayer        :1277-1301
    DiscriminatorEpilogue    :1306-1367
    Discriminator            :1372-1437
    StyleGAN2Loss phases     DissimilarDomains/training/loss.py:44-152

Scope: fp32, c_dim=0, architectures G='skip' D='resnet' (the only ones the
reference's configs use, train.py:264-273), no domain-modulation offsets.
Pinned against the live reference by tests/golden/make_golden.py.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import ops_ref as R

SQRT2 = np.sqrt(2)
SQRT_HALF = np.sqrt(0.5)


def channels_dict(img_resolution, channel_base=32768, channel_max=512):
    log2 = int(np.log2(img_resolution))
    return {2 ** i: min(channel_base // (2 ** i), channel_max) for i in range(2, log2 + 1)}


def num_ws_for(img_resolution):
    # one w per conv layer + one for the last ToRGB: networks.py:1101-1114
    log2 = int(np.log2(img_resolution))
    return 2 * (log2 - 1)


# ----------------------------------------------------------------------------


def fully_connected(P, name, x, activation='linear', lr_multiplier=1.0):
    """networks.py:691-704"""
    w = P[name + '.weight'] * (lr_multiplier / np.sqrt(P[name + '.weight'].shape[1]))
    b = P.get(name + '.bias')
    if b is not None and lr_multiplier != 1:
        b = b * lr_multiplier
    if activation == 'linear' and b is not None:
        return torch.addmm(b.unsqueeze(0), x, w.t())
    x = x.matmul(w.t())
    return R.bias_act(x, b, act=activation)


def conv2d_layer(P, name, x, f, kernel_size, activation='linear', up=1, down=1, gain=1.0, conv_clamp=None):
    """networks.py:748-760"""
    weight = P[name + '.weight']
    w = weight * (1 / np.sqrt(weight.shape[1] * kernel_size ** 2))
    b = P.get(name + '.bias')
    x = R.conv2d_resample(x, w, f=f, up=up, down=down, padding=kernel_size // 2, flip_weight=(up == 1))
    act_gain = R.ACTIVATIONS[activation][1] * gain
    act_clamp = conv_clamp * gain if conv_clamp is not None else None
    return R.bias_act(x, b, act=activation, gain=act_gain, clamp=act_clamp)


def normalize_2nd_moment(x, dim=1, eps=1e-8):
    return x * (x.square().mean(dim=dim, keepdim=True) + eps).rsqrt()


def mapping(P, z, num_ws, num_layers=8, prefix='mapping', truncation_psi=1, truncation_cutoff=None):
    """networks.py:807-842 (c_dim=0; the w_avg EMA update is a side effect handled by the caller)."""
    x = normalize_2nd_moment(z.to(P[f'{prefix}.fc0.weight'].dtype))   # float32 in the reference (:815); the parameter dtype lets tests run an fp64 truth
    for i in range(num_layers):
        x = fully_connected(P, f'{prefix}.fc{i}', x, activation='lrelu', lr_multiplier=0.01)
    x = x.unsqueeze(1).repeat([1, num_ws, 1])
    if truncation_psi != 1:
        w_avg = P[f'{prefix}.w_avg']
        if truncation_cutoff is None:
            x = w_avg.lerp(x, truncation_psi)
        else:
            x = torch.cat([w_avg.lerp(x[:, :truncation_cutoff], truncation_psi), x[:, truncation_cutoff:]], dim=1)
    return x


def synthesis_layer(P, name, x, w, f, resolution, up=1, noise_mode='random', fused_modconv=True, gain=1.0,
                    conv_clamp=None):
    """networks.py:896-922"""
    styles = fully_connected(P, name + '.affine', w)
    noise = None
    if noise_mode == 'random':
        noise = torch.randn([x.shape[0], 1, resolution, resolution]) * P[name + '.noise_strength']
    if noise_mode == 'const':
        noise = P[name + '.noise_const'] * P[name + '.noise_strength']
    x = R.modulated_conv2d(x, P[name + '.weight'], styles, noise=noise, up=up, padding=1, resample_filter=f,
                           flip_weight=(up == 1), fused_modconv=fused_modconv)
    act_gain = SQRT2 * gain
    act_clamp = conv_clamp * gain if conv_clamp is not None else None
    return R.bias_act(x, P[name + '.bias'], act='lrelu', gain=act_gain, clamp=act_clamp)


def torgb_layer(P, name, x, w, fused_modconv=True, conv_clamp=None):
    """networks.py:957-963"""
    weight = P[name + '.weight']
    styles = fully_connected(P, name + '.affine', w) * (1 / np.sqrt(weight.shape[1]))
    x = R.modulated_conv2d(x, weight, styles, demodulate=False, fused_modconv=fused_modconv)
    return R.bias_act(x, P[name + '.bias'], clamp=conv_clamp)


def synthesis(P, ws, img_resolution, noise_mode='random', fused_modconv=False, prefix='synthesis', conv_clamp=None):
    """networks.py:1028-1074,1117-1132 (architecture 'skip').  `fused_modconv` defaults to the training value."""
    f = R.setup_filter([1, 3, 3, 1])
    log2 = int(np.log2(img_resolution))
    x = img = None
    w_idx = 0
    ws = ws.to(P[f'{prefix}.b4.const'].dtype)                          # float32 in the reference (:1119)
    for res in [2 ** i for i in range(2, log2 + 1)]:
        b = f'{prefix}.b{res}'
        if res == 4:
            x = P[b + '.const'].unsqueeze(0).repeat([ws.shape[0], 1, 1, 1])
            x = synthesis_layer(P, b + '.conv1', x, ws[:, w_idx], f, res, noise_mode=noise_mode,
                                fused_modconv=fused_modconv, conv_clamp=conv_clamp)
            n_conv = 1
        else:
            x = synthesis_layer(P, b + '.conv0', x, ws[:, w_idx], f, res, up=2, noise_mode=noise_mode,
                                fused_modconv=fused_modconv, conv_clamp=conv_clamp)
            x = synthesis_layer(P, b + '.conv1', x, ws[:, w_idx + 1], f, res, noise_mode=noise_mode,
                                fused_modconv=fused_modconv, conv_clamp=conv_clamp)
            n_conv = 2
        if img is not None:
            img = R.upsample2d(img, f)
        y = torgb_layer(P, b + '.torgb', x, ws[:, w_idx + n_conv], fused_modconv=fused_modconv, conv_clamp=conv_clamp)
        img = img + y if img is not None else y
        w_idx += n_conv
    return img


def minibatch_std(x, group_size=4, num_channels=1):
    """networks.py:1284-1301"""
    N, C, H, W = x.shape
    G = min(group_size, N) if group_size is not None else N
    Fc = num_channels
    c = C // Fc
    y = x.reshape(G, -1, Fc, c, H, W)
    y = y - y.mean(dim=0)
    y = y.square().mean(dim=0)
    y = (y + 1e-8).sqrt()
    y = y.mean(dim=[2, 3, 4])
    y = y.reshape(-1, Fc, 1, 1)
    y = y.repeat(G, 1, H, W)
    return torch.cat([x, y], dim=1)


def discriminator(P, img, img_resolution, mbstd_group_size=4, conv_clamp=None):
    """networks.py:1244-1272,1341-1367,1427-1437 (architecture 'resnet', c_dim=0)."""
    f = R.setup_filter([1, 3, 3, 1])
    log2 = int(np.log2(img_resolution))
    x = None
    for res in [2 ** i for i in range(log2, 2, -1)]:
        b = f'b{res}'
        if x is None:
            x = conv2d_layer(P, b + '.fromrgb', img, f, 1, activation='lrelu', conv_clamp=conv_clamp)
        y = conv2d_layer(P, b + '.skip', x, f, 1, down=2, gain=SQRT_HALF)
        x = conv2d_layer(P, b + '.conv0', x, f, 3, activation='lrelu', conv_clamp=conv_clamp)
        x = conv2d_layer(P, b + '.conv1', x, f, 3, activation='lrelu', down=2, gain=SQRT_HALF, conv_clamp=conv_clamp)
        x = y + x
    if x is None:  # img_resolution == 4
        x = conv2d_layer(P, 'b4.fromrgb', img, f, 1, activation='lrelu', conv_clamp=conv_clamp)
    x = minibatch_std(x, group_size=mbstd_group_size)
    x = conv2d_layer(P, 'b4.conv', x, f, 3, activation='lrelu', conv_clamp=conv_clamp)
    x = fully_connected(P, 'b4.fc', x.flatten(1), activation='lrelu')
    x = fully_connected(P, 'b4.out', x)
    return x


# ----------------------------------------------------------------------------
# Loss phases (loss.py:69-152), style mixing handled by the caller through `ws`.


def run_G(PG, z, img_resolution, noise_mode='random', style_mixing=None, num_layers=8):
    """loss.py:44-60.  `style_mixing` = None or (cutoff:int, z2) to mirror `:47-55` deterministically."""
    nws = num_ws_for(img_resolution)
    ws = mapping(PG, z, nws, num_layers)
    if style_mixing is not None:
        cutoff, z2 = style_mixing
        ws2 = mapping(PG, z2, nws, num_layers)
        ws = torch.cat([ws[:, :cutoff], ws2[:, cutoff:]], dim=1)
    return ws


def loss_Gmain(PG, PD, z, res, mbstd=4, noise_mode='random', num_layers=8):
    ws = run_G(PG, z, res, num_layers=num_layers)
    img = synthesis(PG, ws, res, noise_mode=noise_mode)
    logits = discriminator(PD, img, res, mbstd_group_size=mbstd)
    return F.softplus(-logits).mean()


def loss_Gpl(PG, z, res, pl_mean, pl_noise=None, pl_batch_shrink=2, pl_decay=0.01, pl_weight=2, noise_mode='random',
             num_layers=8):
    """loss.py:89-111.  Returns (loss, new_pl_mean)."""
    bs = z.shape[0] // pl_batch_shrink
    ws = run_G(PG, z[:bs], res, num_layers=num_layers)
    if not ws.requires_grad:  # loss.py:50-51 set_w_requires_grad
        ws.requires_grad_(True)
    img = synthesis(PG, ws, res, noise_mode=noise_mode)
    if pl_noise is None:
        pl_noise = torch.randn_like(img)
    pl_noise = pl_noise / np.sqrt(img.shape[2] * img.shape[3])
    pl_grads = torch.autograd.grad(outputs=[(img * pl_noise).sum()], inputs=[ws], create_graph=True)[0]
    pl_lengths = pl_grads.square().sum(2).mean(1).sqrt()
    new_mean = pl_mean.lerp(pl_lengths.mean(), pl_decay)
    pl_penalty = (pl_lengths - new_mean).square()   # loss.py:104-106: the new mean is NOT detached here
    loss = (img[:, 0, 0, 0] * 0 + pl_penalty * pl_weight).mean()
    return loss, new_mean.detach()


def loss_Dmain(PG, PD, z, real, res, mbstd=4, noise_mode='random', num_layers=8):
    """loss.py:114-152 with do_Dmain only: softplus(fake) + softplus(-real)."""
    with torch.no_grad():
        ws = run_G(PG, z, res, num_layers=num_layers)
        fake = synthesis(PG, ws, res, noise_mode=noise_mode)
    l_gen = F.softplus(discriminator(PD, fake, res, mbstd_group_size=mbstd)).mean()
    l_real = F.softplus(-discriminator(PD, real, res, mbstd_group_size=mbstd)).mean()
    return l_gen + l_real


def loss_Dr1(PD, real, res, r1_gamma=10.0, mbstd=4):
    """loss.py:127-152 with do_Dr1 only."""
    real = real.detach().requires_grad_(True)
    logits = discriminator(PD, real, res, mbstd_group_size=mbstd)
    r1_grads = torch.autograd.grad(outputs=[logits.sum()], inputs=[real], create_graph=True)[0]
    r1_penalty = r1_grads.square().sum([1, 2, 3])
    return (logits * 0 + (r1_penalty * (r1_gamma / 2)).unsqueeze(1)).mean()


# ----------------------------------------------------------------------------
# Random-init parameter dicts with the reference's names, shapes and init laws.


def init_G_params(img_resolution, channel_base=32768, channel_max=512, z_dim=512, w_dim=512, num_layers=8,
                  generator=None, randomize=False):
    """Shapes per networks.py:685-686, 880-889, 935-941, 997-998; init laws randn / zeros / ones.

    randomize=True additionally randomises the zero-initialised tensors (biases, noise_strength) so
    that parity tests exercise them.
    """
    g = generator
    rn = lambda *s: torch.randn(*s, generator=g)
    P = {}
    cd = channels_dict(img_resolution, channel_base, channel_max)
    for i in range(num_layers):
        fin = z_dim if i == 0 else w_dim
        P[f'mapping.fc{i}.weight'] = rn(w_dim, fin) / 0.01
        P[f'mapping.fc{i}.bias'] = rn(w_dim) if randomize else torch.zeros(w_dim)
    P['mapping.w_avg'] = torch.zeros(w_dim)

    def synth_layer(name, cin, cout, k, res, noise):
        P[name + '.affine.weight'] = rn(cin, w_dim)
        P[name + '.affine.bias'] = torch.ones(cin) + (0.1 * rn(cin) if randomize else 0)
        P[name + '.weight'] = rn(cout, cin, k, k)
        if noise:
            P[name + '.noise_const'] = rn(res, res)
            P[name + '.noise_strength'] = 0.1 * rn(1).reshape([]) if randomize else torch.zeros([])
        P[name + '.bias'] = 0.1 * rn(cout) if randomize else torch.zeros(cout)

    for res, c in cd.items():
        b = f'synthesis.b{res}'
        if res == 4:
            P[b + '.const'] = rn(c, 4, 4)
        else:
            synth_layer(b + '.conv0', cd[res // 2], c, 3, res, True)
        synth_layer(b + '.conv1', c, c, 3, res, True)
        synth_layer(b + '.torgb', c, 3, 1, res, False)
    return P


def init_D_params(img_resolution, channel_base=32768, channel_max=512, generator=None, randomize=False):
    g = generator
    rn = lambda *s: torch.randn(*s, generator=g)
    P = {}
    cd = channels_dict(img_resolution, channel_base, channel_max)
    log2 = int(np.log2(img_resolution))

    def conv(name, cin, cout, k, bias=True):
        P[name + '.weight'] = rn(cout, cin, k, k)
        if bias:
            P[name + '.bias'] = 0.1 * rn(cout) if randomize else torch.zeros(cout)

    for res in [2 ** i for i in range(log2, 2, -1)]:
        b = f'b{res}'
        if res == img_resolution:
            conv(b + '.fromrgb', 3, cd[res], 1)
        conv(b + '.conv0', cd[res], cd[res], 3)
        conv(b + '.conv1', cd[res], cd[res // 2], 3)
        conv(b + '.skip', cd[res], cd[res // 2], 1, bias=False)
    c4 = cd[4]
    conv('b4.conv', c4 + 1, c4, 3)
    P['b4.fc.weight'] = rn(c4, c4 * 16)
    P['b4.fc.bias'] = 0.1 * rn(c4) if randomize else torch.zeros(c4)
    P['b4.out.weight'] = rn(1, c4)
    P['b4.out.bias'] = 0.1 * rn(1) if randomize else torch.zeros(1)
    return P
