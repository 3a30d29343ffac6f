#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ FROM THE LIVE REFERENCE and pin the oracle.

Runs only in the authoring container (needs /root/reference).  It

  1. imports the reference's own `impl='ref'` operators and networks from
     /root/reference/DissimilarDomains (with the external `img is None` guard of
     SURVEY.md section 0.2 -- the reference is not modified),
  2. executes them on seeded inputs (CPU, fp32, plain autograd for all gradients),
  3. asserts that every function in oracle/ reproduces the reference output
     (this is the oracle pin: tolerance 2e-5 max-relative for fp32 reorderings,
     exact for integer bookkeeping),
  4. writes inputs + reference outputs to tests/golden/*.npz.

The GPU box has no /root/reference; tests there read only the .npz files.

    python tests/golden/make_golden.py
"""
import os
import sys
import itertools
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = '/root/reference/DissimilarDomains'
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.set_num_threads(8)

from torch_utils import misc as ref_misc                     # noqa: E402  (reference)
from torch_utils.ops import upfirdn2d as ref_upfirdn2d       # noqa: E402
from torch_utils.ops import bias_act as ref_bias_act         # noqa: E402
from torch_utils.ops import conv2d_resample as ref_c2r       # noqa: E402
from torch_utils.ops import fma as ref_fma                   # noqa: E402

# The guard the fork de-indented away (networks.py:1058-1063); applied from outside.
_as, _up = ref_misc.assert_shape, ref_upfirdn2d.upsample2d
ref_misc.assert_shape = lambda t, s: None if t is None else _as(t, s)
ref_upfirdn2d.upsample2d = lambda x, f, **kw: None if x is None else _up(x, f, **kw)

from training import networks as ref_networks                # noqa: E402
from training import loss as ref_loss                        # noqa: E402

from oracle import ops_ref as R                              # noqa: E402
from oracle import networks_ref as NR                        # noqa: E402
from tests.util import patched_randn                         # noqa: E402

PIN_TOL = 2e-5
pins = []


def pin(name, got, want, tol=PIN_TOL):
    e = R.max_rel_err(got.detach(), want.detach())
    pins.append((name, e))
    assert e <= tol, f'ORACLE PIN FAILED {name}: {e:.3e} > {tol}'


def npy(t):
    return None if t is None else t.detach().cpu().numpy().copy()


def save(name, **arrays):
    arrays = {k: v for k, v in arrays.items() if v is not None}
    path = os.path.join(HERE, name + '.npz')
    np.savez(path, **arrays)
    print(f'  wrote {name}.npz  ({os.path.getsize(path) / 1024:.0f} KiB, {len(arrays)} arrays)')


# ----------------------------------------------------------------------------
def gen_upfirdn2d():
    g = torch.Generator().manual_seed(100)
    out = {}
    f2d = ref_upfirdn2d.setup_filter([1, 3, 3, 1])
    f1d12 = ref_upfirdn2d.setup_filter(np.array([0.015404109327027373, 0.0034907120842174702, -0.11799011114819057,
                                                 -0.048311742585633, 0.4910559419267466, 0.787641141030194,
                                                 0.3379294217276218, -0.07263752278646252, -0.021060292512300564,
                                                 0.04472490177066578, 0.0017677118642428036, -0.007800708325034148]))
    f3 = ref_upfirdn2d.setup_filter([1, 2, 1])
    f5x2 = torch.randn(2, 5, generator=g)
    pin('setup_filter[1331]', R.setup_filter([1, 3, 3, 1]), f2d, 0)
    pin('setup_filter[sym6]', R.setup_filter(f1d12.numpy() * 1.0, normalize=True), ref_upfirdn2d.setup_filter(f1d12.numpy()), 1e-7)
    cases = [
        # name, shape, f, up, down, padding, flip, gain
        ('up2_g4', (2, 3, 8, 8), f2d, 2, 1, [2, 1, 2, 1], False, 4),          # upsample2d of the RGB skip image
        ('down2', (2, 4, 16, 16), f2d, 1, 2, [1, 1, 1, 1], False, 1),          # D skip path
        ('filt_p1_g4', (2, 4, 17, 17), f2d, 1, 1, [1, 1, 1, 1], False, 4),     # after the up-conv
        ('filt_p2', (2, 4, 16, 16), f2d, 1, 1, [2, 2, 2, 2], False, 1),        # before the D stride-2 conv
        ('crop', (1, 2, 12, 10), f2d, 1, 1, [-1, 2, 0, -2], True, 1.5),        # negative padding = crop
        ('sep12_up2', (1, 3, 9, 7), f1d12, 2, 1, [6, 5, 6, 5], False, 4),      # ADA separable path
        ('sep12_down2', (1, 3, 20, 18), f1d12, 1, 2, [5, 5, 5, 5], True, 1),
        ('anis', (2, 2, 7, 9), f5x2, [2, 1], [1, 3], [3, 0, 1, 2], False, 0.7),
        ('f3_up3_down2', (1, 5, 6, 5), R.setup_filter([1, 2, 1]), 3, 2, [2, 2, 1, 3], False, 9),
        ('identity_f', (2, 3, 5, 5), None, 2, 1, 0, False, 1),
        ('single_px', (1, 1, 1, 1), f2d, 2, 1, [2, 1, 2, 1], False, 4),
    ]
    _ = f3
    meta = []
    for name, shape, f, up, down, pad, flip, gain in cases:
        x = torch.randn(*shape, generator=g, requires_grad=True)
        y = ref_upfirdn2d.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain, impl='ref')
        dy = torch.randn(y.shape, generator=g)
        dx, = torch.autograd.grad(y, x, dy)
        yo = R.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
        pin('upfirdn2d/' + name, yo, y)
        dxo, = torch.autograd.grad(yo, x, dy)
        pin('upfirdn2d/' + name + '/dx', dxo, dx)
        # the backward-as-forward identity, upfirdn2d.py:264-283
        if f is not None:
            bw = R.upfirdn2d_backward_args(x.shape, y.shape, f, up, down, pad)
            dxi = R.upfirdn2d(dy, f, flip_filter=(not flip), gain=gain, **bw)
            pin('upfirdn2d/' + name + '/dx_identity', dxi, dx)
        out[name + '.x'] = npy(x); out[name + '.y'] = npy(y); out[name + '.dy'] = npy(dy); out[name + '.dx'] = npy(dx)
        if f is not None:
            out[name + '.f'] = npy(f)
        ups, downs, pads = R.parse_scaling(up), R.parse_scaling(down), R.parse_padding(pad)
        meta.append(f'{name}|{ups[0]},{ups[1]}|{downs[0]},{downs[1]}|{pads[0]},{pads[1]},{pads[2]},{pads[3]}|{int(flip)}|{gain}')
    # convenience wrappers
    x = torch.randn(2, 3, 10, 10, generator=g)
    for wname, rf, of in [('filter2d', ref_upfirdn2d.filter2d, R.filter2d), ('upsample2d', ref_upfirdn2d.upsample2d, R.upsample2d),
                          ('downsample2d', ref_upfirdn2d.downsample2d, R.downsample2d)]:
        for f in (f2d, f1d12):
            yr = rf(x, f, impl='ref')
            pin(f'{wname}/{f.ndim}d', of(x, f), yr)
            out[f'wrap.{wname}.{f.ndim}d.y'] = npy(yr)
    out['wrap.x'] = npy(x); out['wrap.f2d'] = npy(f2d); out['wrap.f1d'] = npy(f1d12)
    out['meta'] = np.array(meta)
    save('upfirdn2d', **out)


# ----------------------------------------------------------------------------
def gen_bias_act():
    g = torch.Generator().manual_seed(200)
    out = {}
    meta = []
    shapes = {'4d': ((3, 5, 4, 6), 1), '2d': ((7, 12), 1), '3d_dim2': ((2, 3, 5), 2)}
    for act, (sname, (shape, dim)), clamp, use_b in itertools.product(
            R.ACTIVATIONS.keys(), shapes.items(), (None, 0.6), (True, False)):
        if sname != '4d' and (clamp is not None or not use_b):
            continue
        gain = None if sname == '4d' else 0.7
        alpha = None if sname != '2d' else 0.3
        name = f'{act}.{sname}.c{0 if clamp is None else 1}.b{int(use_b)}'
        x = (torch.randn(*shape, generator=g) * 1.5).requires_grad_(True)
        b = torch.randn(shape[dim], generator=g).requires_grad_(True) if use_b else None
        y = ref_bias_act.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp, impl='ref')
        yo = R.bias_act(x, b, dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp)
        pin('bias_act/' + name, yo, y, 1e-6)
        dy = torch.randn(shape, generator=g).requires_grad_(True)
        ins = [x] + ([b] if use_b else [])
        grads = torch.autograd.grad(y, ins, dy, create_graph=True)
        dx = grads[0]
        db = grads[1] if use_b else None
        # second order: differentiate <dx, d_dx> wrt dy and x
        d_dx = torch.randn(shape, generator=g)
        gg = torch.autograd.grad(dx, [dy, x], d_dx, allow_unused=True)
        out[name + '.x'] = npy(x); out[name + '.b'] = npy(b); out[name + '.y'] = npy(y)
        out[name + '.dy'] = npy(dy); out[name + '.dx'] = npy(dx); out[name + '.db'] = npy(db)
        out[name + '.d_dx'] = npy(d_dx); out[name + '.gg_dy'] = npy(gg[0])
        out[name + '.gg_x'] = npy(gg[1]) if gg[1] is not None else np.zeros(shape, np.float32)
        # the native kernel's explicit formulas (bias_act.cu:56-142) against autograd
        spec = R.ACTIVATIONS[act]
        a_ = float(alpha if alpha is not None else spec[0]); g_ = float(gain if gain is not None else spec[1])
        c_ = float(clamp if clamp is not None else -1)
        xb = (x + (b.reshape([-1 if i == dim else 1 for i in range(x.ndim)]) if use_b else 0)).detach()
        f1 = R.bias_act_grad_formula(1, act, dy.detach(), xb, y.detach(), None, a_, g_, c_)
        # where |y| hits the clamp exactly or x==0 the one-sided choices differ by measure zero; random data avoids it
        pin('bias_act/' + name + '/grad1_formula', f1, dx, 2e-5)
        if spec[4]:
            f2 = R.bias_act_grad_formula(2, act, d_dx, xb, y.detach(), dy.detach(), a_, g_, c_)
            pin('bias_act/' + name + '/grad2_formula', f2, gg[1], 5e-5)
        meta.append(f'{name}|{act}|{dim}|{alpha}|{gain}|{clamp}')
    out['meta'] = np.array(meta)
    save('bias_act', **out)


# ----------------------------------------------------------------------------
def gen_conv2d_resample():
    g = torch.Generator().manual_seed(300)
    out = {}
    meta = []
    f = ref_upfirdn2d.setup_filter([1, 3, 3, 1])
    cases = [
        # name, N, I, O, H, W, k, up, down, padding, flip_weight, groups
        ('plain3', 2, 8, 6, 9, 9, 3, 1, 1, 1, True, 1),
        ('plain3_noflip', 2, 8, 6, 8, 8, 3, 1, 1, 1, False, 1),
        ('plain1', 2, 8, 3, 8, 8, 1, 1, 1, 0, True, 1),
        ('up2_k3', 2, 8, 6, 8, 8, 3, 2, 1, 1, False, 1),
        ('up2_k3_flip', 1, 4, 4, 5, 5, 3, 2, 1, 1, True, 1),
        ('down2_k3', 2, 6, 8, 16, 16, 3, 1, 2, 1, True, 1),
        ('down2_k1', 2, 6, 8, 16, 16, 1, 1, 2, 0, True, 1),
        ('up2_k1', 2, 6, 4, 8, 8, 1, 2, 1, 0, True, 1),
        ('generic_asym', 1, 4, 4, 9, 9, 3, 1, 1, [1, 0, 2, 1], True, 1),
        ('grouped_plain', 1, 12, 9, 8, 8, 3, 1, 1, 1, True, 3),
        ('grouped_up2', 1, 12, 9, 8, 8, 3, 2, 1, 1, False, 3),
        ('tc_plain3', 2, 32, 32, 16, 16, 3, 1, 1, 1, True, 1),       # tensor-core eligible shapes
        ('tc_plain1', 2, 32, 16, 16, 16, 1, 1, 1, 0, True, 1),
        ('tc_up2', 2, 32, 16, 16, 16, 3, 2, 1, 1, False, 1),
        ('tc_down2', 2, 16, 32, 32, 32, 3, 1, 2, 1, True, 1),
    ]
    for name, N, I, O, H, W, k, up, down, pad, flipw, groups in cases:
        x = torch.randn(N, I, H, W, generator=g, requires_grad=True)
        w = (torch.randn(O, I // groups, k, k, generator=g) / np.sqrt(I // groups * k * k)).requires_grad_(True)
        y = ref_c2r.conv2d_resample(x, w, f=f, up=up, down=down, padding=pad, groups=groups, flip_weight=flipw)
        yo = R.conv2d_resample(x, w, f=f, up=up, down=down, padding=pad, groups=groups, flip_weight=flipw)
        pin('conv2d_resample/' + name, yo, y)
        dy = torch.randn(y.shape, generator=g)
        dx, dw = torch.autograd.grad(y, [x, w], dy)
        out[name + '.x'] = npy(x); out[name + '.w'] = npy(w); out[name + '.y'] = npy(y)
        out[name + '.dy'] = npy(dy); out[name + '.dx'] = npy(dx); out[name + '.dw'] = npy(dw)
        pp = R.parse_padding(pad)
        meta.append(f'{name}|{up}|{down}|{pp[0]},{pp[1]},{pp[2]},{pp[3]}|{int(flipw)}|{groups}')
    out['f'] = npy(f)
    out['meta'] = np.array(meta)
    save('conv2d_resample', **out)


# ----------------------------------------------------------------------------
def gen_modconv():
    g = torch.Generator().manual_seed(400)
    out = {}
    meta = []
    f = ref_upfirdn2d.setup_filter([1, 3, 3, 1])
    cases = [
        # name, N, I, O, R_in, k, up, demod, noise('none'|'rand'|'const'), fused
        ('k3', 3, 8, 6, 8, 3, 1, True, 'rand', False),
        ('k3_fused', 3, 8, 6, 8, 3, 1, True, 'const', True),
        ('k3_up2', 2, 8, 6, 8, 3, 2, True, 'rand', False),
        ('k3_up2_fused', 2, 8, 6, 8, 3, 2, True, 'none', True),
        ('torgb', 2, 8, 3, 8, 1, 1, False, 'none', False),
        ('torgb_fused', 2, 8, 3, 8, 1, 1, False, 'none', True),
        ('nodemod_noise', 2, 4, 4, 4, 3, 1, False, 'rand', False),
        ('tc_k3', 2, 32, 32, 16, 3, 1, True, 'rand', False),
        ('tc_k3_up2', 2, 32, 16, 16, 3, 2, True, 'const', False),
        ('tc_torgb', 2, 32, 3, 16, 1, 1, False, 'none', False),
        ('b4', 4, 16, 16, 4, 3, 1, True, 'rand', False),
    ]
    for name, N, I, O, Rin, k, up, demod, noise_kind, fused in cases:
        Rout = Rin * up
        x = torch.randn(N, I, Rin, Rin, generator=g, requires_grad=True)
        w = torch.randn(O, I, k, k, generator=g, requires_grad=True)
        s = (torch.randn(N, I, generator=g) * 0.5 + 1).requires_grad_(True)
        noise = None
        if noise_kind == 'rand':
            noise = (torch.randn(N, 1, Rout, Rout, generator=g) * 0.3).requires_grad_(True)
        if noise_kind == 'const':
            noise = (torch.randn(Rout, Rout, generator=g) * 0.3).requires_grad_(True)
        kw = dict(noise=noise, up=up, padding=k // 2, resample_filter=f, demodulate=demod, flip_weight=(up == 1),
                  fused_modconv=fused)
        y = ref_networks.modulated_conv2d(x=x, weight=w, styles=s, **kw)
        yo = R.modulated_conv2d(x, w, s, **kw)
        pin('modulated_conv2d/' + name, yo, y)
        dy = torch.randn(y.shape, generator=g)
        ins = [x, w, s] + ([noise] if noise is not None else [])
        grads = torch.autograd.grad(y, ins, dy)
        out[name + '.x'] = npy(x); out[name + '.w'] = npy(w); out[name + '.s'] = npy(s); out[name + '.noise'] = npy(noise)
        out[name + '.y'] = npy(y); out[name + '.dy'] = npy(dy)
        out[name + '.dx'] = npy(grads[0]); out[name + '.dw'] = npy(grads[1]); out[name + '.ds'] = npy(grads[2])
        if noise is not None:
            out[name + '.dnoise'] = npy(grads[3])
        meta.append(f'{name}|{k}|{up}|{int(demod)}|{noise_kind}|{int(fused)}')
    # fma (fma.py:15-58)
    a = torch.randn(2, 4, 5, 5, generator=g); b = torch.randn(2, 4, 1, 1, generator=g); c = torch.randn(2, 1, 5, 5, generator=g)
    pin('fma', R.fma(a, b, c), ref_fma.fma(a, b, c), 1e-7)
    out['fma.a'] = npy(a); out['fma.b'] = npy(b); out['fma.c'] = npy(c); out['fma.y'] = npy(ref_fma.fma(a, b, c))
    out['f'] = npy(f)
    out['meta'] = np.array(meta)
    save('modconv', **out)


# ----------------------------------------------------------------------------
NET = dict(res=64, channel_base=2048, channel_max=32, z_dim=64, w_dim=64, num_layers=2, mbstd=4, batch=4)


def build_ref_nets(PG, PD):
    c = NET
    G = ref_networks.Generator(z_dim=c['z_dim'], c_dim=0, w_dim=c['w_dim'], img_resolution=c['res'], img_channels=3,
                               mapping_kwargs=dict(num_layers=c['num_layers']),
                               synthesis_kwargs=dict(channel_base=c['channel_base'], channel_max=c['channel_max'],
                                                     num_fp16_res=0, conv_clamp=None))
    D = ref_networks.Discriminator(c_dim=0, img_resolution=c['res'], img_channels=3, channel_base=c['channel_base'],
                                   channel_max=c['channel_max'], num_fp16_res=0, conv_clamp=None,
                                   epilogue_kwargs=dict(mbstd_group_size=c['mbstd']))
    for net, P in ((G, PG), (D, PD)):
        sd = net.state_dict()
        missing = [k for k in sd if k not in P and not k.endswith('resample_filter')]
        extra = [k for k in P if k not in sd]
        assert not missing and not extra, (missing, extra)
        for k, v in P.items():
            assert sd[k].shape == v.shape, (k, sd[k].shape, v.shape)
        net.load_state_dict({k: v.clone() for k, v in P.items()}, strict=False)
    return G, D


def gen_networks():
    c = NET
    g = torch.Generator().manual_seed(500)
    PG = NR.init_G_params(c['res'], c['channel_base'], c['channel_max'], c['z_dim'], c['w_dim'], c['num_layers'],
                          generator=g, randomize=True)
    PD = NR.init_D_params(c['res'], c['channel_base'], c['channel_max'], generator=g, randomize=True)
    G, D = build_ref_nets(PG, PD)
    out = {'G.' + k: npy(v) for k, v in PG.items()}
    out.update({'D.' + k: npy(v) for k, v in PD.items()})
    z = torch.randn(c['batch'], c['z_dim'], generator=g)
    real = torch.rand(c['batch'], 3, c['res'], c['res'], generator=g) * 2 - 1
    cnone = torch.zeros(c['batch'], 0)
    out['z'] = npy(z); out['real'] = npy(real)

    # --- forward, eval mode (fused_modconv=True), noise const -------------------------------------------
    G.eval(); D.eval()
    with torch.no_grad():
        ws = G.mapping(z, cnone)
        img = G.synthesis(ws, noise_mode='const')
        img_trunc = G(z, cnone, truncation_psi=0.7, truncation_cutoff=4, noise_mode='const')
        logits = D(img, cnone)
        ws_o = NR.mapping(PG, z, G.num_ws, c['num_layers'])
        img_o = NR.synthesis(PG, ws_o, c['res'], noise_mode='const', fused_modconv=True)
        logits_o = NR.discriminator(PD, img, c['res'], c['mbstd'])
    pin('net/mapping', ws_o, ws); pin('net/synthesis_eval', img_o, img); pin('net/D_eval', logits_o, logits)
    out['eval.ws'] = npy(ws); out['eval.img'] = npy(img); out['eval.logits'] = npy(logits); out['eval.img_trunc'] = npy(img_trunc)

    # --- forward, train mode (non-fused), random noise through the patched RNG ---------------------------
    G.train(); D.train()
    G.mapping.w_avg_beta = None  # keep w_avg untouched; the EMA is a caller-side side effect
    with torch.no_grad(), patched_randn(7):
        img_t = G.synthesis(ws, noise_mode='random')
    with torch.no_grad(), patched_randn(7):
        img_to = NR.synthesis(PG, ws, c['res'], noise_mode='random', fused_modconv=False)
    pin('net/synthesis_train_randnoise', img_to, img_t)
    out['train.img_randnoise7'] = npy(img_t)

    # --- the four loss phases through the reference's own StyleGAN2Loss ---------------------------------
    for p in list(G.parameters()) + list(D.parameters()):
        p.requires_grad_(True)
    loss = ref_loss.StyleGAN2Loss(device=torch.device('cpu'), G_mapping=G.mapping, G_synthesis=G.synthesis, D=D,
                                  augment_pipe=None, style_mixing_prob=0, r1_gamma=10, pl_batch_shrink=2,
                                  pl_decay=0.01, pl_weight=2)

    def grads_of(net):
        return {k: (p.grad.detach().clone() if p.grad is not None else torch.zeros_like(p)) for k, p in net.named_parameters()}

    def zero(net):
        for p in net.parameters():
            p.grad = None

    def oracle_params(P):
        return {k: v.clone().requires_grad_(True) for k, v in P.items()}

    phases = {}
    for phase in ['Gmain', 'Greg', 'Dmain', 'Dreg']:
        zero(G); zero(D)
        loss.pl_mean.zero_()
        with patched_randn(11):
            loss.accumulate_gradients(phase=phase, real_img=real, real_c=cnone, gen_z=z, gen_c=cnone, sync=True, gain=1.0)
        net = G if phase[0] == 'G' else D
        phases[phase] = grads_of(net)
        for k, v in phases[phase].items():
            out[f'{phase}.grad.{k}'] = npy(v)
        if phase == 'Greg':
            out['Greg.pl_mean'] = npy(loss.pl_mean)
        # oracle pin of the same phase
        OG, OD = oracle_params(PG), oracle_params(PD)
        with patched_randn(11):
            if phase == 'Gmain':
                l = NR.loss_Gmain(OG, OD, z, c['res'], c['mbstd'], num_layers=c['num_layers']); tgt = OG
            elif phase == 'Greg':
                l, _ = NR.loss_Gpl(OG, z, c['res'], torch.zeros([]), num_layers=c['num_layers']); tgt = OG
            elif phase == 'Dmain':
                l = NR.loss_Dmain(OG, OD, z, real, c['res'], c['mbstd'], num_layers=c['num_layers']); tgt = OD
            else:
                l = NR.loss_Dr1(OD, real, c['res'], 10.0, c['mbstd']); tgt = OD
        leaves = [v for k, v in tgt.items()]
        og = torch.autograd.grad(l, leaves, allow_unused=True)
        worst = 0.0
        for (k, _), gk in zip(tgt.items(), og):
            if k not in phases[phase]:
                continue
            want = phases[phase][k]
            got = gk if gk is not None else torch.zeros_like(want)
            if want.abs().max() == 0 and got.abs().max() == 0:
                continue
            worst = max(worst, R.max_rel_err(got, want))
        pins.append((f'net/loss/{phase}', worst))
        assert worst <= 2e-4, (phase, worst)
    out['meta'] = np.array([f'{k}={v}' for k, v in c.items()])
    save('networks', **out)


if __name__ == '__main__':
    print('generating goldens from the live reference at', REF)
    gen_upfirdn2d()
    gen_bias_act()
    gen_conv2d_resample()
    gen_modconv()
    gen_networks()
    worst = max(pins, key=lambda t: t[1])
    print(f'oracle pinned against the live reference on {len(pins)} checks; worst max-rel-err {worst[1]:.3e} ({worst[0]})')
    with open(os.path.join(HERE, 'PIN_REPORT.txt'), 'w') as fh:
        fh.write(f'torch {torch.__version__}; {len(pins)} oracle-vs-reference checks, all within tolerance\n')
        for n, e in pins:
            fh.write(f'{e:.3e}  {n}\n')
