"""CPU: the oracle (oracle/) reproduces the committed goldens that were generated from the live reference
by tests/golden/make_golden.py.  This is what keeps the oracle pinned on boxes without /root/reference."""
import numpy as np
import pytest
import torch

from tests.util import load_golden, t, assert_close
from oracle import ops_ref as R
from oracle import networks_ref as NR

PIN = 2e-5


def _meta(g):
    return [str(m).split('|') for m in g['meta']]


def test_upfirdn2d_cases():
    g = load_golden('upfirdn2d')
    for name, up, down, pad, flip, gain in _meta(g):
        up = [int(v) for v in up.split(',')]; down = [int(v) for v in down.split(',')]; pad = [int(v) for v in pad.split(',')]
        f = t(g[name + '.f']) if name + '.f' in g else None
        x = t(g[name + '.x'], requires_grad=True)
        y = R.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=bool(int(flip)), gain=float(gain))
        assert_close(y, g[name + '.y'], PIN, name)
        dx, = torch.autograd.grad(y, x, t(g[name + '.dy']))
        assert_close(dx, g[name + '.dx'], PIN, name + '.dx')
        # integer output-size rule (upfirdn2d.cpp:32-33), bit-exact
        fw, fh = R.filter_size(f)
        assert y.shape[3] == R.upfirdn2d_out_size(x.shape[3], up[0], pad[0], pad[1], fw, down[0])
        assert y.shape[2] == R.upfirdn2d_out_size(x.shape[2], up[1], pad[2], pad[3], fh, down[1])


def test_upfirdn2d_wrappers():
    g = load_golden('upfirdn2d')
    x = t(g['wrap.x'])
    for fname, f in (('2d', t(g['wrap.f2d'])), ('1d', t(g['wrap.f1d']))):
        assert_close(R.filter2d(x, f), g[f'wrap.filter2d.{fname}.y'], PIN)
        assert_close(R.upsample2d(x, f), g[f'wrap.upsample2d.{fname}.y'], PIN)
        assert_close(R.downsample2d(x, f), g[f'wrap.downsample2d.{fname}.y'], PIN)


def test_bias_act_cases():
    g = load_golden('bias_act')
    for name, act, dim, alpha, gain, clamp in _meta(g):
        kw = dict(dim=int(dim), act=act, alpha=None if alpha == 'None' else float(alpha),
                  gain=None if gain == 'None' else float(gain), clamp=None if clamp == 'None' else float(clamp))
        x = t(g[name + '.x'], requires_grad=True)
        b = t(g[name + '.b'], requires_grad=True) if name + '.b' in g else None
        y = R.bias_act(x, b, **kw)
        assert_close(y, g[name + '.y'], 1e-6, name)
        dy = t(g[name + '.dy'], requires_grad=True)
        grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), dy, create_graph=True)
        assert_close(grads[0], g[name + '.dx'], PIN, name + '.dx')
        if b is not None:
            assert_close(grads[1], g[name + '.db'], PIN, name + '.db')
        gg = torch.autograd.grad(grads[0], [dy, x], t(g[name + '.d_dx']), allow_unused=True)
        assert_close(gg[0], g[name + '.gg_dy'], PIN, name + '.gg_dy')
        assert_close(gg[1] if gg[1] is not None else torch.zeros_like(x), g[name + '.gg_x'], 5e-5, name + '.gg_x')


def test_conv2d_resample_cases():
    g = load_golden('conv2d_resample')
    f = t(g['f'])
    for name, up, down, pad, flipw, groups in _meta(g):
        x = t(g[name + '.x'], requires_grad=True); w = t(g[name + '.w'], requires_grad=True)
        y = R.conv2d_resample(x, w, f=f, up=int(up), down=int(down), padding=[int(v) for v in pad.split(',')],
                              groups=int(groups), flip_weight=bool(int(flipw)))
        assert_close(y, g[name + '.y'], PIN, name)
        dx, dw = torch.autograd.grad(y, [x, w], t(g[name + '.dy']))
        assert_close(dx, g[name + '.dx'], PIN, name + '.dx')
        assert_close(dw, g[name + '.dw'], PIN, name + '.dw')


def test_modulated_conv2d_cases():
    g = load_golden('modconv')
    f = t(g['f'])
    for name, k, up, demod, noise_kind, fused in _meta(g):
        x = t(g[name + '.x'], requires_grad=True); w = t(g[name + '.w'], requires_grad=True); s = t(g[name + '.s'], requires_grad=True)
        noise = t(g[name + '.noise'], requires_grad=True) if name + '.noise' in g else None
        y = R.modulated_conv2d(x, w, s, noise=noise, up=int(up), padding=int(k) // 2, resample_filter=f,
                               demodulate=bool(int(demod)), flip_weight=(int(up) == 1), fused_modconv=bool(int(fused)))
        assert_close(y, g[name + '.y'], PIN, name)
        grads = torch.autograd.grad(y, [x, w, s] + ([noise] if noise is not None else []), t(g[name + '.dy']))
        for gname, got in zip(['dx', 'dw', 'ds', 'dnoise'], grads):
            assert_close(got, g[f'{name}.{gname}'], 5e-5, f'{name}.{gname}')
    assert_close(R.fma(t(g['fma.a']), t(g['fma.b']), t(g['fma.c'])), g['fma.y'], 1e-7)


def _net_cfg(g):
    return {kv.split('=')[0]: int(kv.split('=')[1]) for kv in (str(m) for m in g['meta'])}


def test_networks_forward_and_one_loss_phase():
    g = load_golden('networks')
    c = _net_cfg(g)
    PG = {k[2:]: t(v) for k, v in g.items() if k.startswith('G.')}
    PD = {k[2:]: t(v) for k, v in g.items() if k.startswith('D.')}
    z = t(g['z'])
    with torch.no_grad():
        ws = NR.mapping(PG, z, NR.num_ws_for(c['res']), c['num_layers'])
        assert_close(ws, g['eval.ws'], PIN, 'mapping')
        img = NR.synthesis(PG, ws, c['res'], noise_mode='const', fused_modconv=True)
        assert_close(img, g['eval.img'], PIN, 'synthesis')
        assert_close(NR.discriminator(PD, t(g['eval.img']), c['res'], c['mbstd']), g['eval.logits'], PIN, 'D')
    # Dreg (R1, double backward) parameter gradients
    OD = {k: v.clone().requires_grad_(True) for k, v in PD.items()}
    l = NR.loss_Dr1(OD, t(g['real']), c['res'], 10.0, c['mbstd'])
    names = list(OD.keys())
    grads = torch.autograd.grad(l, [OD[k] for k in names], allow_unused=True)
    for k, gk in zip(names, grads):
        want = g['Dreg.grad.' + k]
        if np.abs(want).max() == 0:
            continue
        assert_close(gk, want, 2e-4, 'Dreg.' + k)
