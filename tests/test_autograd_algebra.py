"""CPU tests of the autograd algebra of conv2d_gradfix (first, second and third order) with the kernels replaced by a torch stand-in
(tests/fake_plugin.py) in fp64: every node's backward formula is checked against autograd over the plain torch expression
y = b * conv(a * x, w).  The kernels themselves are tested on the GPU (test_gpu_ops.py); this file pins the calculus."""
import itertools
import numpy as np
import pytest
import torch

import tests.util  # noqa: F401  (installs the operator modules)
from tests.fake_plugin import FakePlugin
from torch_utils.ops import conv2d_gradfix as cg


@pytest.fixture()
def fake_plugin():
    old = cg._plugin
    cg._plugin = FakePlugin()
    yield cg._plugin
    cg._plugin = old


def _ref(x, w, a, b, pad, io, flip):
    v = w.transpose(0, 1) if io else w
    if flip:
        v = v.flip([2, 3])
    xs = x * a[:, :, None, None] if a is not None else x
    y = torch.nn.functional.conv2d(xs, v, padding=pad)
    return y * b[:, :, None, None] if b is not None else y


def _mine(x, w, a, b, pad, io, flip):
    kh, kw = int(w.shape[2]), int(w.shape[3])
    out_hw = (int(x.shape[2]) + 2 * pad - kh + 1, int(x.shape[3]) + 2 * pad - kw + 1)
    if a is None and b is None:
        return cg._conv2d_s1(tuple(w.shape), (pad, pad), out_hw, io, flip).apply(x, w)
    return cg._scaled_conv2d_s1(tuple(w.shape), (pad, pad), out_hw, io, flip, 1.0, None, a is not None, b is not None).apply(x, w, a, b)


def _orders(conv, x, w, a, b, r, which, closed=True):
    cg.closed_scaled_backward = closed
    try:
        ts = [t.clone().requires_grad_(True) if t is not None else None for t in (x, w, a, b)]
        live = [t for t in ts if t is not None]
        y = conv(*ts)
        first = torch.autograd.grad((y * r).sum(), live, create_graph=True)
        pen = sum(first[i].square().sum() * (1.0 + 0.3 * i) for i in which if i < len(first))
        second = torch.autograd.grad(pen, live, create_graph=True, allow_unused=True)
        pen2 = sum(s.square().sum() for s in second if s is not None)
        third = torch.autograd.grad(pen2, live, allow_unused=True, retain_graph=True)
        second_plain = torch.autograd.grad(pen, live, allow_unused=True)          # the non-create_graph route (fused nodes)
        return y, first, second, third, second_plain
    finally:
        cg.closed_scaled_backward = True


def _close(u, v, tol, what):
    if u is None or v is None:
        assert (u is None or float(u.abs().max()) == 0) and (v is None or float(v.abs().max()) == 0), what + ': one side has no gradient'
        return
    err = float((u - v).detach().abs().max()) / max(float(v.detach().abs().max()), 1e-300)
    assert err <= tol, f'{what}: max-rel-err {err:.3e}'


@pytest.mark.parametrize('k,pad,io,flip', [(3, 1, False, False), (3, 1, True, True), (2, 1, False, True), (1, 0, False, False), (2, 0, True, False)])
@pytest.mark.parametrize('scales', ['ab', 'a', 'b'])
def test_scaled_conv_calculus_up_to_third_order(fake_plugin, k, pad, io, flip, scales):
    g = torch.Generator().manual_seed(5 + k)
    N, I, O, H = 2, 4, 3, 6
    x = torch.randn(N, I, H, H, generator=g, dtype=torch.float64)
    w = torch.randn(*((I, O) if io else (O, I)), k, k, generator=g, dtype=torch.float64) / np.sqrt(I * k * k)
    a = (torch.randn(N, I, generator=g, dtype=torch.float64) + 1.5) if 'a' in scales else None
    b = (torch.rand(N, O, generator=g, dtype=torch.float64) + 0.5) if 'b' in scales else None
    r = torch.randn(N, O, H + 2 * pad - k + 1, H + 2 * pad - k + 1, generator=g, dtype=torch.float64)
    n_in = 2 + len(scales)
    for which in [[0], [1], list(range(n_in))] + ([[2]] if n_in > 2 else []) + ([[3]] if n_in > 3 else []):
        want = _orders(lambda *t: _ref(*t, pad, io, flip), x, w, a, b, r, which)
        for closed in (True, False):
            got = _orders(lambda *t: _mine(*t, pad, io, flip), x, w, a, b, r, which, closed=closed)
            tag = f'k{k} pad{pad} io{io} flip{flip} scales={scales} penalty on {which} closed={closed}'
            _close(got[0], want[0], 1e-12, tag + ' y')
            for lvl, name in ((1, 'first'), (2, 'second (create_graph)'), (3, 'third'), (4, 'second')):
                for c, u, v in zip('xwab'[:2] + scales, got[lvl], want[lvl]):
                    _close(u, v, 1e-9, f'{tag}: {name} d{c}')


def test_unscaled_conv_calculus_up_to_third_order(fake_plugin):
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, 5, 5, generator=g, dtype=torch.float64); w = torch.randn(4, 3, 3, 3, generator=g, dtype=torch.float64) * 0.3
    r = torch.randn(2, 4, 5, 5, generator=g, dtype=torch.float64)
    for which in ([0], [1], [0, 1]):
        want = _orders(lambda *t: _ref(*t, 1, False, False), x, w, None, None, r, which)
        got = _orders(lambda *t: _mine(*t, 1, False, False), x, w, None, None, r, which)
        for lvl in (1, 2, 3, 4):
            for c, u, v in zip('xw', got[lvl], want[lvl]):
                _close(u, v, 1e-9, f'unscaled, penalty on {which}, level {lvl}, d{c}')


@pytest.mark.parametrize('k,demodulate', [(3, True), (1, False), (3, False)])
def test_modulated_conv2d_against_the_live_reference_on_cpu(fake_plugin, monkeypatch, k, demodulate):
    """The PRODUCT's modulated_conv2d (styles and demodulation coefficients as in-kernel scales, dcoefs as a small GEMM) against the
    reference's own modulated_conv2d (networks.py:591-668, non-fused and fused form) in fp32 on the CPU (the product serves fp32 only), kernels replaced by the torch
    stand-in: output, gradients w.r.t. x / weight / styles, and the path-length style second-order gradient.  Here the output scale
    is itself a function of the input scale (dcoefs = rsqrt(styles^2 @ wsq)), so this is the check that no path is counted twice."""
    from oracle import live_ref
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent')
    L = live_ref.load()
    from gagan_b200.training import networks as mine
    monkeypatch.setattr(cg, '_check_input', lambda t: None)           # (the device check; the stand-in runs on CPU tensors)
    g = torch.Generator().manual_seed(31 + k)
    N, I, O, H = 3, 5, 4, 6
    x = torch.randn(N, I, H, H, generator=g, dtype=torch.float32)
    w = torch.randn(O, I, k, k, generator=g, dtype=torch.float32)
    s = torch.randn(N, I, generator=g, dtype=torch.float32) * 0.5 + 1
    noise = torch.randn(N, 1, H, H, generator=g, dtype=torch.float32) * 0.1
    r = torch.randn(N, O, H, H, generator=g, dtype=torch.float32)

    def run(fn, **extra):
        ts = [t.clone().requires_grad_(True) for t in (x, w, s)]
        y = fn(x=ts[0], weight=ts[1], styles=ts[2], noise=noise, up=1, padding=k // 2, demodulate=demodulate, flip_weight=True, **extra)
        first = torch.autograd.grad((y * r).sum(), ts, create_graph=True)
        pen = first[2].square().sum() + first[0].square().mean()
        second = torch.autograd.grad(pen, ts, allow_unused=True)
        return [y] + list(first) + list(second)

    mine_out = run(mine.modulated_conv2d.__wrapped__ if hasattr(mine.modulated_conv2d, '__wrapped__') else mine.modulated_conv2d)
    for fused in (False, True):
        want = run(L.networks.modulated_conv2d, fused_modconv=fused)
        for name, u, v in zip(('y', 'dx', 'dw', 'ds', 'ddx', 'ddw', 'dds'), mine_out, want):
            _close(u, v, 2e-4 if name.startswith('dd') else 2e-5, f'k{k} demodulate={demodulate} reference fused_modconv={fused}: {name}')


@pytest.mark.parametrize('case', [
    # name, I, O, R, up, down, k, flip_weight  -- every branch of conv2d_resample the networks take
    ('G conv0 up', 6, 5, 8, 2, 1, 3, False), ('D conv1 down', 5, 6, 16, 1, 2, 3, True), ('D skip down 1x1', 5, 4, 16, 1, 2, 1, True),
    ('plain 3x3', 4, 4, 8, 1, 1, 3, True), ('ToRGB 1x1', 6, 3, 8, 1, 1, 1, True), ('G conv0 up, wide', 4, 4, 16, 2, 1, 3, False),
])
def test_conv2d_resample_against_the_live_reference_on_cpu(fake_plugin, monkeypatch, case):
    """The product's conv2d_resample (stride-2 layers as stride-1 convolutions over phase-major tensors, FIRs fused with the re-layout,
    every node an autograd Function of this build) against the reference's conv2d_resample (conv2d_resample.py:59-156) on the CPU,
    kernels replaced by the torch stand-in: output, gradients, and the R1-style second-order gradient."""
    from oracle import live_ref
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent')
    L = live_ref.load()
    from torch_utils.ops import conv2d_resample as CR, upfirdn2d as U
    name, I, O, R, up, down, k, flip_weight = case
    monkeypatch.setattr(cg, '_check_input', lambda t: None)
    monkeypatch.setattr(U, '_plugin', fake_plugin)
    monkeypatch.setattr(U, '_check_input', lambda t: None)
    g = torch.Generator().manual_seed(sum(map(ord, name)) % 997)
    x = torch.randn(2, I, R, R, generator=g); w = torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)
    f = U.setup_filter([1, 3, 3, 1])
    kw = dict(f=f, up=up, down=down, padding=k // 2, flip_weight=flip_weight)

    def run(fn):
        ts = [t.clone().requires_grad_(True) for t in (x, w)]
        y = fn(ts[0], ts[1], **kw)
        gen = torch.Generator().manual_seed(3)
        first = torch.autograd.grad((y * torch.randn(y.shape, generator=gen)).sum(), ts, create_graph=True)
        second = torch.autograd.grad(first[0].square().sum() + first[1].square().sum(), ts)
        return [y] + list(first) + list(second)

    got, want = run(CR.conv2d_resample.__wrapped__ if hasattr(CR.conv2d_resample, '__wrapped__') else CR.conv2d_resample), run(L.conv2d_resample.conv2d_resample)
    for nm, u, v in zip(('y', 'dx', 'dw', 'ddx', 'ddw'), got, want):
        assert u.shape == v.shape, (name, nm, u.shape, v.shape)
        _close(u, v, 2e-4 if nm.startswith('dd') else 2e-5, f'{name}: {nm}')


@pytest.mark.parametrize('act', ['linear', 'relu', 'lrelu', 'tanh', 'sigmoid', 'elu', 'selu', 'softplus', 'swish'])
@pytest.mark.parametrize('clamp', [None, 0.7])
def test_bias_act_autograd_against_the_live_reference_on_cpu(fake_plugin, monkeypatch, act, clamp):
    """bias_act.py's autograd Functions (forward, the closed gradient op with the fused bias reduction, its second-order backward, the
    noise extension) against the reference's bias_act (bias_act.py:36-88, impl='ref' on CPU), kernels replaced by the stand-in that
    evaluates the native kernel's formulas (bias_act.cu:40-142 as restated in oracle/ops_ref.py)."""
    from oracle import live_ref
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent')
    L = live_ref.load()
    from torch_utils.ops import bias_act as BA
    monkeypatch.setattr(BA, '_plugin', fake_plugin)
    g = torch.Generator().manual_seed(12)
    x = torch.randn(3, 5, 4, 6, generator=g); b = torch.randn(5, generator=g) * 0.5
    r = torch.randn(3, 5, 4, 6, generator=g)
    spec = BA.activation_funcs[act]

    def run(fn):
        ts = [t.clone().requires_grad_(True) for t in (x, b)]
        y = fn(ts[0], ts[1])
        first = torch.autograd.grad((y * r).sum(), ts, create_graph=True)
        pen = first[0].square().sum() + 2.0 * first[1].square().sum()
        second = torch.autograd.grad(pen, ts, allow_unused=True) if pen.requires_grad else (None, None)     # (linear: constant gradients)
        return [y] + list(first) + list(second)

    mine = BA._bias_act_cuda(dim=1, act=act, alpha=None, gain=None, clamp=clamp)
    got = run(lambda x_, b_: mine.apply(x_, b_))
    want = run(lambda x_, b_: L.bias_act.bias_act(x_, b_, act=act, clamp=clamp, impl='ref'))
    for nm, u, v in zip(('y', 'dx', 'db', 'ddx', 'ddb'), got, want):
        _close(u, v, 5e-5, f'{act} clamp={clamp}: {nm}')
    # the noise extension == an explicit add in front of the reference op
    noise = torch.randn(3, 1, 4, 6, generator=g) * 0.3
    if 'x' in spec.ref or spec.has_2nd_grad:
        got_n = run(lambda x_, b_: mine.apply(x_ + noise, b_))
    else:
        got_n = run(lambda x_, b_: mine.apply(x_, b_, noise))
    want_n = run(lambda x_, b_: L.bias_act.bias_act(x_ + noise, b_, act=act, clamp=clamp, impl='ref'))
    for nm, u, v in zip(('y', 'dx', 'db', 'ddx', 'ddb'), got_n, want_n):
        _close(u, v, 5e-5, f'{act} clamp={clamp} with noise: {nm}')


@pytest.fixture()
def host_layer_on_cpu(fake_plugin, monkeypatch):
    """Everything the operator modules need to run on CPU tensors: the stand-in behind all three plugin names, and the three device
    checks of the public entry points taken out.  The public functions themselves (argument parsing, fp16 entry, autograd) are the
    product's."""
    from torch_utils import custom_ops
    from torch_utils.ops import bias_act as BA, upfirdn2d as U
    for name in ('bias_act_plugin', 'upfirdn2d_plugin', 'conv2d_plugin'):
        monkeypatch.setitem(custom_ops._cached_plugins, name, fake_plugin)
    monkeypatch.setattr(BA, '_plugin', fake_plugin); monkeypatch.setattr(U, '_plugin', fake_plugin)
    # the device checks of the public entry points exist so that the PRODUCT never computes on the CPU; here that is the point
    for mod in (cg, U, BA):
        monkeypatch.setattr(mod, '_check_input', lambda t: None)
    return fake_plugin


@pytest.mark.parametrize('fused_callers', [True, False])
def test_reference_networks_on_the_host_layer_with_stand_in_kernels(host_layer_on_cpu, fused_callers):
    """End to end on the CPU: the reference's OWN Generator / Discriminator (from the installed checkout) run on this build's host
    layer -- install(), the fused forwards, modulated_conv2d, conv2d_resample with the phase-major stride-2 forms, upfirdn2d, bias_act
    -- with every kernel replaced by the torch stand-in, against the live reference on its impl='ref' ops: image, logits, and the
    gradients of a non-saturating G loss and of an R1 penalty."""
    from oracle import live_ref
    if not live_ref.available() or not tests.util.HAVE_CHECKOUT:
        pytest.skip('the reference checkouts are absent')
    L = live_ref.load()
    networks = tests.util.reference_networks()
    from gagan_b200.training import networks as host_networks
    host_networks.attach(networks, fused_callers=fused_callers)          # with False the reference's own layer forwards run untouched

    kw_g = dict(z_dim=16, c_dim=0, w_dim=16, img_resolution=16, img_channels=3, mapping_kwargs=dict(num_layers=2),
                synthesis_kwargs=dict(channel_base=256, channel_max=16))
    kw_d = dict(c_dim=0, img_resolution=16, img_channels=3, channel_base=256, channel_max=16, epilogue_kwargs=dict(mbstd_group_size=2))
    torch.manual_seed(4)
    G_ref, D_ref = tests.util.quiet(L.networks.Generator, **kw_g).train(), tests.util.quiet(L.networks.Discriminator, **kw_d).train()
    with torch.no_grad():
        for p_ in list(G_ref.parameters()) + list(D_ref.parameters()):
            if float(p_.abs().max()) == 0:
                p_.copy_(torch.randn(p_.shape) * 0.1)
    G, D = tests.util.quiet(networks.Generator, **kw_g).train(), tests.util.quiet(networks.Discriminator, **kw_d).train()
    G.load_state_dict(G_ref.state_dict()); D.load_state_dict(D_ref.state_dict())
    z = torch.randn(4, 16, generator=torch.Generator().manual_seed(2)); c = torch.zeros(4, 0)
    real = torch.rand(4, 3, 16, 16, generator=torch.Generator().manual_seed(3)) * 2 - 1

    def run(Gn, Dn):
        for p_ in list(Gn.parameters()) + list(Dn.parameters()):
            p_.grad = None
        img = Gn(z, c, noise_mode='const')
        logits = Dn(img, c)
        torch.nn.functional.softplus(-logits).mean().backward()
        g_grads = {k: p_.grad.clone() for k, p_ in Gn.named_parameters() if p_.grad is not None}
        x = real.clone().requires_grad_(True)                       # R1: gradient of a gradient through D
        r1, = torch.autograd.grad(Dn(x, c).sum(), x, create_graph=True)
        for p_ in Dn.parameters():
            p_.grad = None
        r1.square().sum([1, 2, 3]).mean().backward()
        d_grads = {k: p_.grad.clone() for k, p_ in Dn.named_parameters() if p_.grad is not None}
        return img.detach(), logits.detach(), g_grads, d_grads

    try:
        img, logits, gg, dg = run(G, D)
    finally:
        host_networks.attach(networks, fused_callers=True)
    img_r, logits_r, gg_r, dg_r = run(G_ref, D_ref)
    _close(img, img_r, 2e-5, 'image'); _close(logits, logits_r, 2e-5, 'logits')
    assert set(gg) == set(gg_r) and set(dg) == set(dg_r)
    for k in gg_r:
        if float(gg_r[k].abs().max()) > 0:
            _close(gg[k], gg_r[k], 2e-4, 'G gradient ' + k)
    for k in dg_r:
        if float(dg_r[k].abs().max()) > 0:
            _close(dg[k], dg_r[k], 1e-3, 'R1 gradient ' + k)


def test_training_step_on_cpu_matches_the_reference_iteration(host_layer_on_cpu):
    """The training-step driver (ga-gan_b200/training/training_loop.py::TrainingStep: four loss phases with lazy-regularisation gains,
    two accumulation rounds, nan_to_num, Adam with the lazy-reg corrected betas, G_ema) on the host layer with stand-in kernels, against
    the same iteration of the live reference (training_loop.py:293-318, 459-512): parameter updates, last-phase gradients, G_ema, pl_mean.
    The GPU twin of this test is test_gpu_networks.py::test_full_training_iteration_matches_the_live_reference."""
    from oracle import live_ref
    if not live_ref.available() or not tests.util.HAVE_CHECKOUT:
        pytest.skip('the reference checkouts are absent')
    from tests.test_gpu_networks import _cpu_iteration
    from gagan_b200.training.training_loop import TrainingStep
    L = live_ref.load()
    networks = tests.util.reference_networks()
    res, zd = 16, 16
    kw_g = dict(z_dim=zd, c_dim=0, w_dim=zd, img_resolution=res, img_channels=3, mapping_kwargs=dict(num_layers=2),
                synthesis_kwargs=dict(channel_base=256, channel_max=16))
    kw_d = dict(c_dim=0, img_resolution=res, img_channels=3, channel_base=256, channel_max=16, epilogue_kwargs=dict(mbstd_group_size=2))
    torch.manual_seed(9)
    G_cpu, D_cpu = tests.util.quiet(L.networks.Generator, **kw_g).train(), tests.util.quiet(L.networks.Discriminator, **kw_d).train()
    with torch.no_grad():
        for p_ in list(G_cpu.parameters()) + list(D_cpu.parameters()):
            if float(p_.abs().max()) == 0:
                p_.copy_(torch.randn(p_.shape) * 0.1)
    G, D = tests.util.quiet(networks.Generator, **kw_g), tests.util.quiet(networks.Discriminator, **kw_d)
    G.load_state_dict(G_cpu.state_dict()); D.load_state_dict(D_cpu.state_dict())
    before = {('G.' + k): v.clone() for k, v in G_cpu.named_parameters()}
    before.update({('D.' + k): v.clone() for k, v in D_cpu.named_parameters()})
    lrate, gamma, batch, batch_gpu = 1e-3, 1.0, 4, 2
    real = torch.rand(batch, 3, res, res) * 2 - 1
    zs = torch.randn(4, batch, zd)
    L.conv2d_gradfix.enabled = True
    for net in (G_cpu, D_cpu):
        net.requires_grad_(False)
    with tests.util.patched_randn(21):
        G_ema_cpu, loss_cpu = _cpu_iteration(L, G_cpu, D_cpu, real, zs, batch_gpu, lrate, gamma)
    step = TrainingStep(G, D, batch_size=batch, batch_gpu=batch_gpu, device=torch.device('cpu'), lrate=lrate, r1_gamma=gamma, ema_kimg=10.0,
                        style_mixing_prob=0.0, pl_weight=2.0)
    with tests.util.patched_randn(21):
        step.run(real, zs)
    _close(step.loss.pl_mean, loss_cpu.pl_mean, 1e-4, 'pl_mean')
    after_ref = {('G.' + k): v for k, v in G_cpu.named_parameters()}
    after_ref.update({('D.' + k): v for k, v in D_cpu.named_parameters()})
    after = {('G.' + k): v for k, v in step.G.named_parameters()}
    after.update({('D.' + k): v for k, v in step.D.named_parameters()})
    total = off = 0
    for k, b0 in before.items():
        d_ref, d_got = after_ref[k].detach() - b0, after[k].detach() - b0
        assert float(d_ref.abs().max()) > 0, f'{k} was not updated by the reference iteration'
        bad = (d_got - d_ref).abs() > 0.02 * float(d_ref.abs().max())
        total += bad.numel(); off += int(bad.sum())
    assert off <= 2e-3 * total, f'{off} of {total} parameter updates differ from the reference iteration'
    for (k, a_), b_ in zip(step.G_ema.named_parameters(), G_ema_cpu.parameters()):
        _close(a_, b_, 1e-4, 'G_ema ' + k)
    assert {'Loss/G/loss', 'Loss/D/loss', 'Loss/G/reg', 'Loss/D/reg'} <= set(step.read_stats())


def _augment_pair(L, device, p=1.0, res=32):
    """The installed checkout's AugmentPipe (this build's operators) and the live reference's (impl='ref' ops), config 'bgc'."""
    import importlib
    from gagan_b200.training.training_loop import AUGPIPE_BGC
    augment = importlib.import_module('training.augment')
    mine, ref = augment.AugmentPipe(**AUGPIPE_BGC).to(device).train().requires_grad_(False), L.augment.AugmentPipe(**AUGPIPE_BGC).train().requires_grad_(False)
    mine.p.copy_(torch.as_tensor(p)); ref.p.copy_(torch.as_tensor(p))
    imgs = torch.rand(4, 3, res, res, generator=torch.Generator().manual_seed(6)) * 2 - 1
    return mine, ref, imgs


def test_ada_augment_pipe_on_the_host_layer_with_stand_in_kernels(host_layer_on_cpu):
    """The reference's ADA AugmentPipe (augment.py:121-531: geometric transforms through the separable 12-tap sym6 up / down-sampling
    with upfirdn2d, grid_sample, colour transforms) from the installed checkout, i.e. on this build's upfirdn2d module, against the
    live reference with the same random draws: images and the gradient w.r.t. the input images."""
    from oracle import live_ref
    if not live_ref.available() or not tests.util.HAVE_CHECKOUT:
        pytest.skip('the reference checkouts are absent')
    L = live_ref.load()
    L.grid_sample_gradfix.enabled = True
    mine, ref, imgs = _augment_pair(L, torch.device('cpu'))
    outs = []
    for pipe in (mine, ref):
        x = imgs.clone().requires_grad_(True)
        with tests.util.patched_rand(17):
            y = pipe(x)
        gx, = torch.autograd.grad((y * torch.linspace(-1, 1, y.numel()).reshape(y.shape)).sum(), x)
        outs.append((y.detach(), gx))
    _close(outs[0][0], outs[1][0], 2e-5, 'augmented images')
    _close(outs[0][1], outs[1][1], 2e-5, 'gradient w.r.t. the input images')


def _parametrizations():
    from tests.test_gpu_networks import PARAMETRIZATIONS
    return PARAMETRIZATIONS


@pytest.mark.parametrize('param', _parametrizations())
def test_domain_adaptation_parametrizations_on_the_host_layer_with_stand_in_kernels(host_layer_on_cpu, param):
    """CPU twin of test_gpu_networks.py::test_domain_adaptation_parametrizations_match_the_live_reference: every Affine+ / AffineLight+ /
    StyleSpace parameterization of the reference's generator through this build's modulated_conv2d, image and all parameter gradients."""
    from tests.test_gpu_networks import test_domain_adaptation_parametrizations_match_the_live_reference as body
    if not tests.util.HAVE_CHECKOUT:
        pytest.skip('baseline/_ref/DissimilarDomains is absent')
    body(torch.device('cpu'), param)


@pytest.fixture(params=['fused_forwards', 'reference_forwards'])
def nets_on_cpu(host_layer_on_cpu, request):
    """The golden generator / discriminator of tests/golden/networks.npz (outputs of the reference itself) built from the installed
    checkout on CPU tensors -- the same construction as the GPU fixture of test_gpu_networks.py."""
    if not tests.util.HAVE_CHECKOUT:
        pytest.skip('baseline/_ref/DissimilarDomains is absent')
    import gagan_b200.training.networks as mine
    from tests.util import load_golden, t
    networks = tests.util.reference_networks()
    mine.attach(networks, fused_callers=(request.param == 'fused_forwards'))
    g = load_golden('networks')
    cfg = {kv.split('=')[0]: int(kv.split('=')[1]) for kv in (str(m) for m in g['meta'])}
    G = networks.Generator(z_dim=cfg['z_dim'], c_dim=0, w_dim=cfg['w_dim'], img_resolution=cfg['res'], img_channels=3,
                           mapping_kwargs=dict(num_layers=cfg['num_layers']),
                           synthesis_kwargs=dict(channel_base=cfg['channel_base'], channel_max=cfg['channel_max']))
    D = networks.Discriminator(c_dim=0, img_resolution=cfg['res'], img_channels=3, channel_base=cfg['channel_base'],
                               channel_max=cfg['channel_max'], epilogue_kwargs=dict(mbstd_group_size=cfg['mbstd']))
    G.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('G.')}, strict=False)
    D.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('D.')}, strict=False)
    yield G, D, g, cfg
    mine.attach(networks, fused_callers=True)


def test_golden_forwards_on_the_host_layer_with_stand_in_kernels(nets_on_cpu):
    from tests import test_gpu_networks as gpu_tests
    gpu_tests.test_eval_forward(nets_on_cpu, torch.device('cpu'))
    gpu_tests.test_train_forward_random_noise(nets_on_cpu, torch.device('cpu'))


@pytest.mark.parametrize('phase', ['Gmain', 'Greg', 'Dmain', 'Dreg'])
def test_golden_loss_phases_on_the_host_layer_with_stand_in_kernels(nets_on_cpu, phase):
    """The four loss phases of the reference's own StyleGAN2Loss (loss.py:59-133, incl. the double-backward regularisers) on the host
    layer with stand-in kernels, against the golden parameter gradients the reference itself produced."""
    from tests import test_gpu_networks as gpu_tests
    gpu_tests.test_loss_phase_parameter_gradients(nets_on_cpu, torch.device('cpu'), phase)


def test_ga_population_fitness_on_the_host_layer_with_stand_in_kernels(host_layer_on_cpu):
    """CPU twin of test_gpu_networks.py::test_ga_population_fitness_eval_on_the_device (BASELINE configs[3] at toy size; eager path)."""
    if not tests.util.HAVE_CHECKOUT:
        pytest.skip('baseline/_ref/DissimilarDomains is absent')
    from tests import test_gpu_networks as gpu_tests
    gpu_tests.test_ga_population_fitness_eval_on_the_device(torch.device('cpu'))


# ----------------------------------------------------------------------------
# Mixed precision: float16 tensors at the operator boundaries, fp32 arithmetic inside.

HALF_ULP = 2.0 ** -11          # one rounding to float16, relative to the element; the bounds below are relative to max|reference|


def _half_like_reference(L):
    f = L.upfirdn2d.setup_filter([1, 3, 3, 1])
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 6, 12, 12, generator=g).half()
    w = (torch.randn(5, 6, 3, 3, generator=g) * 0.2)
    s = torch.randn(2, 6, generator=g) * 0.5 + 1
    b = (torch.randn(5, generator=g) * 0.3)
    return f, x, w, s, b


def test_float16_operators_round_once_at_their_boundary(host_layer_on_cpu):
    """Every public operator accepts float16 tensors (also channels_last ones), returns float16, sends float16 gradients back, and its
    result is the reference's fp32 result on the same (float16-valued) inputs rounded to float16 ONCE: within 2^-11 of max|y| (plus
    fp32 noise), where the reference's own fp16 evaluation (bias_act.py:127-157, upfirdn2d.py:179-219, conv2d_resample.py:59-154 on
    torch's CPU half kernels) rounds after every elementary step."""
    from oracle import live_ref
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent')
    L = live_ref.load()
    from torch_utils.ops import bias_act as BA, upfirdn2d as U, conv2d_resample as CR
    f, x, w, s, b = _half_like_reference(L)
    xc = x.to(memory_format=torch.channels_last)
    bound = HALF_ULP + 2e-5

    cases = {
        'bias_act lrelu clamp': (lambda t: BA.bias_act(t, b[:1].repeat(6).half(), act='lrelu', clamp=0.8),
                                 lambda t: L.bias_act.bias_act(t, b[:1].repeat(6).half().float(), act='lrelu', clamp=0.8, impl='ref')),
        'bias_act linear gain': (lambda t: BA.bias_act(t, None, act='linear', gain=0.5),
                                 lambda t: L.bias_act.bias_act(t, None, act='linear', gain=0.5, impl='ref')),
        'upfirdn2d up2': (lambda t: U.upsample2d(t, f), lambda t: L.upfirdn2d.upsample2d(t, f, impl='ref')),
        'upfirdn2d down2': (lambda t: U.downsample2d(t, f), lambda t: L.upfirdn2d.downsample2d(t, f, impl='ref')),
        'conv2d_resample plain': (lambda t: CR.conv2d_resample(t, w.half(), padding=1), lambda t: L.conv2d_resample.conv2d_resample(t, w.half().float(), padding=1)),
        'conv2d_resample up2': (lambda t: CR.conv2d_resample(x=t, w=w.half(), f=f, up=2, padding=1, flip_weight=False),
                                lambda t: L.conv2d_resample.conv2d_resample(x=t, w=w.half().float(), f=f, up=2, padding=1, flip_weight=False)),
        'conv2d_resample down2': (lambda t: CR.conv2d_resample(x=t, w=w.half(), f=f, down=2, padding=1),
                                  lambda t: L.conv2d_resample.conv2d_resample(x=t, w=w.half().float(), f=f, down=2, padding=1)),
        'conv2d': (lambda t: cg.conv2d(t, w.half(), padding=1), lambda t: torch.nn.functional.conv2d(t, w.half().float(), padding=1)),
        'conv_transpose2d': (lambda t: cg.conv_transpose2d(input=t, weight=w.half().transpose(0, 1), padding=1),
                             lambda t: torch.nn.functional.conv_transpose2d(t, w.half().float().transpose(0, 1), padding=1)),
    }
    for name, (mine, ref) in cases.items():
        for xin in (x, xc):
            t16 = xin.clone().requires_grad_(True)
            y = mine(t16)
            assert y.dtype == torch.float16, name
            t32 = x.float().requires_grad_(True)
            y32 = ref(t32)
            assert y.shape == y32.shape, name
            _close(y.float(), y32.detach(), bound, name)
            r = torch.randn(y32.shape, generator=torch.Generator().manual_seed(5)).half()
            gx, = torch.autograd.grad(y, t16, r)
            gx32, = torch.autograd.grad(y32, t32, r.float())
            assert gx.dtype == torch.float16, name
            _close(gx.float(), gx32, bound, name + ': dx')


@pytest.mark.parametrize('demodulate,up', [(True, 1), (True, 2), (False, 1)])
def test_float16_modulated_conv2d_against_the_live_reference(host_layer_on_cpu, demodulate, up):
    """modulated_conv2d on float16 activations (networks.py:591-668 with the fp16 pre-normalisation of :621-627): float16 in and out,
    and no further from the reference's fp32 result than ONE float16 rounding -- while the reference's own float16 evaluation (both
    formulations) sits several roundings away."""
    from oracle import live_ref
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent')
    L = live_ref.load()
    import training.networks as N
    f, x, w, s, b = _half_like_reference(L)
    kw = dict(up=up, padding=1, resample_filter=f, demodulate=demodulate, flip_weight=(up == 1))
    noise = torch.randn(12 * up, 12 * up, generator=torch.Generator().manual_seed(3)) * 0.1
    ws = [t.clone().requires_grad_(True) for t in (w, s)]
    y = N.modulated_conv2d(x=x, weight=ws[0], styles=ws[1], noise=noise, **kw)
    assert y.dtype == torch.float16
    wr = [t.clone().requires_grad_(True) for t in (w, s)]
    y32 = L.networks.modulated_conv2d(x=x.float(), weight=wr[0], styles=wr[1], noise=noise, fused_modconv=False, **kw)
    y16 = L.networks.modulated_conv2d(x=x, weight=w, styles=s, noise=noise, fused_modconv=False, **kw)
    e_mine, e_ref16 = tests.util.max_rel_err(y.float(), y32), tests.util.max_rel_err(y16.float(), y32)
    assert e_mine <= HALF_ULP + 2e-5, (e_mine, e_ref16)
    r = torch.randn(y32.shape, generator=torch.Generator().manual_seed(6)).half()
    g = torch.autograd.grad(y, ws, r)
    g32 = torch.autograd.grad(y32, wr, r.float())
    for nm, u, v in zip(('dweight', 'dstyles'), g, g32):
        assert u.dtype == torch.float32                      # parameters stay fp32 (networks.py:897: weight.to(x.dtype) inside the op)
        _close(u, v, 2e-3, f'fp16 modulated_conv2d {nm}')     # the incoming gradient is float16-valued; the rounding of y does not enter


@pytest.mark.parametrize('fused_callers', [True, False])
def test_mixed_precision_networks_on_the_host_layer(host_layer_on_cpu, fused_callers):
    """The reference's default configuration is mixed precision (`num_fp16_res=4`, `conv_clamp=256`, train.py:267-268,425-429): its own
    Generator / Discriminator built that way run on this build's host layer, and land CLOSER to the reference's fp32 evaluation
    (`force_fp32=True`) than the reference's own fp16 evaluation does: image, logits and every parameter gradient of a G loss."""
    from oracle import live_ref
    if not live_ref.available() or not tests.util.HAVE_CHECKOUT:
        pytest.skip('the reference checkouts are absent')
    L = live_ref.load()
    networks = tests.util.reference_networks()
    from gagan_b200.training import networks as host_networks
    host_networks.attach(networks, fused_callers=fused_callers)
    kw_g = dict(z_dim=16, c_dim=0, w_dim=16, img_resolution=32, img_channels=3, mapping_kwargs=dict(num_layers=2),
                synthesis_kwargs=dict(channel_base=512, channel_max=16, num_fp16_res=2, conv_clamp=256))
    kw_d = dict(c_dim=0, img_resolution=32, img_channels=3, channel_base=512, channel_max=16, num_fp16_res=2, conv_clamp=256,
                epilogue_kwargs=dict(mbstd_group_size=2))
    torch.manual_seed(4)
    G_ref, D_ref = tests.util.quiet(L.networks.Generator, **kw_g).train(), tests.util.quiet(L.networks.Discriminator, **kw_d).train()
    with torch.no_grad():
        for p_ in list(G_ref.parameters()) + list(D_ref.parameters()):
            if float(p_.abs().max()) == 0:
                p_.copy_(torch.randn(p_.shape) * 0.1)
    G, D = tests.util.quiet(networks.Generator, **kw_g).train(), tests.util.quiet(networks.Discriminator, **kw_d).train()
    G.load_state_dict(G_ref.state_dict()); D.load_state_dict(D_ref.state_dict())
    z = torch.randn(4, 16, generator=torch.Generator().manual_seed(2)); c = torch.zeros(4, 0)

    def run(Gn, Dn, **kw):
        for p_ in list(Gn.parameters()) + list(Dn.parameters()):
            p_.grad = None
        img = Gn(z, c, noise_mode='const', **kw)
        logits = Dn(img, c, **kw)
        torch.nn.functional.softplus(-logits).mean().backward()
        return img.detach(), logits.detach(), {k: p_.grad.clone() for k, p_ in Gn.named_parameters() if p_.grad is not None}

    try:
        mine = run(G, D)
    finally:
        host_networks.attach(networks, fused_callers=True)
    ref16, ref32 = run(G_ref, D_ref), run(G_ref, D_ref, force_fp32=True)
    err = tests.util.max_rel_err
    assert mine[0].dtype == torch.float32 and set(mine[2]) == set(ref32[2])
    for i, nm in ((0, 'image'), (1, 'logits')):
        e_mine, e_ref16 = err(mine[i], ref32[i]), err(ref16[i], ref32[i])
        assert e_mine <= 1.25 * e_ref16 + 1e-4, (nm, e_mine, e_ref16)        # measured: 0.9e-3 / 1.2e-3 against 1.8e-3 (image)
        assert err(mine[i], ref16[i]) <= 3 * e_ref16 + 1e-4, nm
    live = [k for k in ref32[2] if float(ref32[2][k].abs().max()) > 0]
    worst_mine, worst_ref16 = max(err(mine[2][k], ref32[2][k]) for k in live), max(err(ref16[2][k], ref32[2][k]) for k in live)
    assert worst_mine <= 1.25 * worst_ref16, (worst_mine, worst_ref16)        # measured: 0.08 / 0.09 against 0.105 (fp16 gradients, no loss scaling)


# ----------------------------------------------------------------------------
# The rosinality adapter: SimilarDomains/gan_models/StyleGAN2/model.py on this build's host layer.

def _rosinality_pair(fused_layers, size=32, style_dim=32):
    from oracle import live_ref
    if not live_ref.rosinality_available():
        pytest.skip('oracle/_ref/SimilarDomains is absent')
    ref = live_ref.load_rosinality()
    mine = tests.util.rosinality_model(fused_layers=fused_layers)
    assert mine is not ref and mine.ModulatedConv2d is not ref.ModulatedConv2d
    torch.manual_seed(7)
    G_ref, D_ref = ref.Generator(size, style_dim, 2, channel_multiplier=1), ref.Discriminator(size, channel_multiplier=1)
    with torch.no_grad():
        for p_ in list(G_ref.parameters()) + list(D_ref.parameters()):
            if float(p_.abs().max()) == 0:                          # biases, noise strengths: move them off zero
                p_.copy_(torch.randn(p_.shape) * 0.1)
    G, D = mine.Generator(size, style_dim, 2, channel_multiplier=1), mine.Discriminator(size, channel_multiplier=1)
    G.load_state_dict(G_ref.state_dict()); D.load_state_dict(D_ref.state_dict())
    # float64: these networks are 512 channels wide whatever the image size, and in fp32 a handful of the ~10^6 leaky-ReLU arguments
    # per layer change sign between any two correct implementations (measured: 2 of 524 288 in one layer, each moving the input
    # gradient by 2e-2 of its maximum; the reference's fp32 run has such flips against its own fp64 run as well).  The stand-in
    # kernels are dtype-generic, so the host logic is compared where it can be compared exactly.
    return ref, mine, G_ref.double(), D_ref.double(), G.double(), D.double()


@pytest.mark.parametrize('fused_layers', [True, False])
def test_rosinality_networks_on_the_host_layer_with_stand_in_kernels(host_layer_on_cpu, fused_layers):
    """GA-GAN's second StyleGAN2 code base (SimilarDomains/gan_models/StyleGAN2/model.py: ModulatedConv2d :176-275 with its grouped
    per-sample-weight convolutions, Blur / Upsample :30-88, StyledConv :305-341, ToRGB :342-362, ConvLayer / ResBlock / Discriminator
    :666-795) bound to this build by gagan_b200.install_rosinality, against the same file unmodified on its torch-native ops:
    image (W and S-code entry, fixed noise), logits, generator gradients, R1 gradients."""
    ref, mine, G_ref, D_ref, G, D = _rosinality_pair(fused_layers)
    z = torch.randn(4, 32, generator=torch.Generator().manual_seed(2)).double()
    real = torch.rand(4, 3, 32, 32, generator=torch.Generator().manual_seed(3)).double() * 2 - 1
    noises = [torch.randn(1, 1, 2 ** (2 + (i + 1) // 2), 2 ** (2 + (i + 1) // 2), generator=torch.Generator().manual_seed(10 + i)).double()
              for i in range(G.num_layers)]

    def run(Gn, Dn):
        for p_ in list(Gn.parameters()) + list(Dn.parameters()):
            p_.grad = None
        img, _ = Gn([z], noise=noises)
        logits = Dn(img)
        torch.nn.functional.softplus(-logits).mean().backward()
        g_grads = {k: p_.grad.clone() for k, p_ in Gn.named_parameters() if p_.grad is not None}
        x = real.clone().requires_grad_(True)
        r1, = torch.autograd.grad(Dn(x).sum(), x, create_graph=True)
        for p_ in Dn.parameters():
            p_.grad = None
        r1.square().sum([1, 2, 3]).mean().backward()
        d_grads = {k: p_.grad.clone() for k, p_ in Dn.named_parameters() if p_.grad is not None}
        with torch.no_grad():
            s_codes = Gn.get_s_code([Gn.style(z)], input_is_latent=True) if hasattr(Gn, 'get_s_code') else None
            img_s = Gn(s_codes, is_s_code=True, noise=noises)[0] if s_codes is not None else None
        return img.detach(), logits.detach(), g_grads, d_grads, img_s

    img, logits, gg, dg, img_s = run(G, D)
    img_r, logits_r, gg_r, dg_r, img_s_r = run(G_ref, D_ref)
    _close(img, img_r, 1e-11, 'image'); _close(logits, logits_r, 1e-11, 'logits')
    if img_s_r is not None:
        _close(img_s, img_s_r, 1e-11, 'image from S codes')
    assert set(gg) == set(gg_r) and set(dg) == set(dg_r)
    for k in gg_r:
        if float(gg_r[k].abs().max()) > 0:
            _close(gg[k], gg_r[k], 1e-9, 'G gradient ' + k)
    for k in dg_r:
        if float(dg_r[k].abs().max()) > 0:
            _close(dg[k], dg_r[k], 1e-8, 'R1 gradient ' + k)


def test_rosinality_adapter_pieces(host_layer_on_cpu):
    """The adapter's pieces one by one against the reference's torch-native ops (op/upfirdn2d_torch_native.py:10-59,
    op/fused_act_torch_native.py:23-37) and layers: pad conventions, gains, bias axes, random noise, and the refusal of layer
    shapes the adapter cannot express."""
    ref, mine, *_ = _rosinality_pair(True)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, 3, 9, 11, generator=g)
    k = ref.make_kernel([1, 3, 3, 1])
    for up, down, pad in ((1, 1, (2, 1)), (2, 1, (2, 1)), (1, 2, (1, 1)), (1, 1, (-1, 2))):
        _close(mine.upfirdn2d(x, k, up=up, down=down, pad=pad), ref.upfirdn2d(x, k, up=up, down=down, pad=pad), 1e-6, f'upfirdn2d {up} {down} {pad}')
    b = torch.randn(3, generator=g)
    _close(mine.fused_leaky_relu(x, b), ref.fused_leaky_relu(x, b), 1e-6, 'fused_leaky_relu NCHW')
    x3, b3 = torch.randn(2, 5, 7, generator=g), torch.randn(7, generator=g)
    _close(mine.fused_leaky_relu(x3, b3, 0.1, 1.5), ref.fused_leaky_relu(x3, b3, 0.1, 1.5), 1e-6, 'fused_leaky_relu rank 3')
    x2, b2 = torch.randn(4, 7, generator=g), torch.randn(7, generator=g)
    _close(mine.fused_leaky_relu(x2, b2), ref.fused_leaky_relu(x2, b2), 1e-6, 'fused_leaky_relu rank 2')
    for kw in (dict(upsample=True), dict(downsample=True), dict(), dict(demodulate=False)):
        for ks in (3, 1):
            torch.manual_seed(3)
            lr = ref.ModulatedConv2d(6, 4, ks, 8, **kw)
            lm = mine.ModulatedConv2d(6, 4, ks, 8, **kw)
            lm.load_state_dict(lr.state_dict())
            xi, st = torch.randn(2, 6, 8, 8, generator=g), torch.randn(2, 8, generator=g)
            _close(lm(xi, st), lr(xi, st), 2e-5, f'ModulatedConv2d {kw} k={ks}')
            sc = torch.randn(2, 6, generator=g)
            _close(lm(xi, sc, is_s_code=True), lr(xi, sc, is_s_code=True), 2e-5, f'ModulatedConv2d {kw} k={ks} from an S code')
    # random noise: StyledConv draws N(0,1) of the OUTPUT size when none is given (model.py:284-289)
    torch.manual_seed(5)
    sr, sm = ref.StyledConv(6, 4, 3, 8, upsample=True), mine.StyledConv(6, 4, 3, 8, upsample=True)
    with torch.no_grad():
        sr.noise.weight.fill_(0.3)
    sm.load_state_dict(sr.state_dict())
    xi, st = torch.randn(2, 6, 8, 8, generator=g), torch.randn(2, 8, generator=g)
    with tests.util.patched_randn(9):
        a = sm(xi, st)
    with tests.util.patched_randn(9):
        normal_ = torch.Tensor.normal_
        torch.Tensor.normal_ = lambda self, *a_, **k_: self.copy_(torch.randn(self.shape))     # `new_empty(...).normal_()` through the same generator
        try:
            b_ = sr(xi, st)
        finally:
            torch.Tensor.normal_ = normal_
    _close(a, b_, 2e-5, 'StyledConv with random noise')
    # a Blur whose pads are not the ones of the resampling convolution is refused, not approximated
    odd = mine.ModulatedConv2d(6, 4, 3, 8, upsample=True)
    odd.blur.pad = (2, 1)
    with pytest.raises(NotImplementedError):
        odd(xi, st)


# ----------------------------------------------------------------------------
@pytest.mark.parametrize('c_shape', ['N1HW', '11HW', 'HW'])
def test_fma_rows_route_against_the_live_reference(host_layer_on_cpu, monkeypatch, c_shape):
    """fma (fma.py:15-58) for the shape of its one call site (networks.py:648) takes the single-pass kernel route; values and
    gradients up to second order against the reference's fma, in fp64 through the stand-in kernel."""
    from oracle import live_ref
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent')
    L = live_ref.load()
    from torch_utils.ops import fma as F_
    monkeypatch.setattr(F_, '_on_device', lambda a: True)
    g = torch.Generator().manual_seed(3)
    N, C, H, W = 3, 5, 6, 7
    a = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    b = torch.randn(N, C, 1, 1, generator=g, dtype=torch.float64)
    c = torch.randn({'N1HW': (N, 1, H, W), '11HW': (1, 1, H, W), 'HW': (H, W)}[c_shape], generator=g, dtype=torch.float64)
    r = torch.randn(N, C, H, W, generator=g, dtype=torch.float64)
    assert F_._rows_shape(a, b, c) and not F_._rows_shape(a, b[:, :1], c) and not F_._rows_shape(a, b, c.reshape(-1))

    def run(fn):
        ts = [t.clone().requires_grad_(True) for t in (a, b, c)]
        y = fn(*ts)
        first = torch.autograd.grad((y * r).sum(), ts, create_graph=True)
        second = torch.autograd.grad(sum(t.square().sum() for t in first), ts, allow_unused=True)
        return [y] + list(first) + [t for t in second if t is not None]
    got, want = run(F_.fma), run(L.fma.fma)
    assert len(got) == len(want)
    for i, (u, v) in enumerate(zip(got, want)):
        assert u.shape == v.shape
        _close(u, v, 1e-12, f'fma {c_shape} output {i}')


# ----------------------------------------------------------------------------
@pytest.mark.parametrize('io,flip,scales', [(False, False, ''), (True, True, ''), (False, False, 'ab'), (True, False, 'a')])
def test_unaligned_rows_are_padded_for_the_tensor_core_path(fake_plugin, monkeypatch, io, flip, scales):
    """conv2d_gradfix.pad_unaligned_rows: inputs whose width is not a multiple of 4 reach the kernel zero-padded to the next multiple
    (the TMA row pitch), and nothing changes -- output, first- and second-order gradients -- because the output extent of the
    stride-1 primitive is a free parameter.  Narrow maps, thin layers and large kernels are left alone."""
    seen = []
    plain = type(fake_plugin).conv2d

    def spy(self, x, w, **kw):
        seen.append((int(x.shape[3]), tuple(w.shape)))
        return plain(self, x, w, **kw)
    monkeypatch.setattr(type(fake_plugin), 'conv2d', spy)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 16, 9, 67, generator=g, dtype=torch.float64)
    w = torch.randn(16, 16, 3, 3, generator=g, dtype=torch.float64) / 12
    a = torch.randn(2, 16, generator=g, dtype=torch.float64) * 0.5 + 1 if 'a' in scales else None
    b = torch.randn(2, 16, generator=g, dtype=torch.float64) * 0.5 + 1 if 'b' in scales else None
    r = torch.randn(2, 16, 9, 67, generator=g, dtype=torch.float64)
    results = {}
    for on in (False, True):
        monkeypatch.setattr(cg, 'pad_unaligned_rows', on)
        seen.clear()
        y, first, second, third, second_plain = _orders(lambda x_, w_, a_, b_: _mine(x_, w_, a_, b_, 1, io, flip), x, w, a, b, r, which=(0, 1))
        results[on] = [y] + list(first) + [t for t in second if t is not None]
        widths = {wd for wd, ws in seen}
        assert widths == ({68} if on else {67}), (on, widths)           # every launch (forward, dx, second order) saw aligned rows
    for u, v in zip(results[True], results[False]):
        _close(u, v, 1e-12, 'padded vs unpadded rows')
    # left alone: narrow maps, thin layers, 5x5 kernels
    monkeypatch.setattr(cg, 'pad_unaligned_rows', True)
    assert cg._tma_rows(torch.zeros(1, 16, 4, 33), (16, 16, 3, 3)).shape[3] == 33
    assert cg._tma_rows(torch.zeros(1, 3, 4, 67), (16, 3, 3, 3)).shape[3] == 67
    assert cg._tma_rows(torch.zeros(1, 16, 4, 67), (16, 16, 5, 5)).shape[3] == 67
    assert cg._tma_rows(torch.zeros(1, 16, 4, 67), (16, 16, 1, 1)).shape[3] == 68
    assert cg._tma_rows(torch.zeros(1, 16, 4, 68), (16, 16, 3, 3)).shape[3] == 68

