"""GPU parity tests proper: every op of the hot path, called through the product API (-> ctypes -> C ABI ->
sm_100a kernels), against (a) the committed goldens generated from the live reference and (b) the CPU oracle on
seeded inputs.  Tolerance: BASELINE.json north_star, max relative error 1e-3 (fp32, TF32 off)."""
import itertools
import numpy as np
import pytest
import torch

from tests.util import load_golden, t, assert_close, TOL
from oracle import ops_ref as R

pytestmark = pytest.mark.gpu


def _meta(g):
    return [str(m).split('|') for m in g['meta']]


@pytest.fixture(scope='module')
def ops(device):
    from torch_utils import custom_ops
    from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma
    try:
        from training import networks                 # the reference's module with this build's modulated_conv2d attached
    except ImportError:
        from gagan_b200.training import networks      # no checkout on this box: the function itself
    custom_ops.load_library()
    import types
    return types.SimpleNamespace(upfirdn2d=upfirdn2d, bias_act=bias_act, conv2d_resample=conv2d_resample,
                                 conv2d_gradfix=conv2d_gradfix, fma=fma, networks=networks, custom_ops=custom_ops)


# ------------------------------------------------------------------------------------------------ upfirdn2d
def test_upfirdn2d_golden(ops, device):
    g = load_golden('upfirdn2d')
    for name, up, down, pad, flip, gain in _meta(g):
        up = [int(v) for v in up.split(',')]; down = [int(v) for v in down.split(',')]; pad = [int(v) for v in pad.split(',')]
        f = t(g[name + '.f'], device) if name + '.f' in g else None
        x = t(g[name + '.x'], device, requires_grad=True)
        y = ops.upfirdn2d.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=bool(int(flip)), gain=float(gain))
        assert_close(y, g[name + '.y'], 1e-5, name)
        dx, = torch.autograd.grad(y, x, t(g[name + '.dy'], device))
        assert_close(dx, g[name + '.dx'], 1e-5, name + '.dx')
    x = t(g['wrap.x'], device)
    for fname, f in (('2d', t(g['wrap.f2d'], device)), ('1d', t(g['wrap.f1d'], device))):
        assert_close(ops.upfirdn2d.filter2d(x, f), g[f'wrap.filter2d.{fname}.y'], 1e-5)
        assert_close(ops.upfirdn2d.upsample2d(x, f), g[f'wrap.upsample2d.{fname}.y'], 1e-5)
        assert_close(ops.upfirdn2d.downsample2d(x, f), g[f'wrap.downsample2d.{fname}.y'], 1e-5)


@pytest.mark.parametrize('case', [
    # name, shape, up, down, pad  -- the StyleGAN2 call shapes that take the tiled 4x4 kernels (+ ragged edges)
    ('filt_p1', (2, 3, 129, 129), 1, 1, [1, 1, 1, 1]), ('filt_p2', (2, 3, 128, 128), 1, 1, [2, 2, 2, 2]),
    ('filt_ragged', (1, 2, 77, 203), 1, 1, [1, 1, 1, 1]), ('up2', (2, 3, 64, 64), 2, 1, [2, 1, 2, 1]),
    ('up2_oddpad', (1, 2, 50, 70), 2, 1, [1, 2, 3, 0]), ('down2', (2, 3, 128, 128), 1, 2, [1, 1, 1, 1]),
    ('down2_ragged', (1, 2, 101, 131), 1, 2, [1, 1, 1, 1]), ('filt_crop', (1, 2, 90, 140), 1, 1, [-3, 2, -1, 4]),
    ('up2_bwd_of_down2', (1, 2, 64, 64), 2, 1, [2, 2, 2, 2]),
    # odd widths (rows that are not 16-byte multiples) behind an up-sampling convolution: the scalar-load side of the marching kernel
    ('filt_p1_17', (3, 5, 17, 17), 1, 1, [1, 1, 1, 1]), ('filt_p1_33', (2, 7, 33, 33), 1, 1, [1, 1, 1, 1]),
    ('filt_p1_65', (2, 3, 65, 65), 1, 1, [1, 1, 1, 1]), ('filt_odd_ragged', (1, 3, 19, 45), 1, 1, [3, 0, 2, 1]),
    ('filt_p1_9', (2, 4, 9, 9), 1, 1, [1, 1, 1, 1]),
])
def test_upfirdn2d_tiled_vs_oracle(ops, device, case):
    name, shape, up, down, pad = case
    g = torch.Generator().manual_seed(hash(name) % 1000)
    f = R.setup_filter([1, 3, 3, 1])
    for flip, gain in ((False, 4.0), (True, 1.0)):
        x = torch.randn(*shape, generator=g)
        want = R.upfirdn2d(x, f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
        xg = x.to(device).requires_grad_(True)
        got = ops.upfirdn2d.upfirdn2d(xg, f.to(device), up=up, down=down, padding=pad, flip_filter=flip, gain=gain)
        assert got.shape == want.shape
        assert_close(got, want, 1e-5, name)
        dy = torch.randn(want.shape, generator=g)
        bw = R.upfirdn2d_backward_args(x.shape, want.shape, f, up, down, pad)
        want_dx = R.upfirdn2d(dy, f, flip_filter=(not flip), gain=gain, **bw)
        got_dx, = torch.autograd.grad(got, xg, dy.to(device))
        assert_close(got_dx, want_dx, 1e-5, name + '.dx')


def test_upfirdn2d_double_backward(ops, device):
    g = torch.Generator().manual_seed(5)
    f = R.setup_filter([1, 3, 3, 1]).to(device)
    x = torch.randn(1, 2, 70, 66, generator=g).to(device).requires_grad_(True)
    y = ops.upfirdn2d.upsample2d(x, f)
    dy = torch.randn(y.shape, generator=g).to(device).requires_grad_(True)
    dx, = torch.autograd.grad(y, x, dy, create_graph=True)
    v = torch.randn(dx.shape, generator=g).to(device)
    ddy, = torch.autograd.grad(dx, dy, v)          # linear op: d(dx)/d(dy) applied to v == forward(v)
    assert_close(ddy, ops.upfirdn2d.upsample2d(v, f), 1e-5)


@pytest.mark.parametrize('case', [(2, 5, 16, 16, [2, 2, 2, 2]), (1, 3, 33, 21, [2, 2, 2, 2]), (2, 4, 64, 300, [1, 1, 1, 1]), (1, 2, 7, 9, [2, 1, 2, 1])])
def test_fir_phase_major_fused_layouts(ops, device, case):
    # unit-rate 4x4 FIR fused with space-to-depth (fir_to_pm) / depth-to-space (fir_from_pm), forward and gradient,
    # against the oracle's upfirdn2d composed with the host re-layout helpers
    N, C, H, W, pad = case
    cr = ops.conv2d_resample
    g = torch.Generator().manual_seed(H * 100 + W)
    f = R.setup_filter([1, 3, 3, 1])
    x = torch.randn(N, C, H, W, generator=g)
    oh, ow = H + pad[2] + pad[3] - 3, W + pad[0] + pad[1] - 3
    ys, xs = (oh + 1) // 2 + 1, ((ow + 1) // 2 + 4) // 4 * 4
    xc = x.clone().requires_grad_(True)
    want = cr.space_to_depth(R.upfirdn2d(xc, f, padding=pad, gain=2.0), ys, xs)
    dy = torch.randn(want.shape, generator=g)
    wdx, = torch.autograd.grad(want, xc, dy)
    xg = x.to(device).requires_grad_(True)
    got = ops.upfirdn2d.fir_to_pm(xg, f.to(device), pad, False, 2.0, ys, xs)
    assert_close(got, want, 1e-6, 'fir_to_pm')
    gdx, = torch.autograd.grad(got, xg, dy.to(device))
    assert_close(gdx, wdx, 1e-6, 'fir_to_pm.dx')
    # the other direction: a phase-major tensor with a valid extent smaller than its planes
    z = torch.randn(N, 4 * C, ys, xs, generator=g)
    vh, vw = 2 * ys - 1, 2 * xs - 3
    zc = z.clone().requires_grad_(True)
    want2 = R.upfirdn2d(cr.depth_to_space(zc)[:, :, :vh, :vw], f, padding=pad, flip_filter=True, gain=4.0)
    dy2 = torch.randn(want2.shape, generator=g)
    wdz, = torch.autograd.grad(want2, zc, dy2)
    zg = z.to(device).requires_grad_(True)
    got2 = ops.upfirdn2d.fir_from_pm(zg, f.to(device), pad, True, 4.0, (vh, vw))
    assert_close(got2, want2, 1e-6, 'fir_from_pm')
    gdz, = torch.autograd.grad(got2, zg, dy2.to(device))
    assert_close(gdz, wdz, 1e-6, 'fir_from_pm.dz')


def test_upfirdn2d_errors(ops, device):
    x = torch.randn(1, 1, 2, 2, device=device)
    f = R.setup_filter([1, 3, 3, 1]).to(device)
    with pytest.raises(RuntimeError, match='at least 1x1'):
        ops.upfirdn2d.upfirdn2d(x, f, padding=0)          # 2 - 4 + 1 < 1
    with pytest.raises(RuntimeError, match='fp32'):
        ops.upfirdn2d.upfirdn2d(x.double(), f, padding=1)          # float64 is not served (float16 is: the mixed-precision tests below)


# ------------------------------------------------------------------------------------------------ bias_act
def test_bias_act_golden_all_activations_first_and_second_order(ops, device):
    g = load_golden('bias_act')
    for name, act, dim, alpha, gain, clamp in _meta(g):
        kw = dict(dim=int(dim), act=act, alpha=None if alpha == 'None' else float(alpha),
                  gain=None if gain == 'None' else float(gain), clamp=None if clamp == 'None' else float(clamp))
        x = t(g[name + '.x'], device, requires_grad=True)
        b = t(g[name + '.b'], device, requires_grad=True) if name + '.b' in g else None
        y = ops.bias_act.bias_act(x, b, **kw)
        assert_close(y, g[name + '.y'], 1e-5, name)
        dy = t(g[name + '.dy'], device, requires_grad=True)
        grads = torch.autograd.grad(y, [x] + ([b] if b is not None else []), dy, create_graph=True)
        assert_close(grads[0], g[name + '.dx'], 1e-4, name + '.dx')
        if b is not None:
            assert_close(grads[1], g[name + '.db'], 1e-4, name + '.db')
        gg = torch.autograd.grad(grads[0], [dy, x], t(g[name + '.d_dx'], device), allow_unused=True)
        assert_close(gg[0], g[name + '.gg_dy'], 1e-4, name + '.gg_dy')
        ggx = gg[1] if gg[1] is not None else torch.zeros_like(x)
        assert_close(ggx, g[name + '.gg_x'], 2e-4, name + '.gg_x')


@pytest.mark.parametrize('shape,dim', [((4, 512), 1), ((2, 32, 64, 64), 1), ((3, 16, 4, 4), 1), ((2, 5, 7, 9), 1),
                                       ((1, 8, 33, 128), 1), ((2, 512, 16, 16), 1)])
def test_bias_act_lrelu_shapes_vs_oracle(ops, device, shape, dim):
    g = torch.Generator().manual_seed(len(shape) * 7 + shape[-1])
    for clamp, gain in ((None, None), (0.8, 1.0)):
        x = torch.randn(*shape, generator=g)
        b = torch.randn(shape[dim], generator=g)
        dy = torch.randn(*shape, generator=g)
        xc, bc = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
        want = R.bias_act(xc, bc, dim=dim, act='lrelu', gain=gain, clamp=clamp)
        wdx, wdb = torch.autograd.grad(want, [xc, bc], dy)
        xg, bg = x.to(device).requires_grad_(True), b.to(device).requires_grad_(True)
        got = ops.bias_act.bias_act(xg, bg, dim=dim, act='lrelu', gain=gain, clamp=clamp)
        assert_close(got, want, 1e-6)
        gdx, gdb = torch.autograd.grad(got, [xg, bg], dy.to(device))
        assert_close(gdx, wdx, 1e-6)
        assert_close(gdb, wdb, 1e-4)      # fused reduction: fp32 atomics, order differs


def test_bias_act_r1_style_double_backward_through_bias(ops, device):
    # grad-of-grad w.r.t. the bias parameter, as Dreg needs (loss.py:141-152)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 8, 16, 16, generator=g); b = torch.randn(8, generator=g)

    def penalty(bias_act_fn, x, b):
        x = x.requires_grad_(True)
        y = bias_act_fn(x * x, b, act='lrelu')
        gx, = torch.autograd.grad(y.square().sum(), x, create_graph=True)
        return gx.square().sum()

    bc = b.clone().requires_grad_(True)
    want, = torch.autograd.grad(penalty(R.bias_act, x.clone(), bc), bc)
    bg = b.to(device).requires_grad_(True)
    got, = torch.autograd.grad(penalty(ops.bias_act.bias_act, x.to(device), bg), bg)
    assert_close(got, want, 1e-4)


def test_bias_act_empty_and_errors(ops, device):
    assert ops.bias_act.bias_act(torch.empty(0, 4, device=device), torch.zeros(4, device=device), act='lrelu').shape == (0, 4)
    with pytest.raises(RuntimeError, match='wrong number of elements'):
        ops.bias_act.bias_act(torch.zeros(2, 3, device=device), torch.zeros(4, device=device), act='lrelu')


# ------------------------------------------------------------------------------------------------ conv
def test_conv2d_resample_golden(ops, device):
    g = load_golden('conv2d_resample')
    f = t(g['f'], device)
    for name, up, down, pad, flipw, groups in _meta(g):
        x = t(g[name + '.x'], device, requires_grad=True); w = t(g[name + '.w'], device, requires_grad=True)
        y = ops.conv2d_resample.conv2d_resample(x, w, f=f, up=int(up), down=int(down), padding=[int(v) for v in pad.split(',')],
                                                groups=int(groups), flip_weight=bool(int(flipw)))
        assert_close(y, g[name + '.y'], TOL, name)
        dx, dw = torch.autograd.grad(y, [x, w], t(g[name + '.dy'], device))
        assert_close(dx, g[name + '.dx'], TOL, name + '.dx')
        assert_close(dw, g[name + '.dw'], TOL, name + '.dw')


def _conv_oracle(x, w, stride, pad, transposed, outpad=0):
    if transposed:
        return torch.nn.functional.conv_transpose2d(x, w, stride=stride, padding=pad, output_padding=outpad)
    return torch.nn.functional.conv2d(x, w, stride=stride, padding=pad)


@pytest.mark.parametrize('prec', ['simt', 'auto'])
@pytest.mark.parametrize('case', [
    # N, I, O, H, W, k, stride, pad, transposed
    (2, 3, 32, 32, 32, 1, 1, 0, False), (2, 32, 3, 32, 32, 1, 1, 0, False), (3, 16, 16, 4, 4, 3, 1, 1, False),
    (2, 32, 32, 16, 16, 3, 1, 1, False), (1, 64, 64, 32, 32, 3, 1, 1, False), (2, 128, 64, 16, 16, 3, 1, 1, False),
    (1, 32, 48, 24, 40, 3, 1, 1, False), (2, 33, 17, 9, 11, 3, 1, 1, False), (2, 16, 32, 17, 17, 3, 2, 0, False),
    (2, 32, 16, 8, 8, 3, 2, 0, True), (1, 8, 8, 5, 7, 3, 1, 1, True), (1, 513, 16, 4, 4, 3, 1, 1, False),
    (2, 32, 32, 64, 64, 3, 1, 1, False), (1, 16, 16, 16, 16, 1, 1, 0, False),
    # 2x2 kernels (the phase-major form of the stride-2 layers), ragged channel counts, many tiles per persistent CTA
    (2, 64, 32, 20, 20, 2, 1, 0, False), (2, 32, 64, 16, 16, 2, 1, 1, False), (1, 512, 130, 16, 16, 3, 1, 1, False),
    (3, 48, 48, 40, 40, 3, 1, 1, True), (5, 32, 32, 72, 72, 3, 1, 1, False), (2, 256, 48, 12, 12, 2, 1, 1, True),
    # weight-gradient tilings: 2 kx groups (64 grad channels), 1 kx per group (>=128), ragged widths, 2x2 kernels
    (2, 64, 200, 40, 40, 3, 1, 1, False), (1, 160, 96, 24, 20, 2, 1, 0, False), (2, 96, 40, 33, 17, 3, 1, 1, True),
    (1, 20, 24, 70, 50, 3, 1, 1, False),
])
def test_conv2d_fwd_dgrad_wgrad_vs_oracle(ops, device, case, prec):
    N, I, O, H, W, k, stride, pad, transposed = case
    g = torch.Generator().manual_seed(N * 1000 + I * 10 + k)
    x = torch.randn(N, I, H, W, generator=g)
    wshape = (I, O, k, k) if transposed else (O, I, k, k)
    w = torch.randn(*wshape, generator=g) / np.sqrt(I * k * k)
    xc, wc = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    want = _conv_oracle(xc, wc, stride, pad, transposed)
    dy = torch.randn(want.shape, generator=g)
    wdx, wdw = torch.autograd.grad(want, [xc, wc], dy)
    ops.custom_ops.conv_precision = ops.custom_ops.PREC_FP32_SIMT if prec == 'simt' else ops.custom_ops.PREC_AUTO
    try:
        xg, wg = x.to(device).requires_grad_(True), w.to(device).requires_grad_(True)
        fn = ops.conv2d_gradfix.conv_transpose2d if transposed else ops.conv2d_gradfix.conv2d
        got = fn(xg, wg, stride=stride, padding=pad)
        assert got.shape == want.shape
        tol = 2e-5                                   # both the FFMA and the 3xTF32 tcgen05 path are fp32-faithful
        assert_close(got, want, tol, 'fwd')
        gdx, gdw = torch.autograd.grad(got, [xg, wg], dy.to(device))
        assert_close(gdx, wdx, tol, 'dgrad')
        assert_close(gdw, wdw, tol * 5, 'wgrad')
    finally:
        ops.custom_ops.conv_precision = ops.custom_ops.PREC_AUTO


def test_conv2d_tc_is_fp32_faithful(ops, device):
    # The tensor core truncates its fp32 accumulator toward zero; un-chunked, that is a -5e-5 systematic shrink at
    # 512 channels (tools/tc_rounding.py).  The kernel's per-K-block register accumulation must remove it.
    g = torch.Generator().manual_seed(5)
    plugin = ops.custom_ops.get_plugin('conv2d_plugin')
    x = torch.randn(2, 512, 32, 32, generator=g).abs().to(device)
    w = (torch.randn(64, 512, 3, 3, generator=g).abs() / 68).to(device)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    y = plugin.conv2d(x, w, padding=(1, 1), prec=ops.custom_ops.PREC_TF32X3).double()
    assert plugin.last_conv_prec == 3
    bias = float((y - ref).mean() / ref.abs().mean())
    rms = float((y - ref).square().mean().sqrt() / ref.abs().mean())
    print(f'tf32x3 512ch: mean signed rel err {bias:+.2e}, rms {rms:.2e}')
    assert abs(bias) < 5e-7 and rms < 2e-6            # un-chunked: -5.2e-5; chunked, uncompensated: -1.4e-6


def test_conv2d_wgrad_tc_is_fp32_faithful_and_scaled(ops, device):
    # long reductions (64k pixels) with same-sign products: the per-strip register accumulation must keep the tensor
    # core's truncation bias out; per-sample scales on both operands (the modulated_conv2d factors)
    g = torch.Generator().manual_seed(8)
    plugin = ops.custom_ops.get_plugin('conv2d_plugin')
    x = torch.randn(4, 32, 128, 128, generator=g).abs(); dy = torch.randn(4, 64, 128, 128, generator=g).abs()
    a = torch.rand(4, 32, generator=g) + 0.5; b = torch.rand(4, 64, generator=g) + 0.5
    xd = (x * a[:, :, None, None]).double().requires_grad_(False); dyd = (dy * b[:, :, None, None]).double()
    w = torch.zeros(64, 32, 3, 3, dtype=torch.float64, requires_grad=True)
    ref, = torch.autograd.grad(torch.nn.functional.conv2d(xd, w, padding=1), w, dyd)
    got = plugin.conv2d_wgrad(x.to(device), dy.to(device), (3, 3), padding=(1, 1), a_scale=a.to(device), b_scale=b.to(device),
                              prec=ops.custom_ops.PREC_TF32X3).double().cpu()
    assert plugin.last_wgrad_prec == 3
    bias = float(((got - ref) / ref).mean()); worst = float(((got - ref) / ref).abs().max())
    print(f'wgrad tf32x3 64k pixels: mean signed rel err {bias:+.2e}, worst {worst:.2e}')
    assert abs(bias) < 1e-6 and worst < 1e-5      # un-chunked this would be ~4e-4, chunked but uncompensated -2.7e-6


@pytest.mark.parametrize('case', [(2, 32, 3, 64, 64), (2, 3, 32, 64, 64), (3, 512, 3, 8, 8), (1, 3, 64, 32, 48), (2, 4, 4, 16, 16)])
def test_conv1x1_thin_channels_fwd_dgrad_wgrad_scaled(ops, device, case):
    # ToRGB (C -> 3, modulated) / fromRGB (3 -> C): the HBM-streaming 1x1 kernels, with per-sample scales on both sides
    N, I, O, H, W = case
    g = torch.Generator().manual_seed(I * 7 + O)
    plugin = ops.custom_ops.get_plugin('conv2d_plugin')
    x = torch.randn(N, I, H, W, generator=g); w = torch.randn(O, I, 1, 1, generator=g)
    a = torch.randn(N, I, generator=g); b = torch.randn(N, O, generator=g); dy = torch.randn(N, O, H, W, generator=g)
    want = torch.nn.functional.conv2d(x * a[:, :, None, None], w) * b[:, :, None, None]
    got = plugin.conv2d(x.to(device), w.to(device), in_scale=a.to(device), out_scale=b.to(device))
    assert plugin.last_conv_prec == 0
    assert_close(got, want, 2e-6, 'fwd')
    wt = w.transpose(0, 1).contiguous()                                     # [I,O,1,1] read through the transposed layout
    got_t = plugin.conv2d(x.to(device), wt.to(device), transposed=True, in_scale=a.to(device), out_scale=b.to(device))
    assert_close(got_t, want, 2e-6, 'fwd(transposed layout)')
    wdw = torch.einsum('nohw,nihw->oi', dy * b[:, :, None, None], x * a[:, :, None, None])[:, :, None, None]
    gdw = plugin.conv2d_wgrad(x.to(device), dy.to(device), (1, 1), a_scale=a.to(device), b_scale=b.to(device))
    assert plugin.last_wgrad_prec == 0
    assert_close(gdw, wdw, 2e-5, 'wgrad')
    gdw_t = plugin.conv2d_wgrad(x.to(device), dy.to(device), (1, 1), a_scale=a.to(device), b_scale=b.to(device), out_layout=1)
    assert_close(gdw_t, wdw.transpose(0, 1), 2e-5, 'wgrad [A,B] layout')


def test_conv2d_tc_zero_block_skipping(ops, device):
    # structurally zero (K-block, tap) weight blocks are skipped; the result must not change
    g = torch.Generator().manual_seed(6)
    plugin = ops.custom_ops.get_plugin('conv2d_plugin')
    x = torch.randn(2, 64, 24, 24, generator=g)
    w = torch.randn(160, 64, 2, 2, generator=g) / 16
    w[:, 16:32] = 0                      # a dead K-block
    w[:, 32:48, 1, :] = 0                # dead taps in one K-block
    w[128:, :, :, 1] = 0                 # dead taps in one n-tile
    w[:128, 48:64] = 0                   # a K-block that is dead in the first n-tile only
    want = torch.nn.functional.conv2d(x, w, padding=1)
    got = plugin.conv2d(x.to(device), w.to(device), padding=(1, 1), prec=ops.custom_ops.PREC_TF32X3)
    assert_close(got, want, 1e-5)
    w0 = torch.zeros(32, 64, 3, 3)
    got = plugin.conv2d(x.to(device), w0.to(device), padding=(1, 1), prec=ops.custom_ops.PREC_TF32X3)
    assert float(got.abs().max()) == 0.0


def test_conv2d_scales_fused(ops, device):
    # in_scale / out_scale of gg_conv2d_f32 == the multiply passes of modulated_conv2d (networks.py:642,648-651)
    g = torch.Generator().manual_seed(3)
    plugin = ops.custom_ops.get_plugin('conv2d_plugin')
    for (N, I, O, H, k) in [(2, 32, 32, 16, 3), (3, 8, 5, 6, 3), (2, 64, 3, 32, 1)]:
        x = torch.randn(N, I, H, H, generator=g); w = torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)
        a = torch.randn(N, I, generator=g); b = torch.randn(N, O, generator=g)
        want = torch.nn.functional.conv2d(x * a[:, :, None, None], w, padding=k // 2) * b[:, :, None, None]
        got = plugin.conv2d(x.to(device), w.to(device), padding=(k // 2, k // 2), in_scale=a.to(device), out_scale=b.to(device))
        assert_close(got, want, 1e-4)


@pytest.mark.parametrize('k,pad', [(3, 1), (1, 0), (2, 1)])
def test_scaled_conv_first_and_second_order_gradients(ops, device, k, pad):
    # y = out_scale * conv(in_scale * x, w) with the scales inside the kernel: gradients w.r.t. x, w and both scales, and
    # the path-length-style second-order gradient (gradient of a squared first-order gradient), against torch on the CPU
    g = torch.Generator().manual_seed(20 + k)
    N, I, O, H = 2, 24, 20, 12
    x = torch.randn(N, I, H, H, generator=g); w = torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)
    a = torch.randn(N, I, generator=g) + 1.5; b = torch.rand(N, O, generator=g) + 0.5
    r = torch.randn(N, O, H + 2 * pad - k + 1, H + 2 * pad - k + 1, generator=g)

    def run(conv, x, w, a, b, r, full=False):
        ts = [t_.clone().requires_grad_(True) for t_ in (x, w, a, b)]
        y = conv(*ts)
        first = torch.autograd.grad((y * r).sum(), ts, create_graph=True)
        pen = first[2].square().sum() + first[0].square().mean()          # depends on d/da and d/dx
        if full:                                                          # ... and on d/db and d/dw: every second-order node
            pen = pen + 0.5 * first[3].square().sum() + 3.0 * first[1].square().sum()
        second = torch.autograd.grad(pen, ts, allow_unused=True)
        return y, first, second

    ref = run(lambda x_, w_, a_, b_: torch.nn.functional.conv2d(x_ * a_[:, :, None, None], w_, padding=pad) * b_[:, :, None, None],
              x, w, a, b, r)
    dev = [t_.to(device) for t_ in (x, w, a, b, r)]
    got = run(lambda x_, w_, a_, b_: ops.conv2d_gradfix.conv2d_s1(x_, w_, padding=(pad, pad), in_scale=a_, out_scale=b_), *dev)
    assert_close(got[0], ref[0], 2e-5, 'y')
    for name, g1, r1 in zip('xwab', got[1], ref[1]):
        assert_close(g1, r1, 5e-5, 'd' + name)
    for name, g2, r2 in zip('xwab', got[2], ref[2]):
        assert_close(g2, r2, 2e-4, 'second-order d' + name)
    ref_conv = lambda x_, w_, a_, b_: torch.nn.functional.conv2d(x_ * a_[:, :, None, None], w_, padding=pad) * b_[:, :, None, None]
    my_conv = lambda x_, w_, a_, b_: ops.conv2d_gradfix.conv2d_s1(x_, w_, padding=(pad, pad), in_scale=a_, out_scale=b_)
    ref_full, got_full = run(ref_conv, x, w, a, b, r, full=True), run(my_conv, *dev, full=True)
    for name, g2, r2 in zip('xwab', got_full[2], ref_full[2]):
        assert_close(g2, r2, 2e-4, 'second-order (all four first-order gradients in the penalty) d' + name)
    # the round-1 formulation of the second-order path (broadcast multiplies around the unscaled primitives) gives the same numbers
    ops.conv2d_gradfix.closed_scaled_backward = False
    try:
        old_full = run(my_conv, *dev, full=True)
    finally:
        ops.conv2d_gradfix.closed_scaled_backward = True
    for name, g2, r2 in zip('xwab', old_full[2], ref_full[2]):
        assert_close(g2, r2, 2e-4, 'second-order, spelled-out formulation, d' + name)
    # only one scale present (ToRGB: styles without demodulation)
    for has_a, has_b in ((True, False), (False, True)):
        rc = lambda x_, w_, a_, b_: (torch.nn.functional.conv2d(x_ * a_[:, :, None, None] if has_a else x_, w_, padding=pad)
                                     * (b_[:, :, None, None] if has_b else 1.0)) + 0.0 * (a_.sum() + b_.sum())
        mc = lambda x_, w_, a_, b_: ops.conv2d_gradfix.conv2d_s1(x_, w_, padding=(pad, pad), in_scale=(a_ if has_a else None),
                                                                 out_scale=(b_ if has_b else None)) + 0.0 * (a_.sum() + b_.sum())
        r1, g1 = run(rc, x, w, a, b, r, full=True), run(mc, *dev, full=True)
        for name, gg, rr, like in zip('xwab', g1[2], r1[2], (x, w, a, b)):
            gg = torch.zeros_like(like) if gg is None else gg           # (the absent scale: no second-order gradient on either side)
            rr = torch.zeros_like(like) if rr is None else rr
            assert_close(gg, rr, 2e-4, f'second-order with has_a={has_a} has_b={has_b} d' + name)
    # plain first-order backward (no create_graph) takes the fused kernels
    ts = [t_.clone().requires_grad_(True) for t_ in dev[:4]]
    y = ops.conv2d_gradfix.conv2d_s1(ts[0], ts[1], padding=(pad, pad), in_scale=ts[2], out_scale=ts[3])
    grads = torch.autograd.grad(y, ts, dev[4])
    for name, g1, r1 in zip('xwab', grads, ref[1]):
        assert_close(g1, r1, 5e-5, 'fused d' + name)


def test_scaled_conv_third_order_gradients_fall_back_to_autograd(ops, device):
    # the explicit second-order nodes of the scaled convolution (DxDa, DemodDot, RatioScale) hand higher orders to autograd over the
    # spelled-out expression: gradient of a gradient of a gradient against torch on the CPU
    g = torch.Generator().manual_seed(77)
    N, I, O, H, k, pad = 2, 6, 5, 7, 3, 1
    x = torch.randn(N, I, H, H, generator=g); w = torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)
    a = torch.randn(N, I, generator=g) + 1.5; b = torch.rand(N, O, generator=g) + 0.5
    r = torch.randn(N, O, H, H, generator=g)

    def third(conv, x, w, a, b, r):
        ts = [t_.clone().requires_grad_(True) for t_ in (x, w, a, b)]
        y = conv(*ts)
        first = torch.autograd.grad((y * r).sum(), ts, create_graph=True)
        pen = first[2].square().sum() + first[0].square().mean() + first[3].square().sum()
        second = torch.autograd.grad(pen, ts, create_graph=True)
        pen2 = sum(s_.square().sum() for s_ in second)
        return torch.autograd.grad(pen2, ts, allow_unused=True)

    ref = third(lambda x_, w_, a_, b_: torch.nn.functional.conv2d(x_ * a_[:, :, None, None], w_, padding=pad) * b_[:, :, None, None], x, w, a, b, r)
    got = third(lambda x_, w_, a_, b_: ops.conv2d_gradfix.conv2d_s1(x_, w_, padding=(pad, pad), in_scale=a_, out_scale=b_),
                *[t_.to(device) for t_ in (x, w, a, b, r)])
    for name, g3, r3 in zip('xwab', got, ref):
        assert_close(g3, r3, 1e-3, 'third-order d' + name)


def test_conv2d_double_backward_closure(ops, device):
    # R1-style: gradient of ||d y / d x||^2 w.r.t. the weight goes conv -> dgrad -> (wgrad of dgrad)
    g = torch.Generator().manual_seed(4)
    x = torch.randn(2, 8, 12, 12, generator=g); w = torch.randn(6, 8, 3, 3, generator=g) * 0.2

    def pen(conv, x, w):
        x = x.requires_grad_(True)
        y = conv(x, w)
        gx, = torch.autograd.grad((y * y).sum(), x, create_graph=True)
        return gx.square().sum()

    wc = w.clone().requires_grad_(True)
    want, = torch.autograd.grad(pen(lambda a, b: torch.nn.functional.conv2d(a, b, padding=1), x.clone(), wc), wc)
    wg = w.to(device).requires_grad_(True)
    got, = torch.autograd.grad(pen(lambda a, b: ops.conv2d_gradfix.conv2d(a, b, padding=1), x.to(device), wg), wg)
    assert_close(got, want, 1e-4)
    with ops.conv2d_gradfix.no_weight_gradients():
        xg = x.to(device).requires_grad_(True)
        y = ops.conv2d_gradfix.conv2d(xg, wg, padding=1)
        gx, gw = torch.autograd.grad(y.sum(), [xg, wg], allow_unused=True)
        assert gw is None and gx is not None


# ------------------------------------------------------------------------------------------------ modconv
def test_modulated_conv2d_golden(ops, device):
    g = load_golden('modconv')
    f = t(g['f'], device)
    for name, k, up, demod, noise_kind, fused in _meta(g):
        x = t(g[name + '.x'], device, requires_grad=True); w = t(g[name + '.w'], device, requires_grad=True)
        s = t(g[name + '.s'], device, requires_grad=True)
        noise = t(g[name + '.noise'], device, requires_grad=True) if name + '.noise' in g else None
        y = ops.networks.modulated_conv2d(x=x, weight=w, styles=s, noise=noise, up=int(up), padding=int(k) // 2,
                                          resample_filter=f, demodulate=bool(int(demod)), flip_weight=(int(up) == 1),
                                          fused_modconv=bool(int(fused)))
        assert_close(y, g[name + '.y'], TOL, name)
        grads = torch.autograd.grad(y, [x, w, s] + ([noise] if noise is not None else []), t(g[name + '.dy'], device))
        for gname, got in zip(['dx', 'dw', 'ds', 'dnoise'], grads):
            assert_close(got, g[f'{name}.{gname}'], TOL, f'{name}.{gname}')


@pytest.mark.parametrize('up,k,demod', [(1, 3, True), (2, 3, True), (1, 1, False)])
def test_modulated_conv2d_second_order_vs_oracle(ops, device, up, k, demod):
    # path-length-style regulariser through one modulated conv: the gradient of a squared first-order gradient w.r.t. the
    # styles, taken under no_weight_gradients() exactly like loss.py:98, must match the oracle (torch ops on the CPU)
    g = torch.Generator().manual_seed(40 + up + k)
    N, I, O, H = 2, 24, 16 if k == 3 else 3, 8
    x = torch.randn(N, I, H, H, generator=g); w = torch.randn(O, I, k, k, generator=g)
    s = torch.randn(N, I, generator=g) * 0.5 + 1.0
    noise = torch.randn(N, 1, H * up, H * up, generator=g) * 0.1 if demod else None
    f = R.setup_filter([1, 3, 3, 1])
    r = torch.randn(N, O, H * up, H * up, generator=g)

    def run(mc, ctx, x, w, s, noise, f, r):
        ts = [t_.clone().requires_grad_(True) for t_ in (x, w, s)]
        y = mc(x=ts[0], weight=ts[1], styles=ts[2], noise=noise, up=up, padding=k // 2, resample_filter=f, demodulate=demod,
               flip_weight=(up == 1), fused_modconv=False)
        with ctx():
            gs, gx = torch.autograd.grad((y * r).sum(), [ts[2], ts[0]], create_graph=True)
        pen = gs.square().sum() + gx.square().sum()
        return y, gs, torch.autograd.grad(pen, ts)

    import contextlib
    ref = run(R.modulated_conv2d, contextlib.nullcontext, x, w, s, noise, f, r)
    d = lambda t_: t_.to(device) if t_ is not None else None
    got = run(ops.networks.modulated_conv2d, ops.conv2d_gradfix.no_weight_gradients, d(x), d(w), d(s), d(noise), d(f), d(r))
    assert_close(got[0], ref[0], 2e-5, 'y')
    assert_close(got[1], ref[1], 5e-5, 'd styles')
    for name, a_, b_ in zip(['x', 'w', 'styles'], got[2], ref[2]):
        assert_close(a_, b_, 2e-4, 'second-order d' + name)


# ------------------------------------------------------------------------------------------------ full-size properties
@pytest.mark.parametrize('case', [
    # (N, I, O, R, up, down): config-f layer shapes at BASELINE.json's full size (1024^2 generator / discriminator)
    (2, 32, 32, 1024, 1, 1), (2, 64, 32, 512, 2, 1), (2, 32, 64, 1024, 1, 2), (4, 512, 512, 64, 1, 1), (4, 512, 512, 32, 2, 1),
    (2, 128, 256, 256, 1, 2),
])
def test_full_size_conv_layers_linearity_and_adjoints(ops, device, case):
    """Size-independent properties at full layer sizes (too large for the CPU oracle): the conv2d_resample layers are linear in
    x and in w, their data gradient is the adjoint map (<A x, y> == <x, A^T y>) and their weight gradient the adjoint in w
    (<A_w x, y> == <w, dW(x, y)>).  Inner products are accumulated in fp64."""
    N, I, O, R, up, down = case
    g = torch.Generator(device=device).manual_seed(R + I)
    f = R_setup = ops.upfirdn2d.setup_filter([1, 3, 3, 1]).to(device)
    x1 = torch.randn(N, I, R, R, device=device, generator=g); x2 = torch.randn(N, I, R, R, device=device, generator=g)
    w = torch.randn(O, I, 3, 3, device=device, generator=g) / np.sqrt(9 * I)
    w2 = torch.randn(O, I, 3, 3, device=device, generator=g) / np.sqrt(9 * I)

    def A(x, w_):
        return ops.conv2d_resample.conv2d_resample(x, w_, f=f, up=up, down=down, padding=1, flip_weight=(up == 1))

    def dot(a, b):
        return float((a.detach().double() * b.detach().double()).sum())

    xr = x1.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
    y = A(xr, wr)
    assert y.shape == (N, O, R * up // down, R * up // down)
    yy = torch.randn(y.shape, device=device, generator=g)
    dx, dw = torch.autograd.grad(y, [xr, wr], yy)
    lhs = dot(y, yy)
    assert abs(lhs - dot(x1, dx)) <= 1e-5 * max(abs(lhs), np.sqrt(y.numel())), 'data gradient is not the adjoint'
    assert abs(lhs - dot(w, dw)) <= 1e-5 * max(abs(lhs), np.sqrt(y.numel())), 'weight gradient is not the adjoint'
    with torch.no_grad():
        y1 = y.detach(); y2 = A(x2, w)
        lin = A(0.5 * x1 - 2.0 * x2, w)
        assert_close(lin, 0.5 * y1 - 2.0 * y2, 2e-5, 'linearity in x')
        linw = A(x1, 0.25 * w + 3.0 * w2)
        assert_close(linw, 0.25 * y1 + 3.0 * A(x1, w2), 2e-5, 'linearity in w')


def test_full_size_elementwise_properties(ops, device):
    """bias_act / upfirdn2d at the largest config-f tensors: relu-type idempotence, exact clamp, FIR linearity and the
    adjoint pair upsample2d / its gradient (upfirdn2d.py:264-283)."""
    g = torch.Generator(device=device).manual_seed(5)
    x = torch.randn(4, 32, 1024, 1024, device=device, generator=g); b = torch.randn(32, device=device, generator=g)
    y = ops.bias_act.bias_act(x, b, act='relu', gain=1.0)
    assert torch.equal(ops.bias_act.bias_act(y, None, act='relu', gain=1.0), y)                       # idempotent
    assert float(ops.bias_act.bias_act(x, b, act='lrelu', clamp=0.5).abs().max()) <= 0.5                 # clamp is exact
    lin = ops.bias_act.bias_act(x, b, act='linear', gain=2.0)
    assert_close(lin, 2.0 * (x + b[None, :, None, None]), 1e-6, 'linear')
    f = ops.upfirdn2d.setup_filter([1, 3, 3, 1]).to(device)
    img = torch.randn(4, 3, 512, 512, device=device, generator=g).requires_grad_(True)
    up = ops.upfirdn2d.upsample2d(img, f)
    assert up.shape == (4, 3, 1024, 1024)
    yy = torch.randn(up.shape, device=device, generator=g)
    dimg, = torch.autograd.grad(up, img, yy)
    lhs = float((up.double() * yy.double()).sum())
    assert abs(lhs - float((img.double() * dimg.double()).sum())) <= 1e-6 * max(abs(lhs), 1e3)
    # a constant image stays constant under the normalised up / down filters (DC gain 1)
    ones = torch.ones(2, 8, 256, 256, device=device)
    assert_close(ops.upfirdn2d.upsample2d(ones, f)[:, :, 4:-4, 4:-4], torch.ones(2, 8, 504, 504), 1e-6, 'up2 DC gain')
    assert_close(ops.upfirdn2d.downsample2d(ones, f)[:, :, 2:-2, 2:-2], torch.ones(2, 8, 124, 124), 1e-6, 'down2 DC gain')


@pytest.mark.parametrize('shape,const_noise', [((3, 8, 16, 20), False), ((2, 5, 7, 9), True), ((2, 16, 32, 32), True)])
def test_bias_act_with_folded_noise(ops, device, shape, const_noise):
    # bias_act(x, b, noise=n) == bias_act(x + n, b) of the reference (networks.py:904-921: noise add, then bias_act), with
    # gradients for x, b and the noise
    g = torch.Generator().manual_seed(sum(shape))
    N, C, H, W = shape
    x = torch.randn(shape, generator=g); b = torch.randn(C, generator=g)
    nz = torch.randn(H, W, generator=g) if const_noise else torch.randn(N, 1, H, W, generator=g)
    dy = torch.randn(shape, generator=g)
    xs = [t_.clone().requires_grad_(True) for t_ in (x, b, nz)]
    want = R.bias_act(xs[0] + xs[2], xs[1], act='lrelu', gain=1.3, clamp=2.0)
    wg = torch.autograd.grad(want, xs, dy)
    ds = [t_.to(device).requires_grad_(True) for t_ in (x, b, nz)]
    got = ops.bias_act.bias_act(ds[0], ds[1], act='lrelu', gain=1.3, clamp=2.0, noise=ds[2])
    assert_close(got, want, 1e-6, 'y')
    gg_ = torch.autograd.grad(got, ds, dy.to(device))
    for name, a_, b_ in zip(['dx', 'db', 'dnoise'], gg_, wg):
        assert_close(a_, b_, 2e-6, name)


# ----------------------------------------------------------------------------------------------------------------------
# TMA-fed weight-gradient kernel (csrc/wgrad_tma.cu): G operand in tensor memory, kx shift applied by the converters

def _wgrad_oracle(x, dy, k, pad, a=None, b=None):
    """dw[o,i,ky,kx] = sum dy[n,o,y,x] * x[n,i,y+ky-pad,x+kx-pad] in fp64 through autograd of the CPU conv
    (conv2d_gradfix.py:175-191); dy may be smaller / larger than the natural output: positions outside x read as zero."""
    xd = (x if a is None else x * a[:, :, None, None]).double()
    dyd = (dy if b is None else dy * b[:, :, None, None]).double()
    N, I, H, W = xd.shape
    OH, OW = dyd.shape[2:]
    # pad x on the bottom/right so that the natural output covers dy's extent (extra rows/cols are zero = "outside x")
    eh, ew = max(0, OH + k - 1 - 2 * pad - H), max(0, OW + k - 1 - 2 * pad - W)
    xp = torch.nn.functional.pad(xd, (0, ew, 0, eh))
    w = torch.zeros(dyd.shape[1], I, k, k, dtype=torch.float64, requires_grad=True)
    y = torch.nn.functional.conv2d(xp, w, padding=pad)[:, :, :OH, :OW]
    dw, = torch.autograd.grad(y, w, dyd)
    return dw


@pytest.mark.parametrize('case', [
    # N, A, HA, WA, B, HB, WB, k, pad
    (4, 32, 16, 16, 32, 16, 16, 3, 1),      # 32-channel tile: three kx slots stacked in one MMA
    (2, 64, 32, 32, 64, 32, 32, 3, 1),      # two kx slots + one
    (2, 48, 24, 24, 40, 24, 24, 3, 1),      # channel tails on both operands
    (1, 160, 16, 20, 136, 16, 20, 3, 1),    # several tiles per operand, one kx per CTA group, width % 16 != 0
    (2, 64, 17, 20, 32, 16, 16, 2, 0),      # phase-major down form: a larger than b
    (2, 32, 16, 16, 128, 17, 20, 2, 1),     # phase-major up form: b larger than a (reads outside a count as zero)
    (3, 96, 12, 12, 96, 12, 12, 1, 0),      # 1x1
    (2, 32, 40, 8, 32, 40, 8, 3, 1),        # narrow maps (width < one 16-pixel strip)
    (2, 32, 35, 36, 64, 35, 36, 3, 2),      # rows % 16 != 0, "full" padding
])
@pytest.mark.parametrize('flip,layout', [(False, 0), (True, 1)])
def test_wgrad_tma_kernel_vs_oracle(ops, device, case, flip, layout):
    N, A, HA, WA, B, HB, WB, k, pad = case
    co = ops.custom_ops
    plugin = co.get_plugin('conv2d_plugin')
    g = torch.Generator().manual_seed(A * 31 + B + k)
    x = torch.randn(N, A, HA, WA, generator=g); dy = torch.randn(N, B, HB, WB, generator=g)
    a = torch.rand(N, A, generator=g) + 0.5; b = torch.rand(N, B, generator=g) + 0.5
    want = _wgrad_oracle(x, dy, k, pad, a, b)
    if flip:
        want = want.flip([2, 3])
    if layout:
        want = want.transpose(0, 1)
    got = plugin.conv2d_wgrad(x.to(device), dy.to(device), (k, k), padding=(pad, pad), flip_w=flip, out_layout=layout,
                              a_scale=a.to(device), b_scale=b.to(device), prec=co.PREC_TF32X3)
    assert plugin.last_wgrad_prec == 3 and tuple(got.shape) == tuple(want.shape)
    assert_close(got, want.float(), 1e-5, 'wgrad(tma)')


def test_wgrad_unaligned_rows_take_the_global_load_kernel(ops, device):
    # row pitches that are not multiples of 16 bytes cannot be TMA sources: wgrad_tc.cu's kernel serves them, same results
    co = ops.custom_ops
    plugin = co.get_plugin('conv2d_plugin')
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 32, 33, 21, generator=g); dy = torch.randn(2, 32, 33, 21, generator=g)
    want = _wgrad_oracle(x, dy, 3, 1)
    got = plugin.conv2d_wgrad(x.to(device), dy.to(device), (3, 3), padding=(1, 1), prec=co.PREC_TF32X3)
    assert_close(got, want.float(), 1e-5, 'wgrad(ldg)')


@pytest.mark.parametrize('kind,N,I,O,R,skips', [('down', 2, 64, 128, 32, True), ('down', 2, 32, 64, 64, True), ('up', 2, 128, 128, 16, True),
                                                ('up', 2, 64, 32, 32, False),     # all four phases share one 128-row tile: nothing to skip
                                                ('up', 2, 32, 64, 16, True)])
def test_wgrad_phase_major_hint_skips_only_dead_taps(ops, device, kind, N, I, O, R, skips):
    """gg_conv2d_wgrad_pm_f32: with the structural hint the entries that are zero by construction in the phase-major weight
    (conv2d_resample.phase_major_weight_down / _up) may be left zero; every live entry must be unchanged."""
    co = ops.custom_ops
    cr = ops.conv2d_resample
    plugin = co.get_plugin('conv2d_plugin')
    g = torch.Generator().manual_seed(R + I)
    w2 = (cr.phase_major_weight_down if kind == 'down' else cr.phase_major_weight_up)(torch.ones(O, I, 3, 3))
    live = cr._pm_live(kind, 3, 3)
    if kind == 'down':      # x phase-major [N,4I,R/2+1,..], dy [N,O,R/2,R/2], pad 0
        x = torch.randn(N, 4 * I, R // 2 + 1, (R // 2 + 1 + 3) // 4 * 4, generator=g); dy = torch.randn(N, O, R // 2, R // 2, generator=g)
        pad = 0
    else:                   # x [N,I,R,R], dy phase-major [N,4O,R+1,..], pad 1
        x = torch.randn(N, I, R, R, generator=g); dy = torch.randn(N, 4 * O, R + 1, (R + 1 + 3) // 4 * 4, generator=g)
        pad = 1
    for flip, layout in ((False, 0), (True, 1)):
        full = plugin.conv2d_wgrad(x.to(device), dy.to(device), (2, 2), padding=(pad, pad), flip_w=flip, out_layout=layout, prec=co.PREC_TF32X3)
        # the hint describes dw in the layout of the weight tensor itself; with out_layout=1 dim 0 <-> dim 1 swap roles
        pm = live.pm if not layout else (3 - live.pm[0], live.pm[1])
        hint = plugin.conv2d_wgrad(x.to(device), dy.to(device), (2, 2), padding=(pad, pad), flip_w=flip, out_layout=layout, prec=co.PREC_TF32X3, pm=pm)
        mask = (w2 != 0)
        if layout:
            mask = mask.transpose(0, 1)
        mask = mask.to(device)
        assert float((hint - full).abs()[mask].max()) <= 1e-6 * float(full.abs().max()), 'live entries changed'
        dead = hint[~mask]
        assert (float((dead != 0).float().mean()) < 1.0) == skips   # something was skipped (when a tile lies inside a phase pair) ...
        tol = 1e-6 * float(full.abs().max())                     # (atomic flush order differs between two launches)
        assert bool(((dead == 0) | ((dead - full[~mask]).abs() <= tol)).all())  # ... and what was not skipped is the plain result


# ------------------------------------------------------------------------------------------------ conv + bias_act in one kernel
@pytest.mark.parametrize('case', [
    # (N, I, O, R, up, down, k, act, modulated, noise)
    (2, 32, 32, 64, 1, 1, 3, 'lrelu', True, 'random'),     # G conv1 on the marching kernel
    (2, 64, 48, 32, 1, 1, 3, 'lrelu', True, 'const'),      # tile kernel, ragged output channels
    (2, 512, 512, 8, 1, 1, 3, 'lrelu', False, None),       # D conv0, N = 256 tiles
    (2, 32, 64, 64, 1, 2, 3, 'lrelu', False, None),        # D conv1: the phase-major 2x2 conv is the last operator
    (2, 64, 32, 16, 2, 1, 3, 'lrelu', True, 'random'),     # G conv0: the FIR is last -> one bias_act launch behind it
    (2, 3, 32, 64, 1, 1, 1, 'lrelu', False, None),         # fromRGB: thin kernel + in-place bias_act
    (2, 16, 16, 32, 1, 1, 3, 'linear', False, None),
    (2, 16, 16, 32, 1, 1, 3, 'relu', True, 'const'),       # not invertible: d dcoef takes the extra convolution
    (2, 16, 16, 32, 1, 1, 3, 'tanh', False, None),         # not fusable: conv kernel + bias_act kernel
])
def test_fused_conv_bias_act_matches_the_unfused_pair(ops, device, case):
    """conv2d_resample(..., epilogue=...) / modulated_conv2d(..., epilogue=...) == bias_act(conv(...) + noise, bias, ...) computed by
    the two separate ops: forward, first-order gradients of every input, and a second-order gradient (the closure R1 / path length
    regularisation need)."""
    N, I, O, Rr, up, down, k, act, mod, noise_kind = case
    ops.conv2d_gradfix.fuse_epilogue = True                  # (off by default: measured neutral on the 1024^2 step)
    g = torch.Generator().manual_seed(Rr + I + O)
    x = torch.randn(N, I, Rr, Rr, generator=g); w = torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)
    b = torch.randn(O, generator=g) * 0.3; s = torch.randn(N, I, generator=g) * 0.5 + 1.0
    Ro = Rr * up // down
    noise = None if noise_kind is None else (torch.randn(Ro, Ro, generator=g) if noise_kind == 'const' else torch.randn(N, 1, Ro, Ro, generator=g)) * 0.3
    dy = torch.randn(N, O, Ro, Ro, generator=g)
    f = ops.upfirdn2d.setup_filter([1, 3, 3, 1]).to(device)
    modconv = ops.networks.modulated_conv2d if hasattr(ops.networks, 'modulated_conv2d') else None
    gain, clamp = (0.7, None) if act != 'relu' else (1.3, 2.0)

    def run(fused):
        ts = dict(x=x, w=w, b=b, s=s)
        if noise is not None:
            ts['noise'] = noise
        ts = {n: t.to(device).requires_grad_(True) for n, t in ts.items()}
        epi = dict(bias=ts['b'], act=act, gain=gain, clamp=clamp)
        if mod:
            kw = dict(x=ts['x'], weight=ts['w'], styles=ts['s'], noise=ts.get('noise'), up=up, padding=k // 2, resample_filter=f, flip_weight=(up == 1))
            y = modconv(**kw, epilogue=epi) if fused else ops.bias_act.bias_act(modconv(**kw), ts['b'], act=act, gain=gain, clamp=clamp)
        else:
            kw = dict(x=ts['x'], w=ts['w'], f=f, up=up, down=down, padding=k // 2, flip_weight=(up == 1))
            y = ops.conv2d_resample.conv2d_resample(**kw, epilogue=epi) if fused else \
                ops.bias_act.bias_act(ops.conv2d_resample.conv2d_resample(**kw), ts['b'], act=act, gain=gain, clamp=clamp)
        names = [n for n in ts if mod or n != 's']
        g1 = torch.autograd.grad(y, [ts[n] for n in names], dy.to(device), create_graph=True)
        g2 = torch.autograd.grad(sum(t.square().sum() for t in g1), [ts['x'], ts['w']], allow_unused=True)
        return [y.detach()] + [t.detach() for t in g1] + [t.detach() for t in g2 if t is not None], ['y'] + ['d' + n for n in names] + ['ddx', 'ddw']

    try:
        got, names = run(True)
    finally:
        ops.conv2d_gradfix.fuse_epilogue = False
    want, _ = run(False)
    assert len(got) == len(want)
    for n, a_, b_ in zip(names, got, want):
        assert_close(a_, b_, 2e-5, f'{case}: {n}')


# ------------------------------------------------------------------------------------------------ mixed precision (row f4)
HALF_ULP = 2.0 ** -11          # one rounding to float16


def test_float16_operators_round_once_at_their_boundary(ops, device):
    """float16 tensors (plain and channels_last, as the reference's `num_fp16_res` blocks hand them over: networks.py:1031-1035) through
    every public operator: float16 out, float16 gradients back, and the value is the fp32 oracle's result on the same float16-valued
    inputs rounded to float16 once (2^-11 of max|y| plus fp32 noise).  CPU twin: tests/test_autograd_algebra.py."""
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 6, 20, 20, generator=g).half()
    w = (torch.randn(5, 6, 3, 3, generator=g) * 0.2).half()
    b = (torch.randn(6, generator=g) * 0.3).half()
    f = R.setup_filter([1, 3, 3, 1])
    fd = f.to(device)
    bound = HALF_ULP + 2e-5
    o = ops
    cases = {
        'bias_act lrelu clamp': (lambda t: o.bias_act.bias_act(t, b.to(device), act='lrelu', clamp=0.8),
                                 lambda t: R.bias_act(t, b.float(), act='lrelu', clamp=0.8)),
        'bias_act linear gain': (lambda t: o.bias_act.bias_act(t, None, act='linear', gain=0.5), lambda t: R.bias_act(t, None, act='linear', gain=0.5)),
        'upsample2d': (lambda t: o.upfirdn2d.upsample2d(t, fd), lambda t: R.upsample2d(t, f)),
        'downsample2d': (lambda t: o.upfirdn2d.downsample2d(t, fd), lambda t: R.downsample2d(t, f)),
        'conv2d_resample plain': (lambda t: o.conv2d_resample.conv2d_resample(t, w.to(device), padding=1), lambda t: R.conv2d_resample(t, w.float(), padding=1)),
        'conv2d_resample up2': (lambda t: o.conv2d_resample.conv2d_resample(x=t, w=w.to(device), f=fd, up=2, padding=1, flip_weight=False),
                                lambda t: R.conv2d_resample(x=t, w=w.float(), f=f, up=2, padding=1, flip_weight=False)),
        'conv2d_resample down2': (lambda t: o.conv2d_resample.conv2d_resample(x=t, w=w.to(device), f=fd, down=2, padding=1),
                                  lambda t: R.conv2d_resample(x=t, w=w.float(), f=f, down=2, padding=1)),
        'conv2d stride 2': (lambda t: o.conv2d_gradfix.conv2d(t, w.to(device), stride=2, padding=1),
                            lambda t: torch.nn.functional.conv2d(t, w.float(), stride=2, padding=1)),
        'conv_transpose2d stride 2': (lambda t: o.conv2d_gradfix.conv_transpose2d(input=t, weight=w.to(device).transpose(0, 1), stride=2),
                                      lambda t: torch.nn.functional.conv_transpose2d(t, w.float().transpose(0, 1), stride=2)),
    }
    for name, (mine, ref) in cases.items():
        t32 = x.float().requires_grad_(True)
        y32 = ref(t32)
        r = torch.randn(y32.shape, generator=torch.Generator().manual_seed(5)).half()
        gx32, = torch.autograd.grad(y32, t32, r.float())
        for fmt in (torch.contiguous_format, torch.channels_last):
            t16 = x.to(device).to(memory_format=fmt).requires_grad_(True)
            y = mine(t16)
            assert y.dtype == torch.float16 and y.shape == y32.shape, name
            assert_close(y.float(), y32, bound, name)
            gx, = torch.autograd.grad(y, t16, r.to(device))
            assert gx.dtype == torch.float16, name
            assert_close(gx.float(), gx32, bound, name + ': dx')
    with pytest.raises(RuntimeError):
        o.bias_act.bias_act(x.to(device).double(), None)          # float64 is not served


@pytest.mark.parametrize('demodulate,up,fused', [(True, 1, False), (True, 2, False), (False, 1, False), (True, 1, True)])
def test_float16_modulated_conv2d_vs_oracle(ops, device, demodulate, up, fused):
    """modulated_conv2d on float16 activations (networks.py:591-668 incl. the fp16 pre-normalisation :621-627), without and with the
    fused bias / noise / activation epilogue the fused SynthesisLayer forward hands it: one float16 rounding away from the fp32 oracle."""
    g = torch.Generator().manual_seed(12)
    x = torch.randn(2, 16, 24, 24, generator=g).half()
    w = torch.randn(8, 16, 3, 3, generator=g) * 0.2
    s = torch.randn(2, 16, generator=g) * 0.5 + 1
    b = (torch.randn(8, generator=g) * 0.3).half().float()          # float16-valued: the layer hands `bias.to(x.dtype)` over
    noise = torch.randn(24 * up, 24 * up, generator=g) * 0.1
    f = R.setup_filter([1, 3, 3, 1])
    kw = dict(up=up, padding=1, demodulate=demodulate, flip_weight=(up == 1))
    wr, sr = w.clone().requires_grad_(True), s.clone().requires_grad_(True)
    y32 = R.modulated_conv2d(x.float(), wr, sr, noise=noise, resample_filter=f, fused_modconv=False, **kw)
    if fused:
        y32 = R.bias_act(y32, b, act='lrelu', clamp=256.0)
    wd, sd = w.to(device).requires_grad_(True), s.to(device).requires_grad_(True)
    epi = dict(epilogue=dict(bias=b.to(device).half(), act='lrelu', gain=None, clamp=256.0)) if fused else {}
    y = ops.networks.modulated_conv2d(x=x.to(device), weight=wd, styles=sd, noise=noise.to(device), resample_filter=f.to(device), **kw, **epi)
    assert y.dtype == torch.float16
    assert_close(y.float(), y32, HALF_ULP + 2e-5, 'y')
    r = torch.randn(y32.shape, generator=torch.Generator().manual_seed(6)).half()
    got = torch.autograd.grad(y, [wd, sd], r.to(device))
    want = torch.autograd.grad(y32, [wr, sr], r.float())
    for nm, u, v in zip(('dweight', 'dstyles'), got, want):
        assert u.dtype == torch.float32
        # the rounding of y does not enter: both sides differentiate the unrounded fp32 expression.  With the leaky-ReLU epilogue one
        # of the 9216 arguments may change sign between the GPU and the CPU evaluation (probability ~1 %): that moves dweight by
        # ~1e-2 of its maximum and is not an error, so the fused case carries a bound that only a wiring mistake exceeds
        assert_close(u, v, 5e-2 if fused else 1e-4, nm)


def test_mixed_precision_networks_vs_the_live_reference(device):
    """The reference's default configuration (`num_fp16_res`, `conv_clamp=256`; train.py:267-268,425-429): its own Generator and
    Discriminator built that way, on the library, are closer to the reference's fp32 evaluation (`force_fp32=True`, CPU) than the
    reference's own float16 evaluation is: image, logits, every generator gradient of a non-saturating loss."""
    from oracle import live_ref
    from tests.util import reference_networks, quiet, max_rel_err
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent')
    L = live_ref.load()
    networks = reference_networks()
    kw_g = dict(z_dim=32, c_dim=0, w_dim=32, img_resolution=64, img_channels=3, mapping_kwargs=dict(num_layers=2),
                synthesis_kwargs=dict(channel_base=1024, channel_max=32, num_fp16_res=3, conv_clamp=256))
    kw_d = dict(c_dim=0, img_resolution=64, img_channels=3, channel_base=1024, channel_max=32, num_fp16_res=3, conv_clamp=256,
                epilogue_kwargs=dict(mbstd_group_size=2))
    torch.manual_seed(4)
    G_ref, D_ref = quiet(L.networks.Generator, **kw_g).train(), quiet(L.networks.Discriminator, **kw_d).train()
    with torch.no_grad():
        for p_ in list(G_ref.parameters()) + list(D_ref.parameters()):
            if float(p_.abs().max()) == 0:
                p_.copy_(torch.randn(p_.shape) * 0.1)
    G, D = quiet(networks.Generator, **kw_g).train(), quiet(networks.Discriminator, **kw_d).train()
    G.load_state_dict(G_ref.state_dict()); D.load_state_dict(D_ref.state_dict())
    G, D = G.to(device), D.to(device)
    z = torch.randn(4, 32, generator=torch.Generator().manual_seed(2)); c = torch.zeros(4, 0)

    def run(Gn, Dn, dev, **kw):
        for p_ in list(Gn.parameters()) + list(Dn.parameters()):
            p_.grad = None
        img = Gn(z.to(dev), c.to(dev), noise_mode='const', **kw)
        logits = Dn(img, c.to(dev), **kw)
        torch.nn.functional.softplus(-logits).mean().backward()
        return img.detach().cpu(), logits.detach().cpu(), {k: p_.grad.cpu() for k, p_ in Gn.named_parameters() if p_.grad is not None}

    mine = run(G, D, device)
    ref16, ref32 = run(G_ref, D_ref, 'cpu'), run(G_ref, D_ref, 'cpu', force_fp32=True)
    assert mine[0].dtype == torch.float32 and set(mine[2]) == set(ref32[2])
    for i, nm in ((0, 'image'), (1, 'logits')):
        e_mine, e_ref16 = max_rel_err(mine[i], ref32[i]), max_rel_err(ref16[i], ref32[i])
        assert e_mine <= 1.5 * e_ref16 + 2e-4, (nm, e_mine, e_ref16)
    live = [k for k in ref32[2] if float(ref32[2][k].abs().max()) > 0]
    worst_mine = max(max_rel_err(mine[2][k], ref32[2][k]) for k in live)
    worst_ref16 = max(max_rel_err(ref16[2][k], ref32[2][k]) for k in live)
    assert worst_mine <= 1.5 * worst_ref16 + 1e-3, (worst_mine, worst_ref16)


# ------------------------------------------------------------------------------------------------ rosinality adapter (row f4)
@pytest.mark.parametrize('fused_layers', [True, False])
def test_rosinality_networks_vs_the_unmodified_module(device, fused_layers):
    """GA-GAN's second StyleGAN2 code base (SimilarDomains/gan_models/StyleGAN2/model.py) bound to the library by
    gagan_b200.install_rosinality, against the same file unmodified on the CPU (its torch-native ops): image from W and from S codes,
    logits, generator and R1 gradients.  Image and logits: within 2e-5.  Gradients: these networks are 512 channels wide at every
    size, so between ANY two fp32 implementations an occasional leaky-ReLU argument changes sign and moves every upstream gradient by
    1e-4 ... 2e-2 of its maximum, and the R1 (double-backward) bias gradients are ill-conditioned as in the Dreg study of DESIGN.md.
    Calibrated on the CPU by running this very test against a stand-in whose convolutions sum their channels in a permuted order (10
    seeds: G max-rel typically 1e-6, 4e-3 with a sign flip; R1 90 %-quantile up to 2e-3, max 4e-3; whole-gradient L2 error <= 1e-4),
    the bounds are: per tensor max-rel <= 1e-1, 90 %-quantile <= 1e-2 (G) / 2e-2 (R1), median <= 2e-3, whole-gradient relative L2
    error <= 5e-3.  The exact comparison is the fp64 CPU twin in tests/test_autograd_algebra.py (1e-9)."""
    from oracle import live_ref
    from tests.util import rosinality_model, max_rel_err
    if not live_ref.rosinality_available():
        pytest.skip('oracle/_ref/SimilarDomains is absent')
    ref = live_ref.load_rosinality()
    mine = rosinality_model(fused_layers=fused_layers)
    size, style_dim = 16, 32
    torch.manual_seed(7)
    G_ref, D_ref = ref.Generator(size, style_dim, 2, channel_multiplier=1), ref.Discriminator(size, channel_multiplier=1)
    with torch.no_grad():
        for p_ in list(G_ref.parameters()) + list(D_ref.parameters()):
            if float(p_.abs().max()) == 0:
                p_.copy_(torch.randn(p_.shape) * 0.1)
    G, D = mine.Generator(size, style_dim, 2, channel_multiplier=1), mine.Discriminator(size, channel_multiplier=1)
    G.load_state_dict(G_ref.state_dict()); D.load_state_dict(D_ref.state_dict())
    G, D = G.to(device), D.to(device)
    z = torch.randn(2, style_dim, generator=torch.Generator().manual_seed(2))
    real = torch.rand(2, 3, size, size, generator=torch.Generator().manual_seed(3)) * 2 - 1
    noises = [torch.randn(1, 1, 2 ** (2 + (i + 1) // 2), 2 ** (2 + (i + 1) // 2), generator=torch.Generator().manual_seed(10 + i)) for i in range(G_ref.num_layers)]

    def run(Gn, Dn, dev):
        for p_ in list(Gn.parameters()) + list(Dn.parameters()):
            p_.grad = None
        nz = [n.to(dev) for n in noises]
        img, _ = Gn([z.to(dev)], noise=nz)
        logits = Dn(img)
        torch.nn.functional.softplus(-logits).mean().backward()
        g_grads = {k: p_.grad.cpu() for k, p_ in Gn.named_parameters() if p_.grad is not None}
        x = real.to(dev).requires_grad_(True)
        r1, = torch.autograd.grad(Dn(x).sum(), x, create_graph=True)
        for p_ in Dn.parameters():
            p_.grad = None
        r1.square().sum([1, 2, 3]).mean().backward()
        d_grads = {k: p_.grad.cpu() for k, p_ in Dn.named_parameters() if p_.grad is not None}
        with torch.no_grad():
            img_s = Gn(Gn.get_s_code([Gn.style(z.to(dev))], input_is_latent=True), is_s_code=True, noise=nz)[0]
        return img.detach().cpu(), logits.detach().cpu(), g_grads, d_grads, img_s.cpu()

    img, logits, gg, dg, img_s = run(G, D, device)
    img_r, logits_r, gg_r, dg_r, img_s_r = run(G_ref, D_ref, 'cpu')
    assert_close(img, img_r, 2e-5, 'image'); assert_close(logits, logits_r, 2e-5, 'logits'); assert_close(img_s, img_s_r, 2e-5, 'image from S codes')
    assert set(gg) == set(gg_r) and set(dg) == set(dg_r)
    for nm, got, want, q90 in (('G', gg, gg_r, 1e-2), ('R1', dg, dg_r, 2e-2)):
        errs = sorted((max_rel_err(got[k], want[k]), k) for k in want if float(want[k].abs().max()) > 0)
        whole = float(torch.cat([(got[k] - want[k]).flatten() for k in want]).norm() / torch.cat([want[k].flatten() for k in want]).norm())
        print(f'rosinality {nm} gradients (fused_layers={fused_layers}): median {errs[len(errs) // 2][0]:.1e}, 90 % {errs[int(0.9 * (len(errs) - 1))][0]:.1e}, '
              f'max {errs[-1][0]:.1e} ({errs[-1][1]}), whole-gradient L2 {whole:.1e}')
        assert errs[-1][0] <= 1e-1, (nm, errs[-3:])
        assert errs[int(0.9 * (len(errs) - 1))][0] <= q90, (nm, errs[int(0.9 * (len(errs) - 1))], errs[-3:])
        assert errs[len(errs) // 2][0] <= 2e-3, (nm, errs[len(errs) // 2])
        assert whole <= 5e-3, (nm, whole)


# ------------------------------------------------------------------------------------------------ fma (row a6)
@pytest.mark.parametrize('c_shape', ['N1HW', '11HW', 'HW', 'general'])
def test_fma_vs_oracle(ops, device, c_shape):
    """fma.fma (fma.py:15-58) on the device against the oracle: the shape of its one call site (networks.py:648: activations x
    per-sample channel scale + a noise plane) runs gg_fma_rows_f32 -- vectorised and (odd width) scalar variant --, any other broadcast
    pattern the reference's torch.addcmul; gradients incl. the unbroadcast sums, and a second-order gradient."""
    g = torch.Generator().manual_seed(3)
    for N, C, H, W in ((3, 5, 8, 12), (2, 4, 5, 7)):
        a = torch.randn(N, C, H, W, generator=g)
        b = torch.randn(N, C, 1, 1, generator=g) if c_shape != 'general' else torch.randn(1, C, 1, W, generator=g)
        c = torch.randn({'N1HW': (N, 1, H, W), '11HW': (1, 1, H, W), 'HW': (H, W), 'general': (N, C, 1, 1)}[c_shape], generator=g)
        r = torch.randn(N, C, H, W, generator=g)
        assert ops.fma._rows_shape(a.to(device), b.to(device), c.to(device)) == (c_shape != 'general')

        def run(fn, dev):
            ts = [t.to(dev).requires_grad_(True) for t in (a, b, c)]
            y = fn(*ts)
            first = torch.autograd.grad((y * r.to(dev)).sum(), ts, create_graph=True)
            second = torch.autograd.grad(sum(t.square().sum() for t in first), ts, allow_unused=True)
            return [y] + list(first) + [t for t in second if t is not None]
        n0 = ops.custom_ops.launch_count()
        got, want = run(ops.fma.fma, device), run(R.fma, 'cpu')
        assert (ops.custom_ops.launch_count() > n0) == (c_shape != 'general')
        assert len(got) == len(want)
        for i, (u, v) in enumerate(zip(got, want)):
            assert_close(u, v, 1e-5, f'fma {c_shape} {N}x{C}x{H}x{W} output {i}')


# ------------------------------------------------------------------------------------------------ unaligned conv rows on tcgen05
@pytest.mark.parametrize('case', [
    # N, I, O, H, W, k, pad, transposed  -- widths that are not multiples of 4: the TMA row pitch rule sends them to the FFMA kernel
    (2, 32, 32, 40, 67, 3, 1, False), (1, 64, 64, 70, 130, 3, 1, False), (2, 16, 128, 33, 65, 1, 0, False), (1, 128, 64, 20, 66, 3, 1, True),
])
def test_unaligned_conv_rows_stay_on_the_tensor_cores_when_padded(ops, device, case):
    """conv2d_gradfix.pad_unaligned_rows (default off): with the switch on, such inputs are zero-padded to a multiple
    of 4 columns on the host and run on the tcgen05 kernels (last_conv_prec == 3xTF32) with the results of the FFMA path."""
    N, I, O, H, W, k, pad, transposed = case
    g = torch.Generator().manual_seed(N * 100 + W)
    x = torch.randn(N, I, H, W, generator=g)
    w = torch.randn(*((I, O, k, k) if transposed else (O, I, k, k)), generator=g) / np.sqrt(I * k * k)
    xc, wc = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    want = _conv_oracle(xc, wc, 1, pad, transposed)
    dy = torch.randn(want.shape, generator=g)
    wdx, wdw = torch.autograd.grad(want, [xc, wc], dy)
    fn = ops.conv2d_gradfix.conv_transpose2d if transposed else ops.conv2d_gradfix.conv2d
    plugin = ops.custom_ops.get_plugin('conv2d_plugin')
    used = {}
    for on in (False, True):
        ops.conv2d_gradfix.pad_unaligned_rows = on
        try:
            xg, wg = x.to(device).requires_grad_(True), w.to(device).requires_grad_(True)
            got = fn(xg, wg, padding=pad)
            used[on] = plugin.last_conv_prec
            assert got.shape == want.shape
            assert_close(got, want, 2e-5, f'fwd (padded={on})')
            gdx, gdw = torch.autograd.grad(got, [xg, wg], dy.to(device))
            assert_close(gdx, wdx, 2e-5, f'dgrad (padded={on})')
            assert_close(gdw, wdw, 1e-4, f'wgrad (padded={on})')
        finally:
            ops.conv2d_gradfix.pad_unaligned_rows = False
    assert used[False] == ops.custom_ops.PREC_FP32_SIMT and used[True] == ops.custom_ops.PREC_TF32X3, used


@pytest.mark.parametrize('case', [
    # name, shape, pad -- unit-rate filters on rows that are not 16-byte multiples, wide enough for the shared-memory row exchange of
    # fir_stream (coalesced input window from 256 output columns, coalesced output segments from 384).  The same code path ran the
    # 129^2 and 77x203 cases above while the thresholds stood at 96 columns during development; these are its shipped sizes.
    ('filt_p1_257', (1, 2, 257, 257), [1, 1, 1, 1]), ('filt_p2_384', (1, 2, 384, 384), [2, 2, 2, 2]),
    ('filt_both_odd_wide', (1, 1, 301, 515), [1, 1, 1, 1]), ('filt_p1_1025', (1, 1, 70, 1025), [1, 1, 2, 0]),
], ids=lambda c: c[0])
def test_upfirdn2d_wide_unaligned_rows_vs_oracle(ops, device, case):
    name, shape, pad = case
    test_upfirdn2d_tiled_vs_oracle(ops, device, (name, shape, 1, 1, pad))

