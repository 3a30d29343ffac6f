"""GPU: parity against the CPU oracle AT BASELINE-CONFIG SIZES ("what's weak" 1-3).

  * cfg 1 verbatim -- paper256 256^2 generator forward, batch 4, eval, const noise, seed 0 (ref
    networks.py:1137-1171): the reference's Generator on this build's operators vs the SAME class on the reference's own CPU
    impl='ref' operators (oracle/live_ref.py);
  * config-f single layers at N = 1-2 -- forward, data gradient and weight gradient of 32@1024^2, 64@512^2, 128@256^2,
    512@64^2 and of every up / down variant, plain and modulated, vs the fp64 CPU oracle (oracle/ops_ref.py);
  * one Dreg (R1, double backward) phase of the paper256 256^2 discriminator vs the live reference;
  * the truncation compensation of the tensor-core accumulator on data that is NOT zero-mean noise (post-lrelu sparse,
    constant, heavy-tailed, all-positive);
  * repeat-equality of the deterministic kernels on the layer shape that failed once in round 1 (64->32 up @512^2);
  * the fast mode (one TF32 product): measured and written to gpurun_out/, never asserted against the fp32 tolerance.

Tolerance: max|a-b| / max|b| <= 1e-3 is the contract (tests.util.TOL); the asserted bounds below are much tighter where the
3xTF32 path is expected to be fp32-faithful, so a regression of the numerics shows up long before the contract breaks.
"""
import io
import os
import time
import contextlib
import numpy as np
import pytest
import torch

from tests.util import ROOT, TOL, assert_close, max_rel_err, patched_randn, reference_networks, quiet
from oracle import ops_ref as R

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ops(device):
    import types
    from torch_utils import custom_ops
    from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma
    import gagan_b200.training.networks as mine
    custom_ops.load_library()
    return types.SimpleNamespace(upfirdn2d=upfirdn2d, bias_act=bias_act, conv2d_resample=conv2d_resample, conv2d_gradfix=conv2d_gradfix,
                                 fma=fma, custom_ops=custom_ops, modulated_conv2d=mine.modulated_conv2d)


@pytest.fixture(scope='module')
def live():
    from oracle import live_ref
    if not live_ref.available():
        pytest.skip('oracle/_ref is absent (tools/vendor_reference.py)')
    return live_ref.load()


# ------------------------------------------------------------------------------------------------ cfg 1 verbatim
def test_cfg1_paper256_generator_forward_batch4_matches_the_live_reference(device, live):
    networks = reference_networks()
    kw = dict(z_dim=512, c_dim=0, w_dim=512, img_resolution=256, img_channels=3, mapping_kwargs=dict(num_layers=8),
              synthesis_kwargs=dict(channel_base=16384, channel_max=512, num_fp16_res=0, conv_clamp=None))   # train.py:224,264-268,421-423
    torch.manual_seed(0)
    G_cpu = quiet(live.networks.Generator, **kw).eval()
    z = torch.randn(4, 512); c = torch.zeros(4, 0)
    t0 = time.perf_counter()
    with torch.no_grad():
        ws_cpu = G_cpu.mapping(z, c)
        img_cpu = G_cpu.synthesis(ws_cpu, noise_mode='const')
    t_cpu = time.perf_counter() - t0
    G = quiet(networks.Generator, **kw).eval()
    G.load_state_dict(G_cpu.state_dict())
    G = G.to(device)
    with torch.no_grad():
        ws = G.mapping(z.to(device), c.to(device))
        img = G.synthesis(ws, noise_mode='const')
        img_t = G(z.to(device), c.to(device), truncation_psi=0.5, noise_mode='const')
        img_t_cpu = G_cpu(z, c, truncation_psi=0.5, noise_mode='const')
    assert img.shape == (4, 3, 256, 256)
    e_ws = assert_close(ws, ws_cpu, 1e-5, 'cfg1 mapping')
    e_img = assert_close(img, img_cpu, 2e-5, 'cfg1 synthesis')
    e_t = assert_close(img_t, img_t_cpu, 2e-5, 'cfg1 truncated')
    print(f'cfg1: max-rel-err ws {e_ws:.2e} img {e_img:.2e} truncated {e_t:.2e}; reference CPU forward {t_cpu:.2f} s '
          f'({torch.get_num_threads()} threads)')


# ------------------------------------------------------------------------------------------------ config-f layers vs fp64
LAYERS = [
    # (N, I, O, R_in, up, down, modulated): config-f 1024^2 layer shapes
    (1, 32, 32, 1024, 1, 1, True),     # G b1024.conv1
    (1, 32, 32, 1024, 1, 1, False),    # D b1024.conv0
    (2, 64, 64, 512, 1, 1, True),      # G b512.conv1
    (2, 128, 128, 256, 1, 1, False),   # D b256.conv0
    (2, 512, 512, 64, 1, 1, True),     # G b64.conv1
    (1, 64, 32, 512, 2, 1, True),      # G b1024.conv0 (up)
    (2, 128, 64, 256, 2, 1, True),     # G b512.conv0 (up)
    (2, 512, 512, 32, 2, 1, True),     # G b64.conv0 (up)
    (1, 32, 64, 1024, 1, 2, False),    # D b1024.conv1 (down)
    (2, 64, 128, 512, 1, 2, False),    # D b512.conv1 (down)
    (2, 512, 512, 64, 1, 2, False),    # D b64.conv1 (down)
    (2, 512, 3, 64, 1, 1, 'rgb'),      # G b64.torgb (1x1, demodulate=False)
]


@pytest.mark.parametrize('case', LAYERS, ids=lambda c: 'N{}_{}to{}_r{}_up{}_down{}_{}'.format(*c))
def test_config_f_layer_forward_dgrad_wgrad_match_the_fp64_oracle(ops, device, case):
    N, I, O, Rin, up, down, mod = case
    k = 1 if mod == 'rgb' else 3
    g = torch.Generator().manual_seed(Rin * 7 + I + up + 3 * down)
    x = torch.randn(N, I, Rin, Rin, generator=g)
    w = torch.randn(O, I, k, k, generator=g) / np.sqrt(k * k * I)
    s = torch.randn(N, I, generator=g) * 0.5 + 1.0                      # styles (SURVEY.md section 8(d) cfg 5)
    f = R.setup_filter([1, 3, 3, 1])
    Rout = Rin * up // down
    dy = torch.randn(N, O, Rout, Rout, generator=g)

    def run(x_, w_, s_, f_, dy_, modconv, resample):
        x_ = x_.requires_grad_(True); w_ = w_.requires_grad_(True)
        if mod:
            s_ = s_.requires_grad_(True)
            y = modconv(x=x_, weight=w_, styles=s_, up=up, padding=k // 2, resample_filter=f_, flip_weight=(up == 1),
                        demodulate=(mod != 'rgb'), fused_modconv=False)
            grads = torch.autograd.grad(y, [x_, w_, s_], dy_)
        else:
            y = resample(x=x_, w=w_, f=f_, up=up, down=down, padding=1, flip_weight=(up == 1))
            grads = torch.autograd.grad(y, [x_, w_], dy_)
        return [y.detach()] + [t.detach() for t in grads]

    got = run(x.to(device), w.to(device), s.to(device), f.to(device), dy.to(device), ops.modulated_conv2d, ops.conv2d_resample.conv2d_resample)
    torch.cuda.synchronize()
    want = run(x.double(), w.double(), s.double(), f, dy.double(), R.modulated_conv2d, R.conv2d_resample)
    names = ['y', 'dx', 'dw'] + (['dstyles'] if mod else [])
    errs = {}
    for name, a, b in zip(names, got, want):
        assert a.shape == b.shape, name
        errs[name] = assert_close(a, b, 1e-5, f'{name} of {case}')        # fp32-faithful: two orders below the 1e-3 contract
    print(f'{case}: ' + ' '.join(f'{n} {e:.1e}' for n, e in errs.items()))


# ------------------------------------------------------------------------------------------------ one Dreg phase at 256^2
def test_dreg_phase_paper256_256_matches_the_live_reference(device, live):
    """R1 regularisation of the paper256 discriminator at 256^2, batch 4: parameter gradients after the reference's own
    StyleGAN2Loss.accumulate_gradients('Dreg') -- double backward through every operator -- on the library, against

      * the fp64 truth (the oracle's restatement of the phase, oracle/networks_ref.py, on the same weights),
      * the reference on the CPU (fp32, impl='ref'),
      * the reference ON THE GPU: its impl='ref' torch ops on CUDA tensors = cuDNN fp32 with TF32 off, i.e. what a user of the
        reference gets today.

    These gradients are ILL-CONDITIONED at config size: the bias gradients only flow through the minibatch-std layer and every
    lrelu mask that flips under a rounding error moves them by a finite amount.  profiles/r2_dreg_conditioning.txt (tools/
    dreg_conditioning.py, 4 seeds): the fp64 truth itself moves by up to 1.2e-3 when the input image is perturbed by 1e-7, the
    reference on cuDNN is up to 1.5e-2 away from the truth, the reference on the CPU 1.4e-3, this build 1.7e-4 .. 5e-3.  A fixed
    1e-3 against ONE fp32 realisation is therefore not a meaningful bar for every parameter; the assertions are
      (1) the median parameter is within the 1e-3 contract of the truth and of the reference,
      (2) no parameter is further from the truth than 3x the worse of the two fp32 reference realisations (floor 1e-3)."""
    networks = reference_networks()
    from training import loss as loss_mod                               # the checkout's loss.py on this build's operators
    from oracle import networks_ref as NR
    kw = dict(c_dim=0, img_resolution=256, img_channels=3, channel_base=16384, channel_max=512, num_fp16_res=0, conv_clamp=None,
              epilogue_kwargs=dict(mbstd_group_size=4))
    torch.manual_seed(1)
    D_cpu = quiet(live.networks.Discriminator, **kw).train()
    with torch.no_grad():
        for p in D_cpu.parameters():
            if float(p.abs().max()) == 0:
                p.copy_(torch.randn(p.shape) * 0.1)                       # biases are zero-initialised
    real = torch.rand(4, 3, 256, 256) * 2 - 1
    live.conv2d_gradfix.enabled = True

    def phase(loss_cls, D, dev):
        D.requires_grad_(True)
        c = torch.zeros(4, 0, device=dev); z = torch.zeros(4, 512, device=dev)
        loss_cls(device=dev, G_mapping=None, G_synthesis=None, D=D, r1_gamma=1.0).accumulate_gradients(
            phase='Dreg', real_img=real.to(dev), real_c=c, gen_z=z, gen_c=c, sync=True, gain=16)
        return {n: p.grad.detach().double().cpu() for n, p in D.named_parameters() if p.grad is not None}

    D = quiet(networks.Discriminator, **kw).train(); D.load_state_dict(D_cpu.state_dict())
    ours = phase(loss_mod.StyleGAN2Loss, D.to(device), device)
    D_ref_gpu = quiet(live.networks.Discriminator, **kw).train(); D_ref_gpu.load_state_dict(D_cpu.state_dict())
    ref_gpu = phase(live.loss.StyleGAN2Loss, D_ref_gpu.to(device), device)
    ref_cpu = phase(live.loss.StyleGAN2Loss, D_cpu, torch.device('cpu'))
    PD = {k: v.detach().double().requires_grad_(v.dtype.is_floating_point) for k, v in D_cpu.state_dict().items()}
    names = [n for n, _ in D_cpu.named_parameters()]
    grads = torch.autograd.grad(NR.loss_Dr1(PD, real.double(), 256, 1.0, 4) * 16, [PD[n] for n in names], allow_unused=True)
    truth = {n: g for n, g in zip(names, grads) if g is not None and float(g.abs().max()) > 0}
    rows = sorted(((max_rel_err(ours[n], truth[n]), max_rel_err(ref_gpu[n], truth[n]), max_rel_err(ref_cpu[n], truth[n]),
                    max_rel_err(ours[n], ref_cpu[n]), n) for n in truth), reverse=True)
    print('Dreg 256^2 paper256, parameter gradients, max-rel-err vs the fp64 truth of [this build | reference on cuDNN | reference on CPU] '
          'and of this build vs the CPU reference')
    for e_us, e_rg, e_rc, e_pair, name in rows[:8]:
        print(f'   {e_us:.2e} | {e_rg:.2e} | {e_rc:.2e} || {e_pair:.2e}   {name}')
    med = [float(np.median([r[i] for r in rows])) for i in range(4)]
    print(f'   median over {len(rows)} parameters: {med[0]:.2e} | {med[1]:.2e} | {med[2]:.2e} || {med[3]:.2e}')
    assert med[0] <= TOL and med[3] <= TOL
    for e_us, e_rg, e_rc, e_pair, name in rows:
        assert e_us <= max(TOL, 3 * max(e_rg, e_rc)), f'Dreg grad {name}: {e_us:.2e} from the truth (reference: cuDNN {e_rg:.2e}, CPU {e_rc:.2e})'


# ------------------------------------------------------------------------------------------------ rz_compensation
def _dist(kind, shape, g):
    if kind == 'post_lrelu_sparse':          # what the layers actually see: lrelu output, 80 % of the negative side squashed
        return torch.nn.functional.leaky_relu(torch.randn(shape, generator=g), 0.2) * np.sqrt(2)
    if kind == 'relu_sparse':                # hard zeros on half of the inputs
        return torch.relu(torch.randn(shape, generator=g))
    if kind == 'constant':                   # a flat image: every product has the same sign per weight
        return torch.full(shape, 0.75)
    if kind == 'heavy_tailed':               # cubed normal: a few large entries dominate every dot product
        return torch.randn(shape, generator=g) ** 3
    if kind == 'all_positive':               # running sums only grow: the worst case for a truncating accumulator
        return torch.rand(shape, generator=g) + 0.5
    raise KeyError(kind)


@pytest.mark.parametrize('kind', ['post_lrelu_sparse', 'relu_sparse', 'constant', 'heavy_tailed', 'all_positive'])
@pytest.mark.parametrize('shape', [(2, 512, 512, 32), (2, 32, 32, 128), (2, 128, 128, 64)], ids=['512ch', '32ch', '128ch'])
def test_truncation_compensation_on_structured_data(ops, device, kind, shape):
    """tc_common.cuh::rz_compensation is an EXPECTED-VALUE correction calibrated on random data.  Its assumption (the mean of what
    truncation removes is proportional to the chunk sum) is probed here on data where it is least safe.  Both the weights and
    the inputs follow the distribution for 'all_positive'; otherwise the weights are random-sign.  Reported per case: max
    relative error and the SIGNED mean error relative to the mean magnitude (the bias the compensation is there to remove),
    next to the same two numbers of the exact-fp32 FFMA kernel."""
    N, I, O, Rr = shape
    g = torch.Generator().manual_seed(['post_lrelu_sparse', 'relu_sparse', 'constant', 'heavy_tailed', 'all_positive'].index(kind) * 101 + I)
    x = _dist(kind, (N, I, Rr, Rr), g)
    w = (torch.rand(O, I, 3, 3, generator=g) + 0.1 if kind == 'all_positive' else torch.randn(O, I, 3, 3, generator=g)) / np.sqrt(9 * I)
    dy = _dist(kind, (N, O, Rr, Rr), g)
    plugin = ops.custom_ops.get_plugin('conv2d_plugin')
    co = ops.custom_ops
    want = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    want_dw = torch.nn.grad.conv2d_weight(x.double(), w.shape, dy.double(), padding=1)
    rows = []
    for prec, label in ((co.PREC_TF32X3, 'tcgen05 3xTF32'), (co.PREC_FP32_SIMT, 'FFMA fp32')):
        y = plugin.conv2d(x.to(device), w.to(device), padding=(1, 1), prec=prec).cpu().double()
        dw = plugin.conv2d_wgrad(x.to(device), dy.to(device), (3, 3), padding=(1, 1), prec=prec).cpu().double()
        for name, a, b in (('conv', y, want), ('wgrad', dw, want_dw)):
            rel = float((a - b).abs().max() / b.abs().max())
            bias = float((a - b).mean() / b.abs().mean())
            rows.append((label, name, rel, bias))
            if prec == co.PREC_TF32X3:
                assert rel <= 2e-5, f'{kind} {shape} {name}: max-rel-err {rel:.2e}'
                # measured residual of the expected-value compensation: conv <= 1.4e-7, wgrad <= 2.7e-6 (constant images, where
                # truncation removes about half of what random data loses): a uniform scale error of that size, no spatial structure
                assert abs(bias) <= (5e-7 if name == 'conv' else 5e-6), f'{kind} {shape} {name}: signed mean error {bias:.2e} of the mean magnitude'
    print(f'{kind} {shape}: ' + '; '.join(f'{l} {n}: max {r:.1e} bias {b:+.1e}' for l, n, r, b in rows))


# ------------------------------------------------------------------------------------------------ repeat equality
def test_repeat_equality_of_the_layer_that_failed_once(ops, device):
    """Round 1 saw ONE unexplained failure of the 64->32 up-sampling layer at 512^2 (linearity / adjoint test) in ~30 suite runs.
    The forward and data-gradient kernels are deterministic: 200 repetitions must reproduce the first result bit for bit
    (a race in the TMA / mbarrier / TMEM pipelines would show as a mismatch); the weight gradient is flushed with fp32 atomics
    and must stay within 1e-6 of the first result."""
    cr = ops.conv2d_resample
    f = ops.upfirdn2d.setup_filter([1, 3, 3, 1]).to(device)
    torch.manual_seed(0)
    N, I, O, Rr = 2, 64, 32, 512
    x = torch.randn(N, I, Rr, Rr, device=device); w = torch.randn(O, I, 3, 3, device=device) / np.sqrt(9 * I)
    ref, bad = None, [0, 0, 0]
    for rep in range(200):
        xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
        y = cr.conv2d_resample(xr, wr, f=f, up=2, padding=1, flip_weight=False)
        if rep == 0:
            dy = torch.randn_like(y)
        dx, dw = torch.autograd.grad(y, [xr, wr], dy)
        if ref is None:
            ref = (y.detach().clone(), dx.clone(), dw.clone())
            continue
        bad[0] += int(not torch.equal(y.detach(), ref[0]))
        bad[1] += int(not torch.equal(dx, ref[1]))
        bad[2] += int(float((dw - ref[2]).abs().max()) > 1e-6 * float(ref[2].abs().max()))
    torch.cuda.synchronize()
    assert bad == [0, 0, 0], f'mismatching repetitions (forward, data gradient, weight gradient): {bad}'


@pytest.mark.parametrize('form', ['up', 'down'])
def test_weight_gradient_ring_ownership_regression(ops, device, form):
    """The bug behind the rare `unspecified launch failure`s of rounds 1 and 2 (and the one failure of the layer above): the two
    converter groups of the weight-gradient kernels dealt the row tasks out by their index INSIDE a strip, so whenever a strip had an
    odd number of tasks (2x2 taps) or rows (heights 16 n + 1: the phase-major up-sampling layers) the ownership of the operand ring
    slots flipped between the groups, and a group could wait for release #u-1 of a slot before release #u-2 had happened -- which an
    mbarrier parity wait reports as success.  Overwritten operands, over-arrived barriers; rare with three MMAs per product, immediate
    with one.  The one-product mode is therefore the sensitive probe: 12 launches of the two 2x2 forms at the sizes that failed must
    agree with each other to 2e-6 and with the 3xTF32 result to the tf32 rounding level."""
    co = ops.custom_ops
    plugin = co.get_plugin('conv2d_plugin')
    cr = ops.conv2d_resample
    g = torch.Generator(device=device).manual_seed(3)
    n = 8
    if form == 'up':
        a = torch.randn(n, 64, 512, 512, device=device, generator=g); b = torch.randn(n, 128, 513, 516, device=device, generator=g)
        kw = dict(padding=(1, 1), pm=cr._pm_live('up', 3, 3).pm, flip_w=True, out_layout=1)
    else:
        a = torch.randn(n, 128, 513, 516, device=device, generator=g); b = torch.randn(n, 64, 512, 512, device=device, generator=g)
        kw = dict(padding=(0, 0), pm=cr._pm_live('down', 3, 3).pm)
    ref = plugin.conv2d_wgrad(a, b, (2, 2), prec=co.PREC_AUTO, **kw)
    assert plugin.last_wgrad_prec == co.PREC_TF32X3
    scale = float(ref.abs().max())
    first = None
    for rep in range(12):
        dw = plugin.conv2d_wgrad(a, b, (2, 2), prec=co.PREC_AUTO_FAST, **kw)
        if rep % 2 == 0:
            torch.cuda.synchronize()                        # both launch patterns: isolated and back to back
        first = dw if first is None else first
        assert float((dw - first).abs().max()) <= 2e-6 * scale, f'{form}: launch {rep} differs from the first one'
        assert float((dw - ref).abs().max()) <= 2e-3 * scale, f'{form}: launch {rep} is off the 3xTF32 result'
    for rep in range(12):                                   # and the fp32-faithful mode stays put as well
        dw = plugin.conv2d_wgrad(a, b, (2, 2), prec=co.PREC_AUTO, **kw)
        assert float((dw - ref).abs().max()) <= 2e-6 * scale
    torch.cuda.synchronize()


# ------------------------------------------------------------------------------------------------ fma on the device
def test_fma_forward_and_broadcast_gradients_on_the_device(ops, device):
    g = torch.Generator().manual_seed(1)
    a = torch.randn(2, 4, 9, 9, generator=g); b = torch.randn(2, 4, 1, 1, generator=g); c = torch.randn(2, 1, 9, 9, generator=g)
    dy = torch.randn(2, 4, 9, 9, generator=g)

    def run(fn, dev):
        ts = [t.to(dev).requires_grad_(True) for t in (a, b, c)]
        y = fn(*ts)
        return [y.detach().cpu()] + [t.cpu() for t in torch.autograd.grad(y, ts, dy.to(dev))]
    for u, v in zip(run(ops.fma.fma, device), run(R.fma, 'cpu')):
        assert_close(u, v, 1e-6, 'fma')


# ------------------------------------------------------------------------------------------------ fast mode report
def test_fast_mode_report(ops, device):
    """One TF32 product per MAC instead of three (`--prec tf32x1`): measured, written to gpurun_out/r2_fast_mode_parity.txt,
    and only required to stay finite -- it is NOT fp32-faithful and never the headline."""
    co = ops.custom_ops
    lines = []
    old = co.conv_precision
    try:
        for prec, label in ((co.PREC_AUTO, '3xTF32 (headline)'), (co.PREC_AUTO_FAST, "tf32x1 (fast mode)")):
            co.conv_precision = prec
            for case in [(2, 64, 64, 512, 1, 1), (2, 512, 512, 64, 1, 1), (2, 128, 64, 256, 2, 1), (2, 64, 128, 512, 1, 2)]:
                N, I, O, Rin, up, down = case
                g = torch.Generator().manual_seed(11)
                x = torch.randn(N, I, Rin, Rin, generator=g); w = torch.randn(O, I, 3, 3, generator=g) / np.sqrt(9 * I)
                f = R.setup_filter([1, 3, 3, 1])
                dy = torch.randn(N, O, Rin * up // down, Rin * up // down, generator=g)
                xr = x.to(device).requires_grad_(True); wr = w.to(device).requires_grad_(True)
                y = ops.conv2d_resample.conv2d_resample(xr, wr, f=f.to(device), up=up, down=down, padding=1, flip_weight=(up == 1))
                dx, dw = torch.autograd.grad(y, [xr, wr], dy.to(device))
                xo = x.double().requires_grad_(True); wo = w.double().requires_grad_(True)
                yo = R.conv2d_resample(xo, wo, f=f, up=up, down=down, padding=1, flip_weight=(up == 1))
                dxo, dwo = torch.autograd.grad(yo, [xo, wo], dy.double())
                errs = [max_rel_err(y, yo), max_rel_err(dx, dxo), max_rel_err(dw, dwo)]
                assert all(np.isfinite(e) for e in errs)
                lines.append(f'{label:20s} {str(case):32s} y {errs[0]:.2e}  dx {errs[1]:.2e}  dw {errs[2]:.2e}')
    finally:
        co.conv_precision = old
    out = os.path.join(ROOT, 'gpurun_out')
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, 'r2_fast_mode_parity.txt'), 'w') as fh:
        fh.write('max|a-b|/max|b| vs the fp64 CPU oracle, config-f layer shapes (N, I, O, R_in, up, down)\n' + '\n'.join(lines) + '\n')
    print('\n'.join(lines))
