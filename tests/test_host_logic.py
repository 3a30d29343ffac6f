"""CPU: host-side logic of the product -- integer pad/branch bookkeeping (bit-exact against the oracle),
the C-ABI library (loads, exports every symbol the header declares), and the no-fallback contract."""
import os
import re
import ctypes
import itertools
import pytest
import torch

from tests.util import ROOT, PKG
from oracle import ops_ref as R

from tests.util import CHECKOUT, HAVE_CHECKOUT, reference_networks, quiet

from torch_utils import custom_ops                                                      # after install(): this build's modules
from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample, conv2d_gradfix, fma


def test_conv2d_resample_plan_bit_exact():
    f4 = R.setup_filter([1, 3, 3, 1]); f12 = R.setup_filter(list(range(1, 13))); f3 = R.setup_filter([1, 2, 1])
    n = 0
    for k, up, down, f, pad in itertools.product([1, 3, 5], [1, 2, 4], [1, 2, 3], [None, f4, f12, f3],
                                                 [0, 1, 2, [1, 0], [0, 1, 2, 3], [-1, 1, 0, 2]]):
        got = conv2d_resample.plan((8, 4, k, k), f, up, down, pad)
        want = R.conv2d_resample_plan((8, 4, k, k), f, up, down, pad)
        assert got == want, (k, up, down, pad, got, want)
        n += 1
    assert n == 3 * 3 * 3 * 4 * 6


def test_padding_helpers_and_filter():
    for pad in [0, 3, [1, 2], [1, 2, 3, 4], [-1, 0, 2, -3]]:
        assert upfirdn2d._parse_padding(pad) == R.parse_padding(pad)
    for s in [1, 3, [2, 1]]:
        assert upfirdn2d._parse_scaling(s) == R.parse_scaling(s)
    for spec, kw in [([1, 3, 3, 1], {}), ([1, 2, 1], dict(gain=4)), (list(range(1, 13)), {}), ([1, 3, 3, 1], dict(flip_filter=True, normalize=False)),
                     (None, {}), (2.0, {}), ([1, 3, 3, 1], dict(separable=True))]:
        a = upfirdn2d.setup_filter(spec, **kw); b = R.setup_filter(spec, **kw)
        assert a.shape == b.shape and torch.equal(a, b)
    assert upfirdn2d._get_filter_size(None) == (1, 1)
    assert upfirdn2d._get_filter_size(torch.zeros(3, 5)) == (5, 3)


def test_activation_table_matches_reference_values():
    for name, (alpha, gain, idx, ref, has2) in R.ACTIVATIONS.items():
        s = bias_act.activation_funcs[name]
        assert (s.def_alpha, float(s.def_gain), s.cuda_idx, s.ref, s.has_2nd_grad) == (alpha, float(gain), idx, ref, has2)


def test_cabi_library_loads_and_exports_header_symbols():
    header = open(os.path.join(ROOT, 'include', 'gagan_b200.h')).read()
    declared = re.findall(r'GG_API\s+[\w\s\*]+?\b(gg_\w+)\s*\(', header)
    assert len(declared) >= 8
    lib = custom_ops.load_library()                     # builds with nvcc if the .so is missing; no GPU needed
    for sym in declared:
        assert hasattr(lib, sym), f'{sym} declared in include/gagan_b200.h but not exported by libgagan_b200.so'
    assert set(declared) == set(custom_ops.EXPORTED_SYMBOLS)
    assert lib.gg_version() == 100
    assert isinstance(lib.gg_launch_count(), int)
    # argument validation runs before any CUDA call: usable without a device
    rc = lib.gg_bias_act_f32(None, None, None, None, None, None, None, 0, 3, 0.2, 1.0, -1.0, 16, 1, 1, None)
    assert rc == -1 and b'non-null' in lib.gg_last_error()
    rc = lib.gg_upfirdn2d_f32(ctypes.c_void_p(16), ctypes.c_void_p(16), ctypes.c_void_p(16), 1, 1, 4, 4, 4, 4, 0, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 1, 1, None)
    assert rc == -1 and b'upsampling factor' in lib.gg_last_error()


def test_cabi_argument_validation_precedes_any_cuda_call():
    """Error behaviour at the boundary (TORCH_CHECK -> RuntimeError becomes a negative code + gg_last_error()):
    every entry point validates its arguments before it touches the device, so the checks run on a box without a GPU.  The
    32-bit indexing limit of the reference (upfirdn2d.cpp:34,36, bias_act.cpp:40) is among them."""
    lib = custom_ops.load_library()
    P = ctypes.c_void_p(4096)                            # never dereferenced: every call below must fail validation first
    big = 1 << 16

    def expect(rc, needle):
        msg = lib.gg_last_error() or b''
        assert rc < 0 and needle in msg, (rc, msg)

    expect(lib.gg_bias_act_f32(P, None, None, None, None, P, None, 0, 3, 0.2, 1.0, -1.0, 1 << 31, 1, 1, None), b'too large')
    expect(lib.gg_bias_act_f32(P, None, None, None, None, P, None, 3, 3, 0.2, 1.0, -1.0, 16, 1, 1, None), b'grad must be')
    expect(lib.gg_bias_act_f32(P, None, None, None, None, P, None, 0, 11, 0.2, 1.0, -1.0, 16, 1, 1, None), b'activation')
    expect(lib.gg_upfirdn2d_f32(P, P, P, big, big, 4, 4, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 1, 1, None), b'too large')
    expect(lib.gg_upfirdn2d_f32(P, P, P, 1, 1, 8, 8, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 6, 5, None), b'output size mismatch')
    expect(lib.gg_upfirdn2d_f32(P, P, P, 1, 1, 2, 2, 4, 4, 1, 1, 1, 1, 0, 0, 0, 0, 0, 1.0, 1, 1, None), b'at least 1x1')
    expect(lib.gg_upfirdn2d_f32(P, P, P, 1, 1, 8, 8, 4, 4, 1, 1, 0, 1, 0, 0, 0, 0, 0, 1.0, 5, 5, None), b'downsampling factor')
    used = ctypes.c_int(0)
    expect(lib.gg_conv2d_f32(P, P, P, big, 512, 64, 64, 512, 3, 3, 64, 64, 1, 1, 1, 0, 0, None, None, -1, ctypes.byref(used), None), b'too large')
    expect(lib.gg_conv2d_f32(P, P, P, 1, 8, 16, 16, 8, 3, 3, 17, 16, 2, 1, 1, 0, 0, None, None, -1, ctypes.byref(used), None), b'output size mismatch')
    expect(lib.gg_conv2d_f32(P, P, P, 1, 8, 16, 16, 8, 3, 3, 16, 16, 1, -1, 1, 0, 0, None, None, -1, ctypes.byref(used), None), b'stride/padding')
    expect(lib.gg_conv2d_f32(P, P, P, 1, 8, 16, 16, 8, 3, 3, 16, 16, 1, 1, 1, 0, 0, None, None, 77, ctypes.byref(used), None), b'prec')
    expect(lib.gg_conv2d_wgrad_f32(P, P, P, 1, 8, 16, 16, 8, 16, 16, 0, 3, 1, 1, 1, 0, 0, None, None, -1, ctypes.byref(used), None), b'bad shape')
    expect(lib.gg_conv2d_wgrad_f32(P, P, P, big, 512, 64, 64, 8, 16, 16, 3, 3, 1, 1, 1, 0, 0, None, None, -1, ctypes.byref(used), None), b'too large')
    expect(lib.gg_fma_rows_f32(P, P, P, 7, P, 15, 5, 96, None), b'one [P] plane')
    expect(lib.gg_fma_rows_f32(P, P, P, 0, P, 14, 5, 96, None), b'bad extents')
    expect(lib.gg_scale_rows_f32(P, None, P, 4, 4, None), b'null pointer')
    assert lib.gg_fma_rows_f32(P, P, P, 0, P, 0, 5, 96, None) == 0          # empty tensors are a no-op, not an error


def test_no_cpu_fallback_and_no_ref_impl():
    x = torch.randn(2, 3, 8, 8)
    f = upfirdn2d.setup_filter([1, 3, 3, 1])
    with pytest.raises(RuntimeError, match='no CPU path'):
        bias_act.bias_act(x, torch.zeros(3))
    with pytest.raises(RuntimeError, match='no CPU path'):
        upfirdn2d.upfirdn2d(x, f)
    with pytest.raises(RuntimeError, match='no CPU path'):
        conv2d_gradfix.conv2d(x, torch.randn(4, 3, 3, 3), padding=1)
    with pytest.raises(RuntimeError, match="impl='ref'"):
        bias_act.bias_act(x, impl='ref')
    with pytest.raises(RuntimeError, match="impl='ref'"):
        upfirdn2d.upsample2d(x, f, impl='ref')
    with pytest.raises(RuntimeError):
        custom_ops.get_plugin('no_such_plugin')
    # the product never imports the oracle
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert 'import oracle' not in src and 'from oracle' not in src, fn


def test_fma_matches_oracle_on_cpu():
    # fma is plain torch (no kernel): check forward and the hand-written broadcast-aware backward
    g = torch.Generator().manual_seed(1)
    a = torch.randn(2, 4, 5, 5, generator=g, requires_grad=True)
    b = torch.randn(2, 4, 1, 1, generator=g, requires_grad=True)
    c = torch.randn(2, 1, 5, 5, generator=g, requires_grad=True)
    y = fma.fma(a, b, c)
    yo = R.fma(a, b, c)
    assert torch.allclose(y, yo)
    dy = torch.randn(y.shape, generator=g)
    for u, v in zip(torch.autograd.grad(y, [a, b, c], dy), torch.autograd.grad(yo, [a, b, c], dy)):
        assert torch.allclose(u, v, atol=1e-6)


def test_install_binds_the_operator_modules_into_the_reference_checkout():
    """The drop-in: after gagan_b200.install(checkout) the reference's own modules resolve their operators to this build,
    nothing else of the checkout is shadowed, and its module tree is the reference's code (not a copy shipped here)."""
    import sys
    import gagan_b200
    networks = reference_networks()
    for name in gagan_b200.OPS:
        mod = sys.modules['torch_utils.ops.' + name]
        assert mod.__name__ == f'gagan_b200.torch_utils.ops.{name}', mod
    assert sys.modules['torch_utils.custom_ops'].__name__ == 'gagan_b200.torch_utils.custom_ops'
    for attr in ('conv2d_resample', 'upfirdn2d', 'bias_act', 'fma'):                     # networks.py:16-19
        assert getattr(networks, attr).__name__.startswith('gagan_b200.'), attr
    assert networks.modulated_conv2d.__module__ == 'gagan_b200.training.networks'
    assert os.path.abspath(networks.__file__).startswith(CHECKOUT)
    from training import loss, augment                                                    # loss.py:13, augment.py:14-16
    assert loss.conv2d_gradfix.__name__.startswith('gagan_b200.') and augment.upfirdn2d.__name__.startswith('gagan_b200.')
    assert augment.conv2d_gradfix.__name__.startswith('gagan_b200.')
    from torch_utils import misc, persistence, training_stats                             # the checkout's own, not shadowed
    for mod in (misc, persistence, training_stats, loss, augment):
        assert os.path.abspath(mod.__file__).startswith(CHECKOUT), mod
    # nothing under ga-gan_b200/ defines the reference's module tree or loss
    for dirpath, _, files in os.walk(PKG):
        for fn in files:
            if fn.endswith('.py'):
                src = open(os.path.join(dirpath, fn)).read()
                assert 'class Generator' not in src and 'class StyleGAN2Loss' not in src and 'class SynthesisBlock' not in src, fn


def test_similar_domains_consumers_bind_to_this_build():
    """The second consumer of the module-path boundary: SimilarDomains/gan_models/StyleGAN2/nvidia.py imports `torch_utils.ops.*`
    (nvidia.py:16-21) and needs nothing beyond install(); the rosinality module of the same tree is bound by install_rosinality."""
    import importlib
    import gagan_b200
    from tests.util import SD_CHECKOUT, rosinality_model
    reference_networks()
    model = rosinality_model()
    nvidia = importlib.import_module('gan_models.StyleGAN2.nvidia')
    assert os.path.abspath(nvidia.__file__).startswith(SD_CHECKOUT)
    for attr in ('conv2d_resample', 'upfirdn2d', 'bias_act', 'fma'):
        assert getattr(nvidia, attr).__name__.startswith('gagan_b200.'), attr
    assert model.upfirdn2d.__module__ == 'gagan_b200.rosinality' and model.ModulatedConv2d.forward.__module__ == 'gagan_b200.rosinality'
    st = model._gagan_b200_rosinality
    gagan_b200.install_rosinality(model, fused_layers=False)                 # the module's own layer code on the replaced pieces
    assert model.StyledConv.forward is st['styled_conv_forward'] and model.ConvLayer.forward is st['conv_layer_forward']
    gagan_b200.install_rosinality(model, fused_layers=True)
    assert model.StyledConv.forward.__module__ == 'gagan_b200.rosinality'
    with pytest.raises(RuntimeError):                                        # no CPU path behind the adapter either
        model.Upsample([1, 3, 3, 1])(torch.zeros(1, 1, 4, 4))


def test_reference_networks_build_with_the_golden_state_dict_names():
    from tests.util import load_golden
    networks = reference_networks()
    g = load_golden('networks')
    cfg = {kv.split('=')[0]: int(kv.split('=')[1]) for kv in (str(m) for m in g['meta'])}
    G = networks.Generator(z_dim=cfg['z_dim'], c_dim=0, w_dim=cfg['w_dim'], img_resolution=cfg['res'], img_channels=3,
                           mapping_kwargs=dict(num_layers=cfg['num_layers']),
                           synthesis_kwargs=dict(channel_base=cfg['channel_base'], channel_max=cfg['channel_max']))
    D = networks.Discriminator(c_dim=0, img_resolution=cfg['res'], img_channels=3, channel_base=cfg['channel_base'],
                               channel_max=cfg['channel_max'], epilogue_kwargs=dict(mbstd_group_size=cfg['mbstd']))
    for net, pre in ((G, 'G.'), (D, 'D.')):
        sd = {k: v for k, v in net.state_dict().items() if not k.endswith('resample_filter')}
        gold = {k[2:]: v for k, v in g.items() if k.startswith(pre)}
        assert set(sd) == set(gold), (set(sd) ^ set(gold))
    # parameter counts of the real configs: cfg-f 1024^2 G 30.37 M
    Gf = networks.Generator(512, 0, 512, 1024, 3, mapping_kwargs=dict(num_layers=8), synthesis_kwargs=dict(channel_base=32768))
    assert sum(p.numel() for p in Gf.parameters()) == 30370060 and Gf.num_ws == 18


def _affine_plus_generator(networks):
    return quiet(networks.Generator, z_dim=16, c_dim=0, w_dim=16, img_resolution=64, img_channels=3, mapping_kwargs=dict(num_layers=2),
                 synthesis_kwargs=dict(channel_base=1024, channel_max=16, use_domain_modulation=True,
                                       domain_modulation_parametrization='out_in_5_1,additive,affine_out_in_5_1',
                                       generator_requires_grad_parts=['all']))


def test_trainable_part_filter_matches_the_golden_selection():
    """training_loop.select_parts == the reference's name_filters / set_requires_grad (training_loop.py:57-95); the golden
    selections were produced by the reference function itself (tests/golden/make_parts_golden.py)."""
    import json
    from gagan_b200.training import training_loop
    networks = reference_networks()
    gold = json.load(open(os.path.join(ROOT, 'tests', 'golden', 'requires_grad_parts.json')))
    G = _affine_plus_generator(networks)
    assert [n for n, _ in G.named_parameters()] == gold['parameter_names']
    assert len(gold['specs']) >= 10
    for spec, want in zip(gold['specs'], gold['selected']):
        assert training_loop.select_parts(G, spec) == want, spec
        training_loop.set_requires_grad(G, spec)
        assert [n for n, p in G.named_parameters() if p.requires_grad] == want


def test_phase_major_forms_of_the_stride2_layers_match_the_oracle():
    """Host algebra of conv2d_resample's up=2 / down=2 paths: with torch stand-ins for the two device ops, the
    phase-major (space-to-depth) formulation must reproduce the oracle's conv2d_resample (conv2d_resample.py:119-142)."""
    import torch
    import torch.nn.functional as F
    from torch_utils.ops import conv2d_resample as cr
    from oracle import ops_ref as R

    def conv_s1(x, w, padding, out_hw, live):
        # y[Y,X] = sum x[Y-py+a, X-px+b] w[a,b], zero outside x, free output extent
        kh, kw = w.shape[2:]
        py, px = padding
        H, W = x.shape[2:]
        need_h, need_w = out_hw[0] + kh - 1, out_hw[1] + kw - 1
        xp = F.pad(x, (px, max(need_w - W - px, 0), py, max(need_h - H - py, 0)))[:, :, :need_h, :need_w]
        return F.conv2d(xp, w)

    def fir_to_pm(x, f, padding, flip_filter, gain, ys, xs):
        return cr.space_to_depth(R.upfirdn2d(x, f, padding=padding, flip_filter=flip_filter, gain=gain), ys, xs)

    def fir_from_pm(z, f, padding, flip_filter, gain, valid_hw):
        full = cr.depth_to_space(z)[:, :, :valid_hw[0], :valid_hw[1]]
        return R.upfirdn2d(full, f, padding=padding, flip_filter=flip_filter, gain=gain)

    g = torch.Generator().manual_seed(11)
    f = R.setup_filter([1, 3, 3, 1])
    for (N, I, O, H, W, flip_weight) in [(2, 5, 7, 8, 8, True), (1, 4, 6, 16, 12, False), (2, 3, 4, 5, 9, True)]:
        x = torch.randn(N, I, H, W, generator=g, dtype=torch.float64)
        w = torch.randn(O, I, 3, 3, generator=g, dtype=torch.float64)
        for up, down in [(2, 1), (1, 2)]:
            want = R.conv2d_resample(x, w, f=f, up=up, down=down, padding=1, flip_weight=flip_weight)
            pl = cr.plan(w.shape, f, up, down, 1)
            if up == 2:
                got = cr.up2_phase_major(x, w, f, pl['fir_pad'], flip_weight, False, conv_s1, fir_from_pm)
            else:
                got = cr.down2_phase_major(x, w, f, pl['fir_pad'], flip_weight, False, conv_s1, fir_to_pm)
            assert got.shape == want.shape, (got.shape, want.shape)
            assert float((got - want).abs().max()) < 1e-12, (up, down, flip_weight)
    # 7 of the 16 (phase, tap) blocks of the phase-major weights are structurally zero
    w2 = cr.phase_major_weight_down(torch.ones(2, 3, 3, 3)).reshape(2, 4, 3, 4)
    assert int((w2.abs().sum(dim=(0, 2)) == 0).sum()) == 7
    w2 = cr.phase_major_weight_up(torch.ones(2, 3, 3, 3)).reshape(4, 2, 3, 4)
    assert int((w2.abs().sum(dim=(1, 2)) == 0).sum()) == 7


@pytest.mark.parametrize('kind', ['down', 'up'])
@pytest.mark.parametrize('k', [1, 2, 3, 4])
def test_phase_major_dead_tap_mask_is_the_zero_pattern_of_the_weight(kind, k):
    """The structural hint handed to gg_conv2d_wgrad_pm_f32 (conv2d_resample._pm_live) must mark exactly the (phase group,
    tap) blocks that phase_major_weight_down / _up leave zero for ANY weight -- otherwise the kernel would skip live gradients."""
    O, I = 3, 5
    w = torch.rand(O, I, k, k) + 1.0                                   # no accidental zeros
    w2 = (conv2d_resample.phase_major_weight_down if kind == 'down' else conv2d_resample.phase_major_weight_up)(w)
    live = conv2d_resample._pm_live(kind, k, k)
    pm_dim, dead = live.pm
    assert pm_dim == (2 if kind == 'down' else 1)
    grouped = w2.reshape(O, 4, I, 2, 2) if kind == 'down' else w2.reshape(4, O, I, 2, 2).transpose(0, 1)   # [O, group, I, a, b]
    n_dead = 0
    for g, a, b in itertools.product(range(4), range(2), range(2)):
        block = grouped[:, g, :, a, b]
        is_dead = bool((dead >> (g * 4 + a * 2 + b)) & 1)
        assert bool((block == 0).all()) == is_dead, (kind, k, g, a, b)
        assert is_dead or bool((block != 0).all())
        n_dead += is_dead
    if k == 3:
        assert n_dead == 7 and abs(float(live) - 9.0 / 16.0) < 1e-12    # 9 of the 16 (phase, tap) blocks are live


def test_tcgen05_kernels_keep_their_registers():
    """The consumer warps of the tcgen05 kernels run at the 168-register cap that setmaxnreg gives them; a few bytes of spill stack in
    those loops cost 11 % of the in-step convolution throughput once (round 2: a fatter watchdog expiry path inlined at every mbarrier
    wait).  Every default instantiation must have a zero stack frame (the fused-epilogue <128,2> one is the known exception, off by
    default).  Reads the objects `make` left in csrc/build with cuobjdump: no GPU needed."""
    import shutil
    import subprocess
    build = os.path.join(PKG, 'csrc', 'build')
    cuobjdump = shutil.which('cuobjdump') or '/usr/local/cuda/bin/cuobjdump'
    objs = [os.path.join(build, n + '.o') for n in ('conv_tc', 'conv_march', 'wgrad_tma', 'wgrad_tc')]
    if not os.path.isfile(cuobjdump) or not all(os.path.isfile(o) for o in objs):
        pytest.skip('cuobjdump or the object files are not available (run __graft_entry__.build() first)')
    seen = 0
    for obj in objs:
        out = subprocess.run([cuobjdump, '--dump-resource-usage', obj], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
        for name, regs, stack in re.findall(r'Function (\S+):\s*\n\s*REG:(\d+) STACK:(\d+)', out):
            if not re.search(r'conv_tc_kernel|conv_march_kernel|wgrad_tma_kernel|wgrad_tc_kernel', name):
                continue
            seen += 1
            assert int(regs) <= 168, (name, regs)
            if 'conv_tc_kernelILi128ELi2ELb1' in name:
                continue
            assert int(stack) == 0, f'{name} has a {stack}-byte stack frame: something in its loops spills'
    assert seen >= 12, seen


@pytest.mark.parametrize('K,rows_per_strip', [(2, [16] * 5), (2, [16, 16, 1, 16, 16, 1]), (3, [16] * 4), (3, [16, 16, 1, 16, 7]), (1, [16, 3, 16])])
def test_weight_gradient_ring_indexing_keeps_one_owner_per_slot(K, rows_per_strip):
    """Model of the ring arithmetic of wgrad_tma.cu / wgrad_tc.cu (the converter groups and the MMA issuer derive slot and mbarrier
    phase from running counters): every X slot and every G slot must be filled by ONE converter group for the whole kernel, each group
    must meet the uses of its slots in order, and the issuer must look for every row where its converter put it.  The round-1 / round-2
    kernels dealt tasks out by their index inside a strip; with an odd number of tasks (K = 2) or rows (16 n + 1) per strip slot
    ownership flipped between the groups, a parity wait could pass one release early, and launches failed now and then."""
    XS, GS = 8, 4
    conv = {0: [], 1: []}                                   # per group: (kind, slot, use) in program order
    x_owner, g_owner = {}, {}
    tbase, gn = 0, [0, 0]
    g_of_row = []                                           # converter view: (slot, phase) of every G row in image order
    for rows in rows_per_strip:
        ntask = rows + K - 1
        placed = {}
        for grp in (0, 1):
            for j in range((tbase ^ grp) & 1, ntask, 2):
                pos = tbase + j
                xslot = pos % XS
                assert x_owner.setdefault(xslot, grp) == grp, 'an X slot changed hands'
                conv[grp].append(('x', xslot, pos // XS))
                if j >= K - 1:
                    gslot, phase = grp + 2 * (gn[grp] % (GS // 2)), (gn[grp] // (GS // 2)) & 1
                    assert g_owner.setdefault(gslot, grp) == grp, 'a G slot changed hands'
                    conv[grp].append(('g', gslot, gn[grp] // (GS // 2)))
                    placed[j - (K - 1)] = (gslot, phase)
                    gn[grp] += 1
        g_of_row += [placed[i] for i in range(rows)]
        tbase += ntask
    # a group's uses of one slot come in order 0, 1, 2, ... (so "release #u-1" is always awaited after "release #u-2" was seen)
    for grp, seq in conv.items():
        last = {}
        for kind, slot, use in seq:
            assert use == last.get((kind, slot), -1) + 1, (grp, kind, slot, use)
            last[(kind, slot)] = use
    # the issuer: one row at a time, group = parity of the row's task position, per-group counters
    xq, cnt, seen = 0, [0, 0], []
    for rows in rows_per_strip:
        for _ in range(rows):
            og = (xq + K - 1) & 1
            seen.append((og + 2 * (cnt[og] % (GS // 2)), (cnt[og] // (GS // 2)) & 1))
            cnt[og] += 1
            xq += 1
        xq += K - 1
    assert seen == g_of_row
