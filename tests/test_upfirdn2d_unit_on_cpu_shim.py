"""The whole translation unit ga-gan_b200/csrc/upfirdn2d.cu through its two C-ABI entry points gg_upfirdn2d_f32 / gg_fir4_pm_f32 --
argument checks (the reference's upfirdn2d.cpp:22-36), the dispatch between the kernel families and every family itself:
`upfirdn2d_generic` (any filter, any factors), `fir4_tile` (five instantiations), `fir_resample2` (true 2x up / down, four pad
phases each), `fir_stream` / `fir_march` (unit rate, plain and phase-major sides) -- unmodified, compiled with g++ against
tests/cuda_cpu_shim.h and executed on the CPU, against the oracle (oracle/ops_ref.py::upfirdn2d, pinned to the live reference), and
under ThreadSanitizer / AddressSanitizer with exact-size tensors (the CPU stand-in for the closed `compute-sanitizer`, DESIGN.md
section 2).  tests/test_fir_stream_on_cpu_shim.py holds the wide unaligned shapes of the staged unit-rate kernel."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import ops_ref as R
from tests import cpu_shim as S

SAN_MAIN = r'''
#include <cstdlib>
static float* tensor(size_t n, float scale) {            // exact-size, 16-byte aligned: the sanitizer's red zone starts behind element n-1
    float* p = (float*)aligned_alloc(16, (n * 4 + 15) / 16 * 16);
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N C inH inW fH fW up down padx0 padx1 pady0 pady1
    int a[12]; for (int i = 0; i < 12; ++i) a[i] = atoi(argv[1 + i]);
    const int N = a[0], C = a[1], H = a[2], W = a[3], fH = a[4], fW = a[5], up = a[6], down = a[7];
    const int outW = (W * up + a[8] + a[9] - fW + down) / down, outH = (H * up + a[10] + a[11] - fH + down) / down;
    float *x = tensor((size_t)N * C * H * W, 1.f), *f = tensor((size_t)fH * fW, .25f), *y = tensor((size_t)N * C * outH * outW, 0.f);
    int rc = gg_upfirdn2d_f32(x, f, y, N, C, H, W, fH, fW, up, up, down, down, a[8], a[9], a[10], a[11], 0, 1.5f, outH, outW, nullptr);
    rc |= gg_upfirdn2d_f32(x, f, y, N, C, H, W, fH, fW, up, up, down, down, a[8], a[9], a[10], a[11], 1, 1.f, outH, outW, nullptr);
    double s = 0; for (size_t i = 0; i < (size_t)N * C * outH * outW; ++i) s += y[i];
    printf("rc %d checksum %.5f blocks %ld\n", rc, s, shim_blocks());
    if (rc) printf("%s\n", shim_error());
    free(x); free(f); free(y);
    return rc;
}
'''


def _source():
    return S.translate_unit(open(os.path.join(S.CSRC, 'upfirdn2d.cu')).read(), expect_launches=8)


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('upfirdn2d_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('upfirdn2d_unit', _source(), 'lib'))
    P, I, F = ctypes.c_void_p, ctypes.c_int, ctypes.c_float
    so.gg_upfirdn2d_f32.restype = I
    so.gg_upfirdn2d_f32.argtypes = [P, P, P] + [I] * 15 + [F, I, I, P]
    so.gg_fir4_pm_f32.restype = I
    so.gg_fir4_pm_f32.argtypes = [P, P, P] + [I] * 7 + [F] + [I] * 8 + [P]
    return so


def _run(lib, xt, f, up, down, pad, flip, gain, skew=0):
    """gg_upfirdn2d_f32 on a copy of xt that sits `skew` floats behind a 16-byte boundary; returns (y, oracle)."""
    N, C, H, W = xt.shape
    want = R.upfirdn2d(xt.double(), f, up=up, down=down, padding=pad, flip_filter=flip, gain=gain).numpy()
    outH, outW = want.shape[2:]
    base, _k = S.aligned(np.zeros(xt.numel() + 4))
    x = base[skew: skew + xt.numel()].reshape(xt.shape)
    x[...] = xt.numpy()
    f2 = f if f.ndim == 2 else f.ger(f)
    fa, _f = S.aligned(f2.numpy())
    y, _y = S.aligned(np.full((N, C, outH, outW), np.nan))
    lib.shim_reset()
    rc = lib.gg_upfirdn2d_f32(x.ctypes.data, fa.ctypes.data, y.ctypes.data, N, C, H, W, f2.shape[0], f2.shape[1], up, up, down, down,
                              pad[0], pad[1], pad[2], pad[3], int(flip), gain, outH, outW, None)
    assert rc == 0, lib.shim_error()
    assert not np.isnan(y).any(), 'some output element was never written'
    return y, want


F4 = [1, 3, 3, 1]

# name, N, C, H, W, filter taps (1-D, made separable) or 'asym', up, down, pad [x0,x1,y0,y1], flip, gain, skew, expected kernel family (threads per CTA)
CASES = [
    ('generic_12tap_aug', 1, 2, 20, 24, [1, 2, 3, 4, 5, 6, 6, 5, 4, 3, 2, 1], 2, 1, [6, 5, 6, 5], False, 4.0, 0, 256),     # ADA's 12-tap form (augment.py:357)
    ('generic_down3', 1, 2, 21, 19, [1, 2, 1], 1, 3, [1, 1, 1, 1], False, 1.0, 0, 256),
    ('generic_under_8_columns', 2, 1, 3, 3, F4, 2, 1, [2, 1, 2, 1], False, 4.0, 0, 256),                                    # 6 output columns: below the marching kernels' floor
    ('generic_negative_pad', 1, 1, 12, 12, F4, 1, 1, [-1, -2, -1, 0], False, 1.0, 0, 256),                                  # crop
    ('generic_asym', 1, 1, 9, 10, 'asym', 1, 1, [1, 1, 0, 2], True, 1.0, 0, 256),                                           # 3 x 5 non-separable filter, flipped
    ('tile_unit_padx0_4', 1, 1, 20, 60, F4, 1, 1, [4, -1, 2, 1], False, 1.0, 0, 256),                                       # padx0 = 4: no marching instantiation -> tile kernel
    ('tile_down2_unaligned', 1, 2, 40, 102, F4, 1, 2, [1, 1, 1, 1], False, 1.0, 0, 256),                                    # rows of 102: not 16-byte -> tile kernel, down 2
    ('tile_up2_even_pad', 1, 1, 26, 25, F4, 2, 1, [2, 1, 2, 1], False, 4.0, 0, 256),                                        # rows of 25 -> tile kernel, up 2, PEX 0
    ('tile_up2_odd_pad', 1, 1, 26, 25, F4, 2, 1, [1, 2, 1, 2], True, 4.0, 0, 256),                                          # PEX 1
    ('tile_up2_skewed_base', 1, 1, 10, 28, F4, 2, 1, [2, 1, 2, 1], False, 4.0, 1, 256),                                     # aligned width, unaligned base pointer
    ('march_down2_p1', 2, 2, 16, 32, F4, 1, 2, [1, 1, 1, 1], False, 1.0, 0, 128),                                           # the discriminator's down path (conv2d_resample.py:119-122)
    ('march_down2_p0', 1, 3, 18, 16, F4, 1, 2, [0, 2, 0, 2], True, 1.0, 0, 128),
    ('march_down2_p2', 1, 1, 9, 36, F4, 1, 2, [2, 0, 2, 0], False, 2.0, 0, 128),
    ('march_down2_p3', 1, 1, 12, 20, F4, 1, 2, [3, 1, 3, 1], False, 1.0, 0, 128),
    ('march_up2_p2', 2, 2, 16, 16, F4, 2, 1, [2, 1, 2, 1], False, 4.0, 0, 128),                                             # upsample2d of the skip path (networks.py:1066)
    ('march_up2_p1', 1, 2, 8, 12, F4, 2, 1, [1, 2, 1, 2], False, 4.0, 0, 128),
    ('march_up2_p3', 1, 1, 10, 8, F4, 2, 1, [3, 0, 3, 0], True, 4.0, 0, 128),
    ('march_up2_p0', 1, 1, 7, 24, F4, 2, 1, [0, 3, 0, 3], False, 4.0, 0, 128),
    ('march_up2_tiny', 4, 3, 4, 4, F4, 2, 1, [2, 1, 2, 1], False, 4.0, 0, 128),                                             # 4^2 -> 8^2: column groups of neighbouring planes fill a warp
    ('stream_unit_narrow', 3, 2, 8, 8, F4, 1, 1, [2, 1, 2, 1], False, 1.0, 0, 128),                                         # 8 columns: the smallest marching width
    ('stream_unit_after_up_conv', 1, 2, 33, 33, F4, 1, 1, [1, 1, 1, 1], False, 4.0, 0, 128),                                # the odd maps behind the transposed conv (conv2d_resample.py:139)
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_upfirdn2d_entry_point_source_on_the_cpu(lib, case):
    name, N, C, H, W, taps, up, down, pad, flip, gain, skew, threads = case
    g = torch.Generator().manual_seed(H * 100 + W)
    xt = torch.randn(N, C, H, W, generator=g)
    if taps == 'asym':
        f = torch.randn(3, 5, generator=g)
    else:
        f = R.setup_filter(taps, gain=1)
    y, want = _run(lib, xt, f, up, down, pad, flip, gain, skew)
    assert np.abs(y - want).max() <= 2e-6 * max(1.0, np.abs(want).max()), name
    assert lib.shim_threads() == threads, f'{name}: dispatched to a kernel family with {lib.shim_threads()} threads per CTA'


def test_upfirdn2d_entry_point_checks(lib):
    x = S.aligned(np.zeros((1, 1, 8, 8)))[0]
    f = S.aligned(np.ones((4, 4)))[0]
    y = S.aligned(np.zeros((1, 1, 16, 16)))[0]
    call = lambda *a: lib.gg_upfirdn2d_f32(x.ctypes.data, f.ctypes.data, y.ctypes.data, *a, None)
    base = dict(N=1, C=1, inH=8, inW=8, fH=4, fW=4, upx=1, upy=1, downx=1, downy=1, px0=1, px1=2, py0=1, py1=2, flip=0, gain=1.0, outH=8, outW=8)
    go = lambda **k: call(*[{**base, **k}[n] for n in base])
    assert go() == 0
    assert go(upx=0) == -1 and b'upsampling' in lib.shim_error()                   # upfirdn2d.cpp:27
    assert go(downy=0) == -1 and b'downsampling' in lib.shim_error()               # :28
    assert go(fH=0) == -1                                                          # :26
    assert go(px0=-9, outW=1) == -1 and b'at least 1x1' in lib.shim_error()        # :34
    assert go(outW=9) == -1 and b'mismatch' in lib.shim_error()
    assert go(N=0) == 0 and go(C=0) == 0                                           # empty batch: no launch
    assert lib.gg_upfirdn2d_f32(None, f.ctypes.data, y.ctypes.data, *[base[n] for n in base], None) == -1
    pm = lambda **k: lib.gg_fir4_pm_f32(x.ctypes.data, f.ctypes.data, y.ctypes.data, *[{**dict(N=1, C=1, inH=8, inW=8, px0=1, py0=1, flip=0, gain=1.0, outH=8, outW=8,
                                                                                                in_pm=0, ipH=0, ipW=0, out_pm=0, opH=0, opW=0), **k}[n]
                                                                                       for n in ('N', 'C', 'inH', 'inW', 'px0', 'py0', 'flip', 'gain', 'outH', 'outW', 'in_pm', 'ipH', 'ipW', 'out_pm', 'opH', 'opW')], None)
    assert pm() == 0
    assert pm(in_pm=1, ipH=4, ipW=4, out_pm=1, opH=4, opW=4) == -1 and b'at most one' in lib.shim_error()
    assert pm(in_pm=1, ipH=3, ipW=4) == -1 and b'smaller' in lib.shim_error()
    assert pm(out_pm=1, opH=4, opW=6) == -1                                         # phase-major output rows in 128-bit groups
    assert pm(px0=4) == -1 and b'padx0' in lib.shim_error()


def _to_pm(t):
    """[N,C,2Y,2X] -> [N,(py,px,c),Y,X] (include/gagan_b200.h: t_pm[n,(py,px,c),Y,X] <-> t[n,c,2Y+py,2X+px])."""
    N, C, H, W = t.shape
    return t.reshape(N, C, H // 2, 2, W // 2, 2).permute(0, 3, 5, 1, 2, 4).reshape(N, 4 * C, H // 2, W // 2)


@pytest.mark.parametrize('N,C,H,W,pad,flip', [(1, 2, 16, 16, [1, 1, 1, 1], False), (2, 1, 10, 24, [2, 2, 2, 2], True), (1, 1, 34, 40, [0, 3, 3, 0], False)],
                         ids=['p1', 'p2-flip', 'p0-p3'])
def test_fir4_phase_major_output_source_on_the_cpu(lib, N, C, H, W, pad, flip):
    """The FIR in front of a stride-2 convolution, writing its result space-to-depth (conv2d_resample.py:119-122 in this build's
    phase-major form): logical positions outside the valid output extent are written as zeros."""
    g = torch.Generator().manual_seed(H + W)
    xt = torch.randn(N, C, H, W, generator=g)
    f = R.setup_filter(F4)
    want = R.upfirdn2d(xt.double(), f, padding=pad, flip_filter=flip, gain=1.0)
    outH, outW = want.shape[2:]
    pmH, pmW = (outH + 1) // 2, ((outW + 1) // 2 + 3) // 4 * 4
    full = torch.zeros(N, C, 2 * pmH, 2 * pmW, dtype=torch.float64)
    full[:, :, :outH, :outW] = want
    want_pm = _to_pm(full).numpy()
    x, _x = S.aligned(xt.numpy()); fa, _f = S.aligned(f.numpy()); y, _y = S.aligned(np.full((N, 4 * C, pmH, pmW), np.nan))
    rc = lib.gg_fir4_pm_f32(x.ctypes.data, fa.ctypes.data, y.ctypes.data, N, C, H, W, pad[0], pad[2], int(flip), 1.0, outH, outW, 0, 0, 0, 1, pmH, pmW, None)
    assert rc == 0, lib.shim_error()
    assert not np.isnan(y).any()
    assert np.abs(y - want_pm).max() <= 2e-6 * max(1.0, np.abs(want_pm).max())


@pytest.mark.parametrize('N,C,pmH,pmW,vH,vW,pad,flip', [(1, 2, 8, 8, 16, 16, [1, 1], False), (2, 1, 9, 12, 17, 23, [2, 2], True), (1, 1, 5, 16, 9, 31, [1, 2], False)],
                         ids=['full', 'odd-valid-extent', 'p1-p2'])
def test_fir4_phase_major_input_source_on_the_cpu(lib, N, C, pmH, pmW, vH, vW, pad, flip):
    """The FIR behind a stride-2 transposed convolution that ran as a stride-1 convolution with 4x the output channels: it reads the
    depth-to-space layout directly (conv2d_resample.py:139); only the valid logical extent (vH, vW) of the phase-major planes counts."""
    g = torch.Generator().manual_seed(pmH * 10 + pmW)
    full = torch.randn(N, C, 2 * pmH, 2 * pmW, generator=g)
    f = R.setup_filter(F4)
    padding = [pad[0], pad[0], pad[1], pad[1]]
    want = R.upfirdn2d(full[:, :, :vH, :vW].double(), f, padding=padding, flip_filter=flip, gain=4.0).numpy()
    outH, outW = want.shape[2:]
    x, _x = S.aligned(_to_pm(full).numpy()); fa, _f = S.aligned(f.numpy()); y, _y = S.aligned(np.full((N, C, outH, outW), np.nan))
    rc = lib.gg_fir4_pm_f32(x.ctypes.data, fa.ctypes.data, y.ctypes.data, N, C, vH, vW, pad[0], pad[1], int(flip), 4.0, outH, outW, 1, pmH, pmW, 0, 0, 0, None)
    assert rc == 0, lib.shim_error()
    assert not np.isnan(y).any()
    assert np.abs(y - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


SAN_CASES = [
    ('generic_12tap', (1, 1, 4, 5, 12, 12, 2, 1, 6, 5, 6, 5)),
    ('generic_crop', (1, 2, 12, 12, 4, 4, 1, 1, -1, -2, -1, 0)),
    ('tile_unit', (1, 1, 9, 60, 4, 4, 1, 1, 4, -1, 2, 1)),
    ('tile_down2', (1, 1, 12, 102, 4, 4, 1, 2, 1, 1, 1, 1)),
    ('tile_up2_even', (1, 1, 7, 25, 4, 4, 2, 1, 2, 1, 2, 1)),
    ('tile_up2_odd', (1, 1, 7, 25, 4, 4, 2, 1, 1, 2, 1, 2)),
    ('march_down2', (2, 1, 16, 32, 4, 4, 1, 2, 1, 1, 1, 1)),
    ('march_up2', (1, 2, 8, 16, 4, 4, 2, 1, 2, 1, 2, 1)),
    ('march_up2_tiny', (2, 3, 4, 4, 4, 4, 2, 1, 2, 1, 2, 1)),
    ('stream_unit', (1, 2, 17, 17, 4, 4, 1, 1, 1, 1, 1, 1)),
]


@pytest.mark.parametrize('kind', ['thread', 'address'])
@pytest.mark.parametrize('case', SAN_CASES, ids=lambda c: c[0])
def test_upfirdn2d_translation_unit_under_sanitizers(kind, case):
    """ThreadSanitizer: the shared-memory input tiles of fir4_tile and the filter taps staged by every family (written, barrier, read)
    are race-free.  AddressSanitizer: with exact-size tensors no family touches a byte outside them -- the zero-padding guards of the
    tile loaders, the 128-bit row loads of the marching kernels at the plane borders, the last partial column group."""
    exe = S.build('upfirdn2d_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, case[1])
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum')
