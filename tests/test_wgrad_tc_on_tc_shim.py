"""The whole translation unit ga-gan_b200/csrc/wgrad_tc.cu -- the tcgen05 weight-gradient kernel that loads its rows with ordinary
global loads (no TMA): the path of operands that are not 16-byte aligned or whose width is not a multiple of 4 (odd maps, cropped
images), same operand rings and accumulator ping-pong as wgrad_tma.cu -- compiled with g++ against tests/tc_cpu_shim.h and executed on
the CPU: against float64 autograd on odd-sized and misaligned tensors, under ThreadSanitizer (racecheck of the X / G rings, whose slot
ownership had the parity alias of DESIGN.md section 6) and AddressSanitizer (exact-size tensors: the row loaders' guards at odd widths)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import cpu_shim as S

EXPORTS = r'''
gg::EncodeTiledFn gg::get_encode_fn() { return &shim_encode_tiled; }
namespace gg {      // gg::wgrad_tc hands 16-byte aligned operands over to wgrad_tma.cu (tests/test_wgrad_tma_on_tc_shim.py); here this unit's own kernel runs on everything
bool wgrad_tma_eligible(const float*, const float*, int, int) { return false; }
int wgrad_tma(const float*, const float*, float*, int, int, int, int, int, int, int, int, int, int, int, int, const float*, const float*, int, int, unsigned, cudaStream_t) { return -9; }
}
extern "C" int tc_wgrad(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int K, int pad_y, int pad_x, int flip_w, int out_layout,
                        const float* as, const float* bs, int nprod, int pm_dim, unsigned pm_dead) {
    if (!gg::wgrad_tc_eligible(N, A, HA, WA, B, HB, WB, K, K, 1, pad_y, pad_x)) return -7;
    return gg::wgrad_tc(a, b, dw, N, A, HA, WA, B, HB, WB, K, K, pad_y, pad_x, flip_w, out_layout, as, bs, nprod, pm_dim, pm_dead, nullptr);
}
'''

SAN_MAIN = r'''
static float* tensor(size_t n, float scale) {            // exact-size and deliberately only 4-byte aligned (+1 float): this kernel's reason to exist
    float* p = (float*)malloc((n + 1) * 4) + 1;
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N A B HA WA K pad_y pad_x HB WB nprod
    int v[11]; for (int i = 0; i < 11; ++i) v[i] = atoi(argv[1 + i]);
    const int N = v[0], A = v[1], B = v[2], HA = v[3], WA = v[4], K = v[5], py = v[6], px = v[7], HB = v[8], WB = v[9], nprod = v[10];
    float *a = tensor((size_t)N * A * HA * WA, 1.f), *b = tensor((size_t)N * B * HB * WB, .7f), *dw = tensor((size_t)A * B * K * K, 0.f), *sa = tensor((size_t)N * A, 1.1f),
          *sb = tensor((size_t)N * B, .9f);
    int rc = tc_wgrad(a, b, dw, N, A, HA, WA, B, HB, WB, K, py, px, 0, 0, nullptr, nullptr, nprod, 0, 0);
    rc |= tc_wgrad(a, b, dw, N, A, HA, WA, B, HB, WB, K, py, px, 1, 1, sa, sb, nprod, 0, 0);
    double s = 0; for (size_t i = 0; i < (size_t)A * B * K * K; ++i) s += dw[i];
    printf("rc %d checksum %.5f mma %ld\n", rc, s, shim_mma_instructions());
    if (rc) printf("%s\n", shim_error());
    free(a - 1); free(b - 1); free(dw - 1); free(sa - 1); free(sb - 1);
    return rc;
}
'''


def _source():
    src = S.translate_tc_unit(open(os.path.join(S.CSRC, 'wgrad_tc.cu')).read(), expect_launches=1)
    return '#define GG_NUM_SMS 3\n' + src + EXPORTS


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('wgrad_tc_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('wgrad_tc_unit', _source(), 'lib'))
    P, I = ctypes.c_void_p, ctypes.c_int
    so.tc_wgrad.restype = I
    so.tc_wgrad.argtypes = [P, P, P] + [I] * 12 + [P, P, I, I, ctypes.c_uint]
    so.shim_mma_instructions.restype = ctypes.c_long
    return so


def _skewed(t, skew):
    base = S.aligned(np.zeros(t.numel() + 4))[0]
    v = base[skew: skew + t.numel()].reshape(t.shape)
    v[...] = t.numpy()
    return v


def _p(a):
    return None if a is None else a.ctypes.data


# name, N, A, B, HA, WA, K, pad, flip, out_layout, scales, nprod, skew of the operands in floats
CASES = [
    ('odd_map_3x3', 1, 16, 16, 17, 17, 3, 1, 0, 0, False, 3, 0),
    ('odd_map_modulated_flip', 2, 32, 48, 19, 22, 3, 1, 1, 0, True, 3, 1),
    ('layout_ab_ragged_tiles', 2, 40, 100, 13, 18, 3, 1, 0, 1, True, 3, 3),
    ('1x1', 2, 16, 32, 15, 15, 1, 0, 0, 0, False, 3, 2),
    ('2x2_odd_tasks', 2, 64, 32, 17, 17, 2, 1, 0, 0, True, 3, 1),                    # the ring-ownership regression shape
    ('one_product', 2, 32, 32, 17, 19, 2, 1, 0, 0, False, 1, 0),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_wgrad_tc_source_on_the_hardware_model(lib, case):
    name, N, A, B, HA, WA, K, pad, flip, layout, scales, nprod, skew = case
    g = torch.Generator().manual_seed(len(name) * 13 + A)
    a = torch.randn(N, A, HA, WA, generator=g)
    HB, WB = HA + 2 * pad - K + 1, WA + 2 * pad - K + 1
    b = torch.randn(N, B, HB, WB, generator=g)
    sa = torch.randn(N, A, generator=g) if scales else None
    sb = torch.randn(N, B, generator=g) if scales else None
    ad = a.double() * (sa.double()[:, :, None, None] if scales else 1)
    bd = b.double() * (sb.double()[:, :, None, None] if scales else 1)
    wv = torch.zeros(B, A, K, K, dtype=torch.float64, requires_grad=True)
    (F.conv2d(ad, wv, padding=pad) * bd).sum().backward()
    want = wv.grad
    if flip:
        want = want.flip([2, 3])
    if layout:
        want = want.transpose(0, 1)
    want = want.contiguous().numpy()
    as_, bs_ = _skewed(a, skew), _skewed(b, skew)
    sas = S.aligned(sa.numpy())[0] if scales else None
    sbs = S.aligned(sb.numpy())[0] if scales else None
    dw = S.aligned(np.full(want.shape, np.nan))[0]
    before = lib.shim_mma_instructions()
    assert lib.tc_wgrad(_p(as_), _p(bs_), _p(dw), N, A, HA, WA, B, HB, WB, K, pad, pad, flip, layout, _p(sas), _p(sbs), nprod, 0, 0) == 0, lib.shim_error()
    scale = float(((ad ** 2).sum() * (bd ** 2).sum() / (A * B)).sqrt())
    assert np.abs(dw - want).max() <= (5e-6 if nprod == 3 else 2e-3) * max(scale, np.abs(want).max()), name
    assert lib.shim_mma_instructions() > before


def test_wgrad_tc_eligibility_source(lib):
    a = S.aligned(np.zeros((1, 16, 16, 16)))[0]
    dw = S.aligned(np.zeros((16, 16, 3, 3)))[0]
    go = lambda **k: lib.tc_wgrad(_p(a), _p(a), _p(dw), *[{**dict(N=1, A=16, HA=16, WA=16, B=16, HB=16, WB=16, K=3, py=1, px=1), **k}[n]
                                                         for n in ('N', 'A', 'HA', 'WA', 'B', 'HB', 'WB', 'K', 'py', 'px')], 0, 0, None, None, 3, 0, 0)
    assert go() == 0
    assert go(A=3, B=3) == -7 and go(HB=4, WB=4, HA=4, WA=4) == -7 and go(K=4) == -7 and go(px=3) == -7          # -> the exact FFMA kernel


SAN_CASES = [('odd_3x3', (2, 32, 48, 19, 21, 3, 1, 1, 19, 21, 3)), ('2x2_odd_tasks', (2, 64, 32, 17, 17, 2, 1, 1, 18, 18, 3)), ('1x1_one_product', (2, 16, 32, 15, 15, 1, 0, 0, 15, 15, 1))]


@pytest.mark.parametrize('kind', ['thread', 'address'])
@pytest.mark.parametrize('case', SAN_CASES, ids=lambda c: c[0])
def test_wgrad_tc_pipeline_under_sanitizers(kind, case):
    exe = S.build('wgrad_tc_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, case[1], timeout=1500)
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum')


def test_the_racecheck_does_report_a_broken_global_load_weight_gradient():
    old = 'if (has_g) mbar_wait_spin(BAR_G_EMPTY(gslot), ((t.gc / (GS / 2)) & 1) ^ 1);'
    src = _source()
    assert src.count(old) == 1
    exe = S.build('wgrad_tc_mutant', src.replace(old, ''), 'thread', SAN_MAIN)
    reported, out = S.mutant_is_reported(exe, SAN_CASES[1][1])
    if reported is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert reported, out[-2000:]
