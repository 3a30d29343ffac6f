"""The whole translation unit ga-gan_b200/csrc/wgrad_tma.cu -- the tcgen05 weight-gradient kernel (22 % of the training step): row
tasks staged by two TMA producers, converted by two groups of warps into an X operand ring in shared memory and a G operand ring in
TENSOR MEMORY (tcgen05.st), multiplied by the .ts form of tcgen05.mma, drained and flushed with fp32 atomics; and its host code
(work partition, tensor maps) -- compiled with g++ against tests/tc_cpu_shim.h and executed on the CPU.

This is the kernel whose operand rings had the mbarrier parity alias of DESIGN.md section 6 (ownership of the ring slots flipped
between the converter groups whenever a strip had an odd number of tasks: 2x2 taps, heights 16 n + 1).  `compute-sanitizer` is closed
on the GPU pool; here the shipped source runs those shapes under ThreadSanitizer (mbarriers are acquire / release atomics in the model,
TMA copies and MMAs run in the issuing thread: an operand overwritten before its release is a reported data race, an over-arrival
aborts the model) and under AddressSanitizer with exact-size tensors and shared memory, and agrees with float64 autograd."""
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
extern "C" int tc_wgrad(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int K, int pad_y, int pad_x, int flip_w, int out_layout,
                        const float* as, const float* bs, int nprod, int pm_dim, unsigned pm_dead) {
    if (!gg::wgrad_tma_eligible(a, b, WA, WB)) return -7;
    return gg::wgrad_tma(a, b, dw, N, A, HA, WA, B, HB, WB, K, pad_y, pad_x, flip_w, out_layout, as, bs, nprod, pm_dim, pm_dead, nullptr);
}
'''

SAN_MAIN = r'''
static float* tensor(size_t n, float scale) {            // exact-size, 16-byte aligned: the sanitizer's red zone starts behind element n-1
    float* p = (float*)aligned_alloc(16, (n * 4 + 15) / 16 * 16);
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
    free(a); free(b); free(dw); free(sa); free(sb);
    return rc;
}
'''


def _source():
    return '#define GG_NUM_SMS 3\n' + S.translate_tc_unit(open(os.path.join(S.CSRC, 'wgrad_tma.cu')).read(), expect_launches=1) + EXPORTS


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('wgrad_tma_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('wgrad_tma_unit', _source(), 'lib'))
    P, I = ctypes.c_void_p, ctypes.c_int
    so.tc_wgrad.restype = I
    so.tc_wgrad.argtypes = [P, P, P] + [I] * 12 + [P, P, I, I, ctypes.c_uint]
    so.shim_mma_instructions.restype = ctypes.c_long
    return so


def _a(t):
    return None if t is None else S.aligned(t.numpy())[0]


def _p(a):
    return None if a is None else a.ctypes.data


# name, N, A, B, HA, WA, K, (pad_y, pad_x), (HB, WB) or None = natural, flip, out_layout, scales, nprod
CASES = [
    ('3x3_one_tile', 1, 16, 16, 16, 16, 3, (1, 1), None, 0, 0, False, 3),
    ('3x3_modulated_flip', 2, 32, 48, 20, 24, 3, (1, 1), None, 1, 0, True, 3),                   # two row strips (16 + 4), RB = 64
    ('3x3_layout_ab_128', 1, 40, 100, 12, 16, 3, (1, 1), None, 0, 1, True, 3),                   # NTA = 64 with a ragged A tile, RB = 128, dw as [A,B,k,k]
    ('1x1', 1, 16, 32, 8, 8, 1, (0, 0), None, 0, 0, False, 3),
    ('2x2_odd_tasks', 2, 64, 32, 16, 16, 2, (1, 1), (17, 20), 0, 0, True, 3),                    # K = 2, 17 gradient rows: the ring-ownership regression shape
    ('2x2_up_layer_form', 1, 32, 64, 33, 36, 2, (0, 0), (32, 36), 0, 1, True, 3),                # heights 16 n + 1 on the input side
    ('free_extent_pad2', 1, 16, 32, 10, 12, 3, (2, 0), (12, 12), 0, 0, False, 3),                # gradient larger than the natural correlation output
    ('wide_rows_two_column_strips', 1, 16, 16, 6, 136, 3, (1, 1), None, 0, 0, False, 3),         # WA > one 128-column strip
    ('one_product', 2, 32, 32, 17, 16, 2, (1, 1), (18, 20), 0, 0, False, 1),                     # the mode in which the old alias surfaced within a few iterations
]


def _want(a, b, sa, sb, K, py, px, flip, layout):
    ad = a.double() * (sa.double()[:, :, None, None] if sa is not None else 1)
    bd = b.double() * (sb.double()[:, :, None, None] if sb is not None else 1)
    HB, WB = b.shape[2:]
    HA, WA = a.shape[2:]
    ap = F.pad(ad, [px, max(WB + K - 1 - WA - px, 0), py, max(HB + K - 1 - HA - py, 0)])[:, :, :HB + K - 1, :WB + K - 1]
    wv = torch.zeros(b.shape[1], a.shape[1], K, K, dtype=torch.float64, requires_grad=True)
    (F.conv2d(ap, wv) * bd).sum().backward()
    g = wv.grad
    if flip:
        g = g.flip([2, 3])
    if layout:
        g = g.transpose(0, 1)
    return g.contiguous().numpy(), float(((ad ** 2).sum() * (bd ** 2).sum() / (a.shape[1] * b.shape[1])).sqrt())


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_wgrad_tma_source_on_the_hardware_model(lib, case):
    name, N, A, B, HA, WA, K, (py, px), ext, flip, layout, scales, nprod = case
    g = torch.Generator().manual_seed(len(name) * 11 + A)
    a = torch.randn(N, A, HA, WA, generator=g)
    HB, WB = ext if ext else (HA + 2 * py - K + 1, WA + 2 * px - K + 1)
    b = torch.randn(N, B, HB, WB, generator=g)
    sa = torch.randn(N, A, generator=g) if scales else None
    sb = torch.randn(N, B, generator=g) if scales else None
    want, scale = _want(a, b, sa, sb, K, py, px, flip, layout)
    dw = S.aligned(np.full(want.shape, np.nan))[0]
    as_, bs_, sas, sbs = _a(a), _a(b), _a(sa), _a(sb)
    before = lib.shim_mma_instructions()
    rc = lib.tc_wgrad(_p(as_), _p(bs_), _p(dw), N, A, HA, WA, B, HB, WB, K, py, px, flip, layout, _p(sas), _p(sbs), nprod, 0, 0)
    assert rc == 0, lib.shim_error()
    tol = 5e-6 if nprod == 3 else 2e-3
    assert np.abs(dw - want).max() <= tol * max(scale, np.abs(want).max()), name
    assert lib.shim_mma_instructions() > before


def test_wgrad_tma_phase_major_hint_source_on_the_hardware_model(lib):
    """gg_conv2d_wgrad_pm_f32: dimension 0 of dw holds four phase groups, bit (group*4 + ky*2 + kx) of pm_dead marks taps that are zero by
    construction in the phase-major stride-2 weight -- the kernel may leave them zero; every other entry must be exact."""
    N, A, B, HA, WA, K = 2, 32, 64, 12, 16, 2
    g = torch.Generator().manual_seed(3)
    a, b = torch.randn(N, A, HA, WA, generator=g), torch.randn(N, B, HA + 1, WA + 4, generator=g)
    want, scale = _want(a, b, None, None, K, 1, 1, 0, 0)
    dead = 0
    for grp in range(4):                                    # an arbitrary but structured pattern: group g keeps tap g only
        for t in range(4):
            if t != grp:
                dead |= 1 << (grp * 4 + t)
    dw = S.aligned(np.full(want.shape, np.nan))[0]
    as_, bs_ = _a(a), _a(b)
    assert lib.tc_wgrad(_p(as_), _p(bs_), _p(dw), N, A, HA, WA, B, HA + 1, WA + 4, K, 1, 1, 0, 0, None, None, 3, 1, dead) == 0, lib.shim_error()
    grp = np.arange(B) // (B // 4)
    live = np.zeros((B, A, 2, 2), bool)
    for o in range(B):
        live[o, :, grp[o] // 2, grp[o] % 2] = True
    assert np.abs(dw - want)[live].max() <= 5e-6 * max(scale, np.abs(want).max())
    assert ((dw == 0) | (np.abs(dw - want) <= 5e-6 * max(scale, np.abs(want).max())))[~live].all()       # dead taps: zero or computed, never garbage


def test_wgrad_tma_eligibility_source(lib):
    a = S.aligned(np.zeros((1, 16, 8, 8)))[0]
    dw = S.aligned(np.zeros((16, 16, 3, 3)))[0]
    assert lib.tc_wgrad(_p(a), _p(a), _p(dw), 1, 16, 8, 8, 16, 8, 8, 3, 1, 1, 0, 0, None, None, 3, 0, 0) == 0
    assert lib.tc_wgrad(_p(a) + 4, _p(a), _p(dw), 1, 16, 8, 8, 16, 8, 8, 3, 1, 1, 0, 0, None, None, 3, 0, 0) == -7        # unaligned rows -> wgrad_tc / FFMA
    assert lib.tc_wgrad(_p(a), _p(a), _p(dw), 1, 16, 8, 6, 16, 8, 6, 3, 1, 1, 0, 0, None, None, 3, 0, 0) == -7


SAN_CASES = [
    ('3x3', (2, 32, 48, 20, 24, 3, 1, 1, 20, 24, 3)),
    ('2x2_odd_tasks', (2, 64, 32, 16, 16, 2, 1, 1, 17, 20, 3)),
    ('2x2_odd_tasks_one_product', (2, 32, 32, 17, 16, 2, 1, 1, 18, 20, 1)),
    ('heights_16n_plus_1', (1, 32, 64, 33, 36, 2, 0, 0, 32, 36, 3)),
    ('1x1_rb128', (1, 16, 100, 20, 16, 1, 0, 0, 20, 16, 3)),
]


SAN_RUNS = [(kind, c) for c in SAN_CASES for kind in ('thread', 'address')]
SAN_DEFAULT = {'thread-2x2_odd_tasks', 'thread-2x2_odd_tasks_one_product', 'thread-heights_16n_plus_1', 'address-3x3', 'address-2x2_odd_tasks', 'thread-1x1_rb128'}


@pytest.mark.parametrize('kind,case', S.subset(SAN_RUNS, SAN_DEFAULT, id_of=lambda p: p[0] + '-' + p[1][0]))
def test_wgrad_tma_pipeline_under_sanitizers(kind, case):
    """ThreadSanitizer = racecheck of the staging ring (TMA -> converters), the X ring (shared memory) and the G ring (tensor memory) with
    their FULL / EMPTY barriers and of the accumulator ping-pong, incl. the odd task counts; AddressSanitizer = memcheck."""
    exe = S.build('wgrad_tma_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, case[1], timeout=1500)
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum')


# (with 64 input channels per tile the G ring has 4 slots against 8 X slots, and the wait for G slot t-4 already orders the converter
#  behind every MMA that read X slot t-8: the X_EMPTY wait only matters with the 8-slot G ring of the 32-channel tile -- first mutant)
MUTANTS = [
    ('converter-does-not-wait-for-the-x-slot', 'mbar_wait_block(BAR_X_EMPTY(xslot), ((xc / XS) & 1) ^ 1);', '', (2, 32, 32, 16, 16, 2, 1, 1, 17, 20, 3)),
    ('converter-does-not-wait-for-the-g-slot', 'mbar_wait_block(BAR_G_EMPTY(gslot), gphase ^ 1u);', '', (2, 64, 32, 16, 16, 2, 1, 1, 17, 20, 3)),
    ('converter-does-not-wait-for-the-tma', 'mbar_wait_block(BAR_STG_FULL(slot), (tcn / ST) & 1);', '', (2, 64, 32, 16, 16, 2, 1, 1, 17, 20, 3)),
]


@pytest.mark.parametrize('name,old,new,args', S.subset(MUTANTS, {'converter-does-not-wait-for-the-x-slot'}))
def test_the_racecheck_does_report_a_broken_weight_gradient_pipeline(name, old, new, args):
    src = _source()
    assert src.count(old) == 1, old
    exe = S.build('wgrad_tma_mutant_' + name.replace('-', '_'), src.replace(old, new), 'thread', SAN_MAIN)
    reported, out = S.mutant_is_reported(exe, args)
    if reported is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert reported, out[-2000:]
