"""The whole translation unit ga-gan_b200/csrc/conv_march.cu -- the row-marching tcgen05 convolution for <= 64 output channels (the
32- and 64-channel layers at 512^2 / 1024^2): one raw image row per TMA box, converted once into a SWIZZLE_128B K-major operand ring
whose descriptor START ADDRESSES march row by row (so the hardware's swizzle works on absolute shared-memory address bits, DESIGN.md
section 4.1), all three ky taps in one N = 3 NT instruction, weights resident or ringed -- with its pack kernel and host code, compiled
with g++ against tests/tc_cpu_shim.h and executed on the CPU: against float64 convolutions, under ThreadSanitizer (racecheck of the raw /
converted / weight / accumulator rings; mbarriers are acquire / release atomics in the model) and under AddressSanitizer (exact-size
tensors, scratch and shared memory).  The model applies the 128-byte swizzle exactly as the kernel's converter assumes the tensor core
does; that the two agree here, and the kernel agrees with the hardware in the GPU suite, pins the model as well."""
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
void gg::keep_pool_memory() {}
extern "C" int march_conv(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int OH, int OW, int pad_y, int pad_x, int flip_w, int w_is_IO,
                          const float* is, const float* os, int nprod, const float* bias, const float* noise, long long noise_bs, int act, float alpha, float gain, float clamp) {
    if (!gg::conv2d_march_eligible(x, N, I, H, W, O, 3, 3, OH, OW, 1, pad_y, pad_x)) return -7;
    ggtc::ConvEpilogue e{bias, noise, noise_bs, act, alpha, gain, clamp};
    return gg::conv2d_march(x, w, y, N, I, H, W, O, 3, OH, OW, pad_y, pad_x, flip_w, w_is_IO, is, os, nprod, act ? &e : nullptr, nullptr);
}
'''

SAN_MAIN = r'''
static float* tensor(size_t n, float scale) {
    float* p = (float*)aligned_alloc(16, (n * 4 + 15) / 16 * 16);
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N I H W O pad  -- plain, modulated + flipped, fused-epilogue launches on exact-size tensors
    const int N = atoi(argv[1]), I = atoi(argv[2]), H = atoi(argv[3]), W = atoi(argv[4]), O = atoi(argv[5]), pad = atoi(argv[6]);
    const int OH = H + 2 * pad - 2, OW = W + 2 * pad - 2;
    float *x = tensor((size_t)N * I * H * W, 1.f), *w = tensor((size_t)O * I * 9, .5f), *y = tensor((size_t)N * O * OH * OW, 0.f), *si = tensor((size_t)N * I, 1.1f),
          *so = tensor((size_t)N * O, .9f), *bias = tensor(O, .3f), *noise = tensor((size_t)OH * OW, .2f);
    int rc = march_conv(x, w, y, N, I, H, W, O, OH, OW, pad, pad, 0, 0, nullptr, nullptr, 3, nullptr, nullptr, 0, 0, 0.f, 1.f, -1.f);
    rc |= march_conv(x, w, y, N, I, H, W, O, OH, OW, pad, pad, 1, 1, si, so, 3, nullptr, nullptr, 0, 0, 0.f, 1.f, -1.f);
    rc |= march_conv(x, w, y, N, I, H, W, O, OH, OW, pad, pad, 0, 0, si, so, 1, bias, noise, 0, 3, .2f, 1.4f, 2.f);
    double s = 0; for (size_t i = 0; i < (size_t)N * O * OH * OW; ++i) s += y[i];
    printf("rc %d checksum %.5f mma %ld scratch %ld\n", rc, s, shim_mma_instructions(), shim_scratch_blocks_live());
    if (rc) printf("%s\n", shim_error());
    free(x); free(w); free(y); free(si); free(so); free(bias); free(noise);
    return rc;
}
'''


def _source():
    return '#define GG_NUM_SMS 3\n' + S.translate_tc_unit(open(os.path.join(S.CSRC, 'conv_march.cu')).read(), expect_launches=2) + EXPORTS


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('conv_march_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('conv_march_unit', _source(), 'lib'))
    P, I, F32, LL = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
    so.march_conv.restype = I
    so.march_conv.argtypes = [P, P, P] + [I] * 11 + [P, P, I, P, P, LL, I, F32, F32, F32]
    so.shim_mma_instructions.restype = ctypes.c_long
    so.shim_scratch_blocks_live.restype = ctypes.c_long
    return so


def _a(t):
    return None if t is None else S.aligned(t.numpy())[0]


def _p(a):
    return None if a is None else a.ctypes.data


# name, N, I, H, W, O, (pad_y, pad_x), (OH, OW) or None, flip, w_is_IO, scales, nprod
CASES = [
    ('nt32_resident_weights', 1, 16, 8, 64, 32, (1, 1), None, 0, 0, False, 3),
    ('nt64_modulated_flip', 2, 32, 20, 72, 48, (1, 1), None, 1, 0, True, 3),                 # ragged 128-pixel segment, two bands
    ('nt64_weight_ring', 1, 64, 40, 136, 64, (1, 1), None, 0, 0, True, 3),                    # 4 x 3 weight blocks > the resident stages: the ring path
    ('nt32_io_layout_odd_channels', 1, 24, 12, 64, 20, (1, 1), None, 0, 1, True, 3),          # the data gradient's weight layout, I % 16 != 0, O % 32 != 0
    ('full_correlation_pad2', 1, 16, 10, 64, 16, (2, 2), None, 1, 0, False, 3),
    ('asym_pad_free_extent', 1, 16, 9, 68, 32, (0, 2), (10, 64), 0, 0, False, 3),
    ('many_units_per_cta', 3, 16, 40, 132, 32, (1, 1), None, 0, 0, True, 3),                  # image changes restage the scales; units wrap every ring
    ('one_product', 1, 16, 8, 64, 32, (1, 1), None, 0, 0, False, 1),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_conv_march_source_on_the_hardware_model(lib, case):
    name, N, I, H, W, O, (py, px), ext, flip, w_io, scales, nprod = case
    K = 3
    g = torch.Generator().manual_seed(len(name) * 5 + O)
    x, w = torch.randn(N, I, H, W, generator=g), torch.randn(O, I, K, K, generator=g)
    si = torch.randn(N, I, generator=g) if scales else None
    so = torch.randn(N, O, generator=g) if scales else None
    OH, OW = ext if ext else (H + 2 * py - K + 1, W + 2 * px - K + 1)
    xd = x.double() * (si.double()[:, :, None, None] if scales else 1)
    xp = F.pad(xd, [px, max(OW + K - 1 - W - px, 0), py, max(OH + K - 1 - H - py, 0)])[:, :, :OH + K - 1, :OW + K - 1]
    want = F.conv2d(xp, w.double().flip([2, 3]) if flip else w.double())
    if scales:
        want = want * so.double()[:, :, None, None]
    want = want.numpy()
    xs, ws, sis, sos = _a(x), _a(w.transpose(0, 1).contiguous() if w_io else w), _a(si), _a(so)
    y = S.aligned(np.full((N, O, OH, OW), np.nan))[0]
    before = lib.shim_mma_instructions()
    rc = lib.march_conv(_p(xs), _p(ws), _p(y), N, I, H, W, O, OH, OW, py, px, flip, w_io, _p(sis), _p(sos), nprod, None, None, 0, 0, 0.0, 1.0, -1.0)
    assert rc == 0, lib.shim_error()
    assert not np.isnan(y).any(), 'an output element was never written'
    assert np.abs(y - want).max() <= (5e-6 if nprod == 3 else 2e-3) * np.abs(want).max(), name
    assert lib.shim_mma_instructions() > before and lib.shim_scratch_blocks_live() == 0


def test_conv_march_fused_epilogue_source_on_the_hardware_model(lib):
    N, I, H, W, O = 2, 32, 12, 64, 40
    g = torch.Generator().manual_seed(4)
    x, w = torch.randn(N, I, H, W, generator=g), torch.randn(O, I, 3, 3, generator=g) * 0.1
    si, so, b, nz = torch.randn(N, I, generator=g), torch.randn(N, O, generator=g), torch.randn(O, generator=g), torch.randn(N, H * W, generator=g)
    v = F.conv2d(x.double() * si.double()[:, :, None, None], w.double(), padding=1) * so.double()[:, :, None, None]
    v = v + b.double()[None, :, None, None] + nz.double().reshape(N, 1, H, W)
    v = (torch.where(v > 0, v, v * 0.2) * 1.4142).clamp(-1.5, 1.5)
    y = S.aligned(np.full((N, O, H, W), np.nan))[0]
    xs, ws, sis, sos, bs, ns = _a(x), _a(w), _a(si), _a(so), _a(b), _a(nz)
    assert lib.march_conv(_p(xs), _p(ws), _p(y), N, I, H, W, O, H, W, 1, 1, 0, 0, _p(sis), _p(sos), 3, _p(bs), _p(ns), H * W, 3, 0.2, 1.4142, 1.5) == 0, lib.shim_error()
    assert np.abs(y - v.numpy()).max() <= 5e-6 * max(1.0, float(v.abs().max()))


def test_conv_march_eligibility_source(lib):
    x = S.aligned(np.zeros((1, 16, 8, 64)))[0]
    w = S.aligned(np.zeros((32, 16, 3, 3)))[0]
    y = S.aligned(np.zeros((1, 32, 8, 64)))[0]
    call = lambda xp=None, **k: lib.march_conv(xp or _p(x), _p(w), _p(y), *[{**dict(N=1, I=16, H=8, W=64, O=32, OH=8, OW=64, py=1, px=1), **k}[n]
                                                                            for n in ('N', 'I', 'H', 'W', 'O', 'OH', 'OW', 'py', 'px')], 0, 0, None, None, 3, None, None, 0, 0, 0.0, 1.0, -1.0)
    assert call() == 0
    # wider / narrower layers, small maps and unaligned rows belong to conv_tc.cu or the FFMA kernels
    assert call(O=65) == -7 and call(O=8) == -7 and call(I=8) == -7 and call(OW=60, W=60) == -7 and call(OH=4, H=4) == -7 and call(px=3) == -7
    assert call(xp=_p(x) + 4) == -7


SAN_CASES = [('nt32', (2, 16, 12, 64, 32, 1)), ('nt64_weight_ring', (1, 64, 20, 72, 48, 1)), ('pad2_many_units', (3, 16, 24, 132, 16, 2))]


SAN_RUNS = [(kind, c) for c in SAN_CASES for kind in ('thread', 'address')]
SAN_DEFAULT = {'thread-nt64_weight_ring', 'thread-pad2_many_units', 'address-nt64_weight_ring', 'thread-nt32'}


@pytest.mark.parametrize('kind,case', S.subset(SAN_RUNS, SAN_DEFAULT, id_of=lambda p: p[0] + '-' + p[1][0]))
def test_conv_march_pipeline_under_sanitizers(kind, case):
    exe = S.build('conv_march_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, case[1], timeout=1500)
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum') and out.rstrip().endswith('scratch 0')


MUTANTS = [
    ('converter-does-not-wait-for-the-tma', 'mbar_wait(BAR_RAW_FULL(s), (sc >> 1) & 1);', ''),
    ('drain-does-not-wait-for-the-mma', 'mbar_wait(BAR_ACC_FULL(s), (k >> 1) & 1);', ''),
    ('issuer-does-not-wait-for-the-converter', 'mbar_wait(BAR_CVT_FULL(cs), cph);', ''),
]


@pytest.mark.parametrize('name,old,new', S.subset(MUTANTS, {'issuer-does-not-wait-for-the-converter'}))
def test_the_racecheck_does_report_a_broken_marching_pipeline(name, old, new):
    src = _source()
    assert src.count(old) == 1, old
    exe = S.build('conv_march_mutant_' + name.replace('-', '_'), src.replace(old, new), 'thread', SAN_MAIN)
    reported, out = S.mutant_is_reported(exe, (2, 32, 12, 64, 32, 1))
    if reported is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert reported, out[-2000:]
