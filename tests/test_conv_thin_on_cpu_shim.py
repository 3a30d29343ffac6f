"""The whole translation unit ga-gan_b200/csrc/conv_thin.cu -- the HBM-streaming 1x1 kernels of ToRGB (C -> 3), fromRGB (3 -> C),
their data gradients (transposed weight layout) and weight gradients, with the per-sample input / output scales of the modulated
form: eligibility checks, launch arithmetic and the three kernels, unmodified -- compiled with g++ against tests/cuda_cpu_shim.h and
executed on the CPU, against float64 numpy, and under ThreadSanitizer / AddressSanitizer with exact-size tensors (the CPU stand-in for
the closed `compute-sanitizer`).  Replaces, in the product, the ATen conv calls of the reference's
conv2d_gradfix.py:141-146,178-188 for training/networks.py:957-963 (ToRGBLayer) and :1254-1258 (fromrgb)."""
import ctypes
import os

import numpy as np
import pytest

from tests import cpu_shim as S

EXPORTS = r'''
extern "C" int thin_fwd(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int w_io, const float* is, const float* os) {
    if (!gg::conv1x1_thin_eligible(x, y, N, I, H, W, O, 1, 1, H, W, 1, 0, 0)) return -7;
    return gg::conv1x1_thin(x, w, y, N, I, H, W, O, w_io, is, os, nullptr);
}
extern "C" int thin_wgrad(const float* a, const float* b, float* dw, int N, int A, int H, int W, int B, int out_layout, const float* as, const float* bs) {
    if (!gg::wgrad1x1_thin_eligible(a, b, N, A, H, W, B, H, W, 1, 1, 1, 0, 0)) return -7;
    return gg::wgrad1x1_thin(a, b, dw, N, A, H, W, B, out_layout, as, bs, nullptr);
}
extern "C" int thin_eligible(const float* x, const float* y, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int stride, int py, int px) {
    return gg::conv1x1_thin_eligible(x, y, N, I, H, W, O, KH, KW, OH, OW, stride, py, px) ? 1 : 0;
}
'''

SAN_MAIN = r'''
#include <cstdlib>
static float* tensor(size_t n, float scale) {            // exact-size, 16-byte aligned: the sanitizer's red zone starts behind element n-1
    float* p = (float*)aligned_alloc(16, (n * 4 + 15) / 16 * 16);
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N C T H W  -- C wide channels, T thin ones: forward both ways, both weight layouts, weight gradient both ways
    const int N = atoi(argv[1]), C = atoi(argv[2]), T = atoi(argv[3]), H = atoi(argv[4]), W = atoi(argv[5]);
    const size_t P = (size_t)H * W;
    float *wide = tensor(N * C * P, 1.f), *thin = tensor(N * T * P, .7f), *w = tensor((size_t)C * T, .5f), *dw = tensor((size_t)C * T, 0.f),
          *sc = tensor((size_t)N * C, 1.2f), *st = tensor((size_t)N * T, .8f), *ywide = tensor(N * C * P, 0.f), *ythin = tensor(N * T * P, 0.f);
    int rc = 0;
    rc |= thin_fwd(wide, w, ythin, N, C, H, W, T, 0, sc, st);          // ToRGB, modulated
    rc |= thin_fwd(wide, w, ythin, N, C, H, W, T, 1, nullptr, nullptr); // data gradient of fromRGB (weights read transposed)
    rc |= thin_fwd(thin, w, ywide, N, T, H, W, C, 0, nullptr, sc);     // fromRGB
    rc |= thin_fwd(thin, w, ywide, N, T, H, W, C, 1, st, nullptr);     // data gradient of ToRGB
    rc |= thin_wgrad(wide, thin, dw, N, C, H, W, T, 0, sc, st);        // ToRGB weight gradient: a = wide input, b = thin gradient
    rc |= thin_wgrad(thin, wide, dw, N, T, H, W, C, 1, nullptr, nullptr);
    double s = 0; for (size_t i = 0; i < N * T * P; ++i) s += ythin[i]; for (size_t i = 0; i < N * C * P; ++i) s += ywide[i]; for (int i = 0; i < C * T; ++i) s += dw[i];
    printf("rc %d checksum %.5f\n", rc, s);
    free(wide); free(thin); free(w); free(dw); free(sc); free(st); free(ywide); free(ythin);
    return rc;
}
'''


def _source():
    return S.translate_unit(open(os.path.join(S.CSRC, 'conv_thin.cu')).read(), expect_launches=3) + EXPORTS


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('conv_thin_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('conv_thin_unit', _source(), 'lib'))
    P, I = ctypes.c_void_p, ctypes.c_int
    so.thin_fwd.restype = I
    so.thin_fwd.argtypes = [P, P, P, I, I, I, I, I, I, P, P]
    so.thin_wgrad.restype = I
    so.thin_wgrad.argtypes = [P, P, P, I, I, I, I, I, I, P, P]
    so.thin_eligible.restype = I
    so.thin_eligible.argtypes = [P, P] + [I] * 12
    return so


def _p(a):
    return None if a is None else a.ctypes.data


def _r(rng, *shape):
    return S.aligned(rng.standard_normal(shape))[0]


@pytest.mark.parametrize('N,I,O,H,W', [(2, 32, 3, 8, 8), (1, 512, 3, 4, 4), (3, 7, 1, 2, 6), (2, 3, 32, 8, 8), (1, 1, 512, 4, 4), (2, 4, 4, 16, 16), (1, 3, 5, 64, 68)],
                         ids=['torgb', 'torgb-512', 'odd-to-1', 'fromrgb', 'one-to-512', 'four-to-four', 'grid-stride'])
@pytest.mark.parametrize('w_io', [0, 1], ids=['w[o,i]', 'w[i,o]'])
@pytest.mark.parametrize('scales', ['none', 'in', 'both'])
def test_thin_1x1_convolution_source_on_the_cpu(lib, N, I, O, H, W, w_io, scales):
    rng = np.random.default_rng(N + 10 * I + 100 * O + H)
    x = _r(rng, N, I, H * W)
    w = _r(rng, *((I, O) if w_io else (O, I)))
    si = _r(rng, N, I) if scales in ('in', 'both') else None
    so = _r(rng, N, O) if scales == 'both' else None
    y = S.aligned(np.full((N, O, H * W), np.nan))[0]
    lib.shim_reset()
    assert lib.thin_fwd(_p(x), _p(w), _p(y), N, I, H, W, O, w_io, _p(si), _p(so)) == 0, lib.shim_error()
    wm = (w.T if w_io else w).astype(np.float64)
    xs = x.astype(np.float64) * (1 if si is None else si[:, :, None])
    want = np.einsum('oi,nip->nop', wm, xs) * (1 if so is None else so[:, :, None])
    assert np.abs(y - want).max() <= 3e-6 * np.abs(want).max()
    gx = max(1, min(-(-(H * W // 4) // 256), -(-148 * 8 // N)))
    assert lib.shim_blocks_since_reset() == gx * N and lib.shim_threads() == 256


@pytest.mark.parametrize('N,A,B,H,W', [(2, 32, 3, 8, 8), (2, 3, 32, 8, 8), (1, 40, 2, 4, 12), (3, 1, 17, 2, 2), (1, 3, 16, 128, 136), (2, 4, 4, 8, 8)],
                         ids=['torgb', 'fromrgb', 'three-passes', 'one-float4', 'pixel-chunks', 'four-by-four'])
@pytest.mark.parametrize('out_layout', [0, 1], ids=['dw[b,a]', 'dw[a,b]'])
@pytest.mark.parametrize('scales', [False, True], ids=['plain', 'scaled'])
def test_thin_1x1_weight_gradient_source_on_the_cpu(lib, N, A, B, H, W, out_layout, scales):
    """dw[b,a] = sum_{n,p} gs[n,b] G[n,b,p] * xs[n,a] X[n,a,p]; the wide side is cut into passes of 16 channels (grid.z), the pixels
    into chunks (grid.x) that meet in fp32 atomics."""
    rng = np.random.default_rng(N + 10 * A + 100 * B + H)
    a, b = _r(rng, N, A, H * W), _r(rng, N, B, H * W)
    sa = _r(rng, N, A) if scales else None
    sb = _r(rng, N, B) if scales else None
    dw = S.aligned(np.full((A, B) if out_layout else (B, A), np.nan))[0]
    lib.shim_reset()
    assert lib.thin_wgrad(_p(a), _p(b), _p(dw), N, A, H, W, B, out_layout, _p(sa), _p(sb)) == 0, lib.shim_error()
    a64 = a.astype(np.float64) * (1 if sa is None else sa[:, :, None])
    b64 = b.astype(np.float64) * (1 if sb is None else sb[:, :, None])
    want = np.einsum('nbp,nap->ba', b64, a64)
    got = dw.T if out_layout else dw
    scale = np.sqrt((a64 ** 2).sum() * (b64 ** 2).sum() / (A * B))
    assert np.abs(got - want).max() <= 3e-6 * scale
    chunks = max(1, min(-(-(H * W // 4) // 4096), -(-148 * 8 // N)))
    assert lib.shim_blocks_since_reset() == chunks * N * -(-max(A, B) // 16)


def test_thin_eligibility_source(lib):
    x = S.aligned(np.zeros(64))[0]
    ok = lambda **k: lib.thin_eligible(_p(k.get('x', x)), _p(x), *[k.get(n, d) for n, d in
                                       (('N', 2), ('I', 32), ('H', 4), ('W', 4), ('O', 3), ('KH', 1), ('KW', 1), ('OH', 4), ('OW', 4), ('stride', 1), ('py', 0), ('px', 0))])
    assert ok() == 1
    assert ok(KH=3) == 0 and ok(stride=2) == 0 and ok(py=1) == 0 and ok(OH=5) == 0                  # only the plain 1x1 form
    assert ok(I=5, O=5) == 0 and ok(I=4, O=512) == 1 and ok(I=4, O=513) == 0 and ok(I=513) == 0     # one side <= 4, the other <= 512
    assert ok(H=3, W=3, OH=3, OW=3) == 0 and ok(H=2, W=6, OH=2, OW=6) == 1                                                   # planes in 128-bit groups
    assert ok(N=0) == 0 and ok(N=65536) == 0                                                         # grid.y
    assert lib.thin_eligible(_p(x) + 4, _p(x), 2, 32, 4, 4, 3, 1, 1, 4, 4, 1, 0, 0) == 0             # 16-byte alignment


@pytest.mark.parametrize('kind', ['thread', 'address'])
@pytest.mark.parametrize('args', [(2, 20, 3, 4, 4), (1, 8, 1, 2, 2), (1, 16, 4, 128, 136)], ids=['rgb', 'one-float4', 'chunks'])
def test_conv_thin_translation_unit_under_sanitizers(kind, args):
    """ThreadSanitizer: the shared weight tile (written, barrier, read) and the weight gradient's warp / block reduction are race-free.
    AddressSanitizer: with exact-size tensors nothing is touched outside them -- incl. the channel guards `c0 + c < CW`, `t < CT`
    of the 16 x 4 register tile and the last pixel chunk."""
    if kind == 'thread' and args[3] >= 128:
        pytest.skip('covered by the address run; under ThreadSanitizer the 64 warp reductions per thread take minutes')
    exe = S.build('conv_thin_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, args)
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum')
