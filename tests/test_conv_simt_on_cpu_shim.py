"""The whole translation unit ga-gan_b200/csrc/conv_simt.cu -- the exact-fp32 (FFMA) implicit-GEMM convolution and weight-gradient
kernels that serve every shape the tcgen05 path does not take (and that the GPU suite uses as the on-device cross-check of the tensor
core kernels): launch arithmetic (tile grid, split-K decision) and both kernels, unmodified -- compiled with g++ against
tests/cuda_cpu_shim.h and executed on the CPU, against float64 torch convolutions, and under ThreadSanitizer / AddressSanitizer with
exact-size tensors (the CPU stand-in for the closed `compute-sanitizer`).  The operator semantics are those of
include/gagan_b200.h::gg_conv2d_f32 / gg_conv2d_wgrad_f32, i.e. of the reference's conv2d_gradfix.py:37-58,138-148,175-191 with the
per-sample scales of training/networks.py:641-653 folded in."""
import ctypes
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import cpu_shim as S

EXPORTS = r'''
extern "C" int simt_conv(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int KH, int KW, int OH, int OW, int stride, int pad_y, int pad_x,
                         int transposed, int flip_w, const float* is, const float* os) {
    return gg::conv2d_simt(x, w, y, N, I, H, W, O, KH, KW, OH, OW, stride, pad_y, pad_x, transposed, flip_w, is, os, nullptr);
}
extern "C" int simt_wgrad(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int KH, int KW, int stride, int pad_y, int pad_x,
                          int flip_w, int out_layout, const float* as, const float* bs) {
    return gg::conv2d_wgrad_simt(a, b, dw, N, A, HA, WA, B, HB, WB, KH, KW, stride, pad_y, pad_x, flip_w, out_layout, as, bs, nullptr);
}
'''

SAN_MAIN = r'''
#include <cstdlib>
static float* tensor(size_t n, float scale) {            // exact-size: the sanitizer's red zone starts behind element n-1
    float* p = (float*)malloc(n * 4 ? n * 4 : 4);
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N I O H W K stride pad  -- correlation, its transposed form, and the weight gradient (plain and split-K) on exact-size tensors
    const int N = atoi(argv[1]), I = atoi(argv[2]), O = atoi(argv[3]), H = atoi(argv[4]), W = atoi(argv[5]), K = atoi(argv[6]), s = atoi(argv[7]), pad = atoi(argv[8]);
    const int OH = (H + 2 * pad - K) / s + 1, OW = (W + 2 * pad - K) / s + 1;
    const int TH = (OH - 1) * s - 2 * pad + K, TW = (OW - 1) * s - 2 * pad + K;            // conv_transpose2d of the output: back to (about) H x W
    float *x = tensor((size_t)N * I * H * W, 1.f), *w = tensor((size_t)O * I * K * K, .5f), *y = tensor((size_t)N * O * OH * OW, 0.f),
          *xt = tensor((size_t)N * I * TH * TW, 0.f), *dw = tensor((size_t)O * I * K * K, 0.f), *si = tensor((size_t)N * I, 1.1f), *so = tensor((size_t)N * O, .9f);
    int rc = 0;
    rc |= simt_conv(x, w, y, N, I, H, W, O, K, K, OH, OW, s, pad, pad, 0, 0, si, so);
    rc |= simt_conv(x, w, y, N, I, H, W, O, K, K, OH, OW, s, pad, pad, 0, 1, nullptr, nullptr);
    rc |= simt_conv(y, w, xt, N, O, OH, OW, I, K, K, TH, TW, s, pad, pad, 1, 0, so, si);      // data gradient: w [O,I,..] read as the transposed-conv weight
    rc |= simt_wgrad(x, y, dw, N, I, H, W, O, OH, OW, K, K, s, pad, pad, 0, 0, si, so);
    rc |= simt_wgrad(x, y, dw, N, I, H, W, O, OH, OW, K, K, s, pad, pad, 1, 1, nullptr, nullptr);
    double c = 0; for (size_t i = 0; i < (size_t)N * O * OH * OW; ++i) c += y[i]; for (size_t i = 0; i < (size_t)N * I * TH * TW; ++i) c += xt[i];
    for (size_t i = 0; i < (size_t)O * I * K * K; ++i) c += dw[i];
    printf("rc %d checksum %.5f\n", rc, c);
    free(x); free(w); free(y); free(xt); free(dw); free(si); free(so);
    return rc;
}
'''


def _source():
    return S.translate_unit(open(os.path.join(S.CSRC, 'conv_simt.cu')).read(), expect_launches=2) + EXPORTS


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('conv_simt_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('conv_simt_unit', _source(), 'lib'))
    P, I = ctypes.c_void_p, ctypes.c_int
    so.simt_conv.restype = I
    so.simt_conv.argtypes = [P, P, P] + [I] * 14 + [P, P]
    so.simt_wgrad.restype = I
    so.simt_wgrad.argtypes = [P, P, P] + [I] * 14 + [P, P]
    return so


def _p(a):
    return None if a is None else a.ctypes.data


def _np(t):
    return np.ascontiguousarray(t.numpy(), np.float32)


# N, I, O, H, W, K, stride, pad, transposed, flip, scales
FWD = [
    (2, 5, 7, 6, 6, 3, 1, 1, 0, 0, True),       # the synthesis layer form, one tile
    (1, 3, 70, 9, 7, 3, 1, 1, 0, 1, False),     # two n-tiles (O > 64), flipped weights (conv2d_resample.py:35-36)
    (3, 4, 6, 13, 11, 3, 2, 1, 0, 0, True),     # stride 2 (discriminator down path without the FIR), M > one 128-pixel tile
    (2, 6, 4, 5, 5, 3, 2, 1, 1, 0, True),       # conv_transpose2d stride 2 (the up-sampling layer)
    (1, 8, 3, 4, 4, 1, 1, 0, 0, 0, True),       # 1x1 (ToRGB shape on the exact path)
    (2, 3, 5, 7, 8, 4, 1, 2, 1, 1, False),      # even kernel, transposed, flipped
    (1, 2, 2, 20, 20, 3, 1, 0, 0, 0, False),    # no padding, 324 pixels: three m-tiles, the last one ragged
]


@pytest.mark.parametrize('N,I,O,H,W,K,stride,pad,transposed,flip,scales', FWD,
                         ids=['synthesis', 'two-n-tiles-flip', 'stride2', 'transposed2', '1x1', 'even-transposed-flip', 'ragged-m'])
def test_conv_simt_source_on_the_cpu(lib, N, I, O, H, W, K, stride, pad, transposed, flip, scales):
    g = torch.Generator().manual_seed(N * 1000 + I * 100 + O * 10 + K)
    x = torch.randn(N, I, H, W, generator=g)
    w = torch.randn((I, O, K, K) if transposed else (O, I, K, K), generator=g)
    si = torch.randn(N, I, generator=g) if scales else None
    so = torch.randn(N, O, generator=g) if scales else None
    xd = x.double() * (si.double()[:, :, None, None] if scales else 1)
    wd = w.double().flip([2, 3]) if flip else w.double()
    want = F.conv_transpose2d(xd, wd, stride=stride, padding=pad) if transposed else F.conv2d(xd, wd, stride=stride, padding=pad)
    if scales:
        want = want * so.double()[:, :, None, None]
    OH, OW = want.shape[2:]
    y = np.full((N, O, OH, OW), np.nan, np.float32)
    xs, ws = _np(x), _np(w)
    sis, sos = (_np(si), _np(so)) if scales else (None, None)
    lib.shim_reset()
    assert lib.simt_conv(_p(xs), _p(ws), _p(y), N, I, H, W, O, K, K, OH, OW, stride, pad, pad, transposed, flip, _p(sis), _p(sos)) == 0, lib.shim_error()
    assert np.abs(y - want.numpy()).max() <= 3e-6 * float(want.abs().max())
    assert lib.shim_blocks_since_reset() == -(-N * OH * OW // 128) * -(-O // 64) and lib.shim_threads() == 256


def test_conv_simt_source_free_output_extent_at_stride_1(lib):
    """stride 1: OH / OW are free and (pad_y, pad_x) is the top / left padding -- positions that read outside x see zeros.  This is what
    makes the operator closed under differentiation without copies (conv2d_gradfix.conv2d_s1); different pads per axis."""
    g = torch.Generator().manual_seed(5)
    N, I, O, H, W, K = 2, 3, 4, 5, 6, 3
    x, w = torch.randn(N, I, H, W, generator=g), torch.randn(O, I, K, K, generator=g)
    pad_y, pad_x, OH, OW = 2, 0, 8, 5
    xp = F.pad(x.double(), [pad_x, OW + K - 1 - W - pad_x, pad_y, OH + K - 1 - H - pad_y])
    want = F.conv2d(xp, w.double())
    assert tuple(want.shape[2:]) == (OH, OW)
    y = np.full((N, O, OH, OW), np.nan, np.float32)
    xs, ws = _np(x), _np(w)
    assert lib.simt_conv(_p(xs), _p(ws), _p(y), N, I, H, W, O, K, K, OH, OW, 1, pad_y, pad_x, 0, 0, None, None) == 0
    assert np.abs(y - want.numpy()).max() <= 3e-6 * float(want.abs().max())


def test_conv_simt_source_degenerate_extents(lib):
    y = np.full((2, 3, 4, 4), np.nan, np.float32)
    x = np.ones((2, 1, 4, 4), np.float32)
    assert lib.simt_conv(_p(x), _p(x), _p(y), 2, 0, 4, 4, 3, 3, 3, 4, 4, 1, 1, 1, 0, 0, None, None) == 0 and (y == 0).all()    # no input channels: zeros
    y[:] = 7
    assert lib.simt_conv(_p(x), _p(x), _p(y), 0, 1, 4, 4, 3, 3, 3, 4, 4, 1, 1, 1, 0, 0, None, None) == 0 and (y == 7).all()    # empty batch: untouched
    dw = np.full((3, 1, 3, 3), np.nan, np.float32)
    assert lib.simt_wgrad(_p(x), _p(y), _p(dw), 0, 1, 4, 4, 3, 4, 4, 3, 3, 1, 1, 1, 0, 0, None, None) == 0 and (dw == 0).all()  # empty batch: zero gradient


# N, A, B, HA, WA, K, stride, pad, flip, out_layout, scales
WG = [
    (2, 5, 7, 6, 6, 3, 1, 1, 0, 0, True),       # split-K over the pixels (atomics): few tiles, 72 pixels
    (1, 70, 66, 4, 4, 3, 1, 1, 0, 0, False),    # 2 x 10 tiles of 64 x 64
    (3, 4, 6, 13, 11, 3, 2, 1, 1, 0, True),     # stride 2, flipped gradient
    (2, 6, 4, 9, 9, 3, 1, 1, 0, 1, True),       # transposed output layout [A,B,KH,KW]
    (1, 3, 5, 8, 8, 1, 1, 0, 0, 0, False),      # 1x1
    (4, 2, 3, 33, 31, 3, 1, 1, 0, 0, False),    # 4092 pixels: many K splits with a ragged last one
]


@pytest.mark.parametrize('N,A,B,HA,WA,K,stride,pad,flip,out_layout,scales', WG, ids=['split-k', 'tiles', 'stride2-flip', 'layout-ab', '1x1', 'ragged-split'])
def test_wgrad_simt_source_on_the_cpu(lib, N, A, B, HA, WA, K, stride, pad, flip, out_layout, scales):
    g = torch.Generator().manual_seed(N * 1000 + A * 100 + B * 10 + K)
    a = torch.randn(N, A, HA, WA, generator=g)
    HB, WB = (HA + 2 * pad - K) // stride + 1, (WA + 2 * pad - K) // stride + 1
    b = torch.randn(N, B, HB, WB, generator=g)
    sa = torch.randn(N, A, generator=g) if scales else None
    sb = torch.randn(N, B, generator=g) if scales else None
    ad = a.double() * (sa.double()[:, :, None, None] if scales else 1)
    bd = b.double() * (sb.double()[:, :, None, None] if scales else 1)
    wv = torch.zeros(B, A, K, K, dtype=torch.float64, requires_grad=True)
    (F.conv2d(ad, wv, stride=stride, padding=pad) * bd).sum().backward()
    want = wv.grad
    if flip:
        want = want.flip([2, 3])
    if out_layout:
        want = want.transpose(0, 1)
    want = want.contiguous().numpy()
    dw = np.full(want.shape, np.nan, np.float32)
    as_, bs_ = _np(a), _np(b)
    sas, sbs = (_np(sa), _np(sb)) if scales else (None, None)
    assert lib.simt_wgrad(_p(as_), _p(bs_), _p(dw), N, A, HA, WA, B, HB, WB, K, K, stride, pad, pad, flip, out_layout, _p(sas), _p(sbs)) == 0, lib.shim_error()
    scale = float(((ad ** 2).sum() * (bd ** 2).sum() / (A * B)).sqrt())
    assert np.abs(dw - want).max() <= 3e-6 * max(scale, np.abs(want).max())


def test_wgrad_simt_source_gradient_extent_larger_than_natural(lib):
    """`b` may be larger than the natural correlation output (the stride-1 operator with a free extent): positions of `a` outside its
    extent count as zeros."""
    g = torch.Generator().manual_seed(9)
    N, A, B, HA, WA, K, pad_y, pad_x, HB, WB = 2, 3, 4, 5, 6, 3, 2, 0, 8, 5
    a, b = torch.randn(N, A, HA, WA, generator=g), torch.randn(N, B, HB, WB, generator=g)
    ap = F.pad(a.double(), [pad_x, WB + K - 1 - WA - pad_x, pad_y, HB + K - 1 - HA - pad_y])
    wv = torch.zeros(B, A, K, K, dtype=torch.float64, requires_grad=True)
    (F.conv2d(ap, wv) * b.double()).sum().backward()
    dw = np.full((B, A, K, K), np.nan, np.float32)
    as_, bs_ = _np(a), _np(b)
    assert lib.simt_wgrad(_p(as_), _p(bs_), _p(dw), N, A, HA, WA, B, HB, WB, K, K, 1, pad_y, pad_x, 0, 0, None, None) == 0
    assert np.abs(dw - wv.grad.numpy()).max() <= 3e-6 * float(wv.grad.abs().max())


@pytest.mark.parametrize('kind', ['thread', 'address'])
@pytest.mark.parametrize('args', [(2, 5, 7, 6, 6, 3, 1, 1), (1, 3, 66, 7, 5, 3, 2, 1), (1, 9, 4, 4, 4, 1, 1, 0)], ids=['3x3', 'stride2-two-n-tiles', '1x1'])
def test_conv_simt_translation_unit_under_sanitizers(kind, args):
    """ThreadSanitizer: the double-buffered operand tiles (register prefetch -> shared store -> barrier -> FFMA) of both kernels and the
    split-K atomics are race-free.  AddressSanitizer: with exact-size, unpadded tensors the gather addressing (stride, zero padding,
    transposed form, ragged M / N / K tiles) never touches a byte outside them."""
    exe = S.build('conv_simt_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, args)
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum')
