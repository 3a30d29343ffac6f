"""The whole translation unit ga-gan_b200/csrc/conv_tc.cu -- the tcgen05 / TMEM / TMA convolution kernel that carries 35 % of the
training step, with its weight-packing kernel and its host code (n-tile choice, tensor-map encoding, shared-memory layout) -- compiled
with g++ against tests/tc_cpu_shim.h, a functional model of mbarriers, TMA box loads, tensor memory and tcgen05.mma, and EXECUTED ON THE
CPU: 384 std::threads per CTA play the TMA producers, the MMA issuers and the eight converter / drain warps of the shipped source.

What this buys ("sanitizer evidence"): `compute-sanitizer` is closed on the GPU pool, so the racecheck / memcheck of
the mbarrier pipelines that the round-1 review asked for runs here instead --
  * against float64 torch convolutions (the descriptors, the im2col-by-start-address, the zero-block skipping, the chunked
    accumulation and the epilogue produce the right numbers through the MODEL of the hardware, not only on it);
  * under ThreadSanitizer: the model implements mbarriers with acquire / release atomics and runs TMA copies and MMAs in the issuing
    thread, so every read of a stage that is not ordered behind its FULL barrier, and every refill not ordered behind its EMPTY
    barrier, is a data race the tool reports (two mutation tests remove one wait each and must fail);
  * under AddressSanitizer with exact-size tensors, scratch and shared memory.
GG_NUM_SMS is set to 3 for these builds, so that each persistent CTA walks many tiles (ring wrap-around, n-tile and image changes)."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from tests import cpu_shim as S

EXPORTS = r'''
extern "C" int tc_conv(const float* x, const float* w, float* y, int N, int I, int H, int W, int O, int K, int OH, int OW, int pad_y, int pad_x, int flip_w, int w_is_IO,
                       const float* is, const float* os, int nprod, const float* bias, const float* noise, long long noise_bs, int act, float alpha, float gain, float clamp) {
    if (!gg::conv2d_tc_eligible(N, I, H, W, O, K, K, OH, OW, 1, pad_y, pad_x, 0)) return -7;
    ggtc::ConvEpilogue e{bias, noise, noise_bs, act, alpha, gain, clamp};
    return gg::conv2d_tc(x, w, y, N, I, H, W, O, K, K, OH, OW, pad_y, pad_x, flip_w, w_is_IO, is, os, nprod, act ? &e : nullptr, nullptr);
}
'''

SAN_MAIN = r'''
static float* tensor(size_t n, float scale) {            // exact-size, 16-byte aligned: the sanitizer's red zone starts behind element n-1
    float* p = (float*)aligned_alloc(16, (n * 4 + 15) / 16 * 16);
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N I H W O K pad sparse (0: dense, m: keep the taps with (K-block + tap) % m == 0)  -- plain, modulated + flipped, and fused-epilogue launches on exact-size tensors
    const int N = atoi(argv[1]), I = atoi(argv[2]), H = atoi(argv[3]), W = atoi(argv[4]), O = atoi(argv[5]), K = atoi(argv[6]), pad = atoi(argv[7]), sparse = atoi(argv[8]);
    const int OH = H + 2 * pad - K + 1, OW = W + 2 * pad - K + 1;
    float *x = tensor((size_t)N * I * H * W, 1.f), *w = tensor((size_t)O * I * K * K, .5f), *y = tensor((size_t)N * O * OH * OW, 0.f), *si = tensor((size_t)N * I, 1.1f),
          *so = tensor((size_t)N * O, .9f), *bias = tensor(O, .3f), *noise = tensor((size_t)N * OH * OW, .2f);
    if (sparse)          // structurally zero (K-block, tap) blocks, as the phase-major stride-2 weights have them: skipped by every warp role
        for (int o = 0; o < O; ++o) for (int i = 0; i < I; ++i) for (int t = 0; t < K * K; ++t) if (((i / 16) + t) % sparse != 0) w[((size_t)o * I + i) * K * K + t] = 0.f;
    int rc = 0;
    rc |= tc_conv(x, w, y, N, I, H, W, O, K, OH, OW, pad, pad, 0, 0, nullptr, nullptr, 3, nullptr, nullptr, 0, 0, 0.f, 1.f, -1.f);
    rc |= tc_conv(x, w, y, N, I, H, W, O, K, OH, OW, pad, pad, 1, 0, si, so, 3, nullptr, nullptr, 0, 0, 0.f, 1.f, -1.f);
    rc |= tc_conv(x, w, y, N, I, H, W, O, K, OH, OW, pad, pad, 0, 0, si, so, 1, bias, noise, (long long)OH * OW, 3, .2f, 1.4f, 2.f);
    double s = 0; for (size_t i = 0; i < (size_t)N * O * OH * OW; ++i) s += y[i];
    printf("rc %d checksum %.5f mma %ld scratch %ld\n", rc, s, shim_mma_instructions(), shim_scratch_blocks_live());
    if (rc) printf("%s\n", shim_error());
    free(x); free(w); free(y); free(si); free(so); free(bias); free(noise);
    return rc;
}
'''


def _source():
    return '#define GG_NUM_SMS 3\n' + S.translate_tc_unit(open(os.path.join(S.CSRC, 'conv_tc.cu')).read(), expect_launches=2) + EXPORTS


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('conv_tc_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('conv_tc_unit', _source(), 'lib'))
    P, I, F32, LL = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_longlong
    so.tc_conv.restype = I
    so.tc_conv.argtypes = [P, P, P] + [I] * 12 + [P, P, I, P, P, LL, I, F32, F32, F32]
    so.shim_mma_instructions.restype = ctypes.c_long
    so.shim_scratch_blocks_live.restype = ctypes.c_long
    return so


def _a(t):
    return None if t is None else S.aligned(t.numpy())[0]


def _p(a):
    return None if a is None else a.ctypes.data


# name, N, I, H, W, O, K, pad (y, x), OH/OW (None = natural), flip, w_is_IO, scales, nprod, sparse
CASES = [
    ('nt32_3x3', 1, 16, 16, 16, 32, 3, (1, 1), None, 0, 0, False, 3, False),
    ('nt64_modulated_flip', 2, 32, 20, 24, 48, 3, (1, 1), None, 1, 0, True, 3, False),           # ragged tiles in x and y, two K-blocks
    ('nt128_io_layout', 1, 24, 12, 16, 100, 3, (1, 1), None, 0, 1, True, 3, False),              # weights [I,O,k,k] (the data gradient), I % 16 != 0
    ('nt256_two_n_tiles', 1, 40, 12, 16, 300, 3, (1, 1), None, 0, 0, True, 3, False),            # <256,1>: 8-pixel-wide tiles, O % 256 != 0
    ('k1', 2, 48, 8, 8, 16, 1, (0, 0), None, 0, 0, True, 3, False),
    ('k2_asym_pad', 1, 16, 9, 12, 32, 2, (1, 0), None, 0, 0, False, 3, False),                   # the 2x2 phase-major forms
    ('free_extent', 1, 16, 10, 12, 32, 3, (2, 0), (13, 9), 0, 0, False, 3, False),               # stride-1 operator with a free output extent
    ('pad_x_2_unaligned_box', 1, 16, 18, 20, 32, 3, (2, 2), None, 1, 0, False, 3, False),        # full correlation: the box start is shifted / clamped
    ('sparse_blocks_chunks_span', 2, 64, 20, 20, 32, 3, (1, 1), None, 0, 0, True, 3, True),      # zero blocks skipped; chunks span K-blocks
    ('many_tiles_per_cta', 3, 16, 40, 36, 40, 3, (1, 1), None, 0, 0, True, 3, False),            # 27 tiles over 3 CTAs: image changes restage the scales
    ('one_product', 1, 16, 16, 16, 32, 3, (1, 1), None, 0, 0, False, 1, False),                  # the tf32x1 fast mode
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_conv_tc_source_on_the_hardware_model(lib, case):
    name, N, I, H, W, O, K, (py, px), ext, flip, w_io, scales, nprod, sparse = case
    g = torch.Generator().manual_seed(len(name) * 7 + O)
    x = torch.randn(N, I, H, W, generator=g)
    w = torch.randn(O, I, K, K, generator=g)
    if sparse:
        keep = ((torch.arange(I)[:, None] // 16 + torch.arange(K * K)[None, :]) % 3 == 0).reshape(1, I, K, K)
        w = w * keep
    si = torch.randn(N, I, generator=g) if scales else None
    so = torch.randn(N, O, generator=g) if scales else None
    OH, OW = ext if ext else (H + 2 * py - K + 1, W + 2 * px - K + 1)
    xd = x.double() * (si.double()[:, :, None, None] if scales else 1)
    xp = F.pad(xd, [px, OW + K - 1 - W - px, py, OH + K - 1 - H - py])
    want = F.conv2d(xp, w.double().flip([2, 3]) if flip else w.double())
    if scales:
        want = want * so.double()[:, :, None, None]
    want = want.numpy()
    xs, ws = _a(x), _a(w.transpose(0, 1).contiguous() if w_io else w)
    sis, sos = _a(si), _a(so)
    y = S.aligned(np.full((N, O, OH, OW), np.nan))[0]
    before = lib.shim_mma_instructions()
    rc = lib.tc_conv(_p(xs), _p(ws), _p(y), N, I, H, W, O, K, OH, OW, py, px, flip, w_io, _p(sis), _p(sos), nprod, None, None, 0, 0, 0.0, 1.0, -1.0)
    assert rc == 0, lib.shim_error()
    assert not np.isnan(y).any(), 'an output element was never written'
    tol = 5e-6 if nprod == 3 else 2e-3                                     # 3xTF32 is fp32-faithful; one product is ~2^-11 per operand
    assert np.abs(y - want).max() <= tol * np.abs(want).max(), name
    assert lib.shim_mma_instructions() > before and lib.shim_scratch_blocks_live() == 0      # tensor cores did the work; the packed weights were freed
    if sparse:                                                              # 1/3 of the (K-block, tap) blocks are live: the others cost no MMA
        NT = 32
        tiles = N * -(-OH // 16) * -(-OW // 16) * -(-O // NT)
        assert lib.shim_mma_instructions() - before == tiles * 2 * (I // 16) * 3 * 6


@pytest.mark.parametrize('act,alpha,gain,clamp,noise,bias', [(3, 0.2, 1.4142, -1.0, 'sample', True), (1, 0.0, 1.0, 0.7, 'shared', True), (2, 0.0, 1.0, -1.0, None, True),
                                                             (3, 0.2, 1.4142, 0.5, 'shared', False)], ids=['lrelu-noise', 'linear-clamp', 'relu-bias', 'lrelu-clamp-nobias'])
def test_conv_tc_fused_epilogue_source_on_the_hardware_model(lib, act, alpha, gain, clamp, noise, bias):
    """gg_conv2d_act_f32's tensor-core path: y = clamp(act(out_scale * conv(in_scale * x, w) + bias[o] + noise[pixel]) * gain) in the
    store loop (SynthesisLayer.forward, reference training/networks.py:904-921, as one launch)."""
    N, I, H, W, O, K = 2, 32, 20, 16, 40, 3
    g = torch.Generator().manual_seed(act * 10 + int(clamp * 10))
    x, w = torch.randn(N, I, H, W, generator=g), torch.randn(O, I, K, K, generator=g) * 0.1
    si, so = torch.randn(N, I, generator=g), torch.randn(N, O, generator=g)
    b = torch.randn(O, generator=g) if bias else None
    nz = None if noise is None else torch.randn((N, H * W) if noise == 'sample' else (1, H * W), generator=g)
    v = F.conv2d(x.double() * si.double()[:, :, None, None], w.double(), padding=1) * so.double()[:, :, None, None]
    if bias:
        v = v + b.double()[None, :, None, None]
    if nz is not None:
        v = v + nz.double().reshape(-1, 1, H, W)
    v = {1: v, 2: v.clamp(min=0), 3: torch.where(v > 0, v, v * alpha)}[act] * gain
    if clamp >= 0:
        v = v.clamp(-clamp, clamp)
    y = S.aligned(np.full((N, O, H, W), np.nan))[0]
    xs, ws, sis, sos, bs, ns = _a(x), _a(w), _a(si), _a(so), _a(b), _a(nz)
    rc = lib.tc_conv(_p(xs), _p(ws), _p(y), N, I, H, W, O, K, H, W, 1, 1, 0, 0, _p(sis), _p(sos), 3, _p(bs), _p(ns), H * W if noise == 'sample' else 0, act, alpha, gain, clamp)
    assert rc == 0, lib.shim_error()
    assert np.abs(y - v.numpy()).max() <= 5e-6 * max(1.0, float(v.abs().max()))


def test_conv_tc_eligibility_and_alignment_source(lib):
    x = S.aligned(np.zeros((1, 16, 8, 8)))[0]
    y = S.aligned(np.zeros((1, 16, 8, 8)))[0]
    w = S.aligned(np.zeros((16, 16, 3, 3)))[0]
    call = lambda xp=None, **k: lib.tc_conv(xp or _p(x), _p(w), _p(y), *[{**dict(N=1, I=16, H=8, W=8, O=16, K=3, OH=8, OW=8, py=1, px=1), **k}[n]
                                                                         for n in ('N', 'I', 'H', 'W', 'O', 'K', 'OH', 'OW', 'py', 'px')], 0, 0, None, None, 3, None, None, 0, 0, 0.0, 1.0, -1.0)
    assert call() == 0
    assert call(I=8) == -7 and call(O=8) == -7 and call(K=4) == -7 and call(px=3) == -7 and call(W=6, OW=6) == -7     # -> the FFMA kernels
    assert call(xp=_p(x) + 4) == -1 and b'16-byte' in lib.shim_error()


SAN_CASES = [('nt32', (2, 48, 20, 20, 32, 3, 1, 0)), ('nt64_sparse', (2, 64, 20, 16, 48, 3, 1, 3)), ('nt32_chunks_span_k_blocks', (1, 96, 16, 16, 32, 3, 1, 5)), ('nt256', (1, 32, 12, 12, 272, 3, 1, 0)), ('k1', (2, 32, 16, 16, 32, 1, 0, 0)),
             ('k2', (1, 48, 12, 12, 32, 2, 1, 0))]


SAN_RUNS = [(kind, c) for c in SAN_CASES for kind in ('thread', 'address')]
SAN_DEFAULT = {'thread-nt64_sparse', 'thread-nt32_chunks_span_k_blocks', 'thread-nt256', 'address-nt256', 'address-k2', 'thread-k1'}


@pytest.mark.parametrize('kind,case', S.subset(SAN_RUNS, SAN_DEFAULT, id_of=lambda p: p[0] + '-' + p[1][0]))
def test_conv_tc_pipeline_under_sanitizers(kind, case):
    """ThreadSanitizer = racecheck of the four mbarrier rings (raw box, converted tile, weight stage, accumulator set) over several tiles
    per CTA and >= 3 K-blocks per tile, so that every ring wraps; AddressSanitizer = memcheck of global tensors, the weight scratch and
    the CTA's shared memory (a heap block of exactly the launch's size)."""
    exe = S.build('conv_tc_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, case[1], timeout=1500)
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum') and out.rstrip().endswith('scratch 0')


# (with dense weights every K-block closes an accumulator chunk, and the drain of K-block j-2 already orders the converter behind the MMAs
#  that read slot s: the CVT_EMPTY wait only matters when chunks span K-blocks -- the first mutant runs the sparse case)
MUTANTS = [
    ('converter-does-not-wait-for-the-mma', 'mbar_wait(BAR_CVT_EMPTY(s), ((kbc >> 1) & 1) ^ 1);', '', (1, 96, 16, 16, 32, 3, 1, 5)),
    ('drain-does-not-wait-for-the-mma', 'mbar_wait(BAR_ACC_FULL(s), (k >> 1) & 1);', '', (2, 64, 20, 16, 48, 3, 1, 0)),
    ('weight-producer-does-not-wait-for-the-mma', 'mbar_wait(BAR_W_EMPTY(ws), wph ^ 1);', '', (2, 64, 20, 16, 48, 3, 1, 0)),
]


@pytest.mark.parametrize('name,old,new,args', S.subset(MUTANTS, {'converter-does-not-wait-for-the-mma'}))
def test_the_racecheck_does_report_a_broken_pipeline(name, old, new, args):
    """Mutation check: with one wait of the pipeline removed, the ThreadSanitizer run must report a data race (or the model must abort
    on an over-arrival / deadlock) -- the run above is not vacuous.  (A consumer that merely runs AHEAD of its producer -- the MMA issuer
    without its wait for the weight stage -- reads stale data in an order the EMPTY barrier still sequences: no tool calls that a race;
    the comparisons against float64 above are what catches it.)"""
    src = _source()
    assert src.count(old) == 1, old
    exe = S.build('conv_tc_mutant_' + name.replace('-', '_'), src.replace(old, new), 'thread', SAN_MAIN)
    reported, out = S.mutant_is_reported(exe, args)
    if reported is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert reported, out[-2000:]
