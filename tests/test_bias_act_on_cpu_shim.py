"""The whole translation unit ga-gan_b200/csrc/bias_act.cu -- the C-ABI entry points gg_bias_act_f32 / gg_bias_act_noise_f32, their
argument checks, the launch code and the three kernel templates, unmodified -- compiled with g++ against tests/cuda_cpu_shim.h and
executed on the CPU (one std::thread per CUDA thread; warp_sum's shuffles and the bias-gradient atomics emulated), against the oracle,
and under ThreadSanitizer / AddressSanitizer with exact-size tensors: the CPU stand-in for `compute-sanitizer` racecheck / memcheck,
which is closed on the GPU pool."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from tests.util import PKG, ROOT
from oracle import ops_ref as R

ACTS = {1: 'linear', 2: 'relu', 3: 'lrelu', 4: 'tanh', 5: 'sigmoid', 6: 'elu', 7: 'selu', 8: 'softplus', 9: 'swish'}


def _source():
    common = open(os.path.join(PKG, 'csrc', 'common.cuh')).read()
    warp_sum = common[common.index('__device__ __forceinline__ float warp_sum(float v) {'):]
    warp_sum = warp_sum[:warp_sum.index('\n}\n') + 3]
    src = open(os.path.join(PKG, 'csrc', 'bias_act.cu')).read()
    body = src[src.index('#include "common.cuh"') + len('#include "common.cuh"'):]
    body, n = re.subn(r'(\w+<[^;]*?>)<<<\(unsigned\)grid, 256, 0, st>>>\(p\);', r'SHIM_LAUNCH((\1), (unsigned)grid, 256, p);', body)
    assert n == 3, 'expected the three launches of bias_act.cu'
    return '#include "cuda_cpu_shim.h"\nnamespace gg {\n' + warp_sum + '}\n' + body


SAN_MAIN = r'''
#include <cstdlib>
#include <cstring>
static float* tensor(size_t n, float scale) {            // exact-size, 16-byte aligned: the sanitizer's red zone starts behind element n-1
    float* p = (float*)aligned_alloc(16, (n * 4 + 15) / 16 * 16);
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N C P act  -- forward (+noise), then grad 1 with the fused bias gradient, dense NCHW (stepB = P) and channels-last (stepB = 1)
    const int N = atoi(argv[1]), C = atoi(argv[2]), P = atoi(argv[3]), act = atoi(argv[4]);
    const size_t n = (size_t)N * C * P;
    float *x = tensor(n, 2.f), *b = tensor(C, .5f), *noise = tensor((size_t)N * P, .3f), *y = tensor(n, 0.f), *dy = tensor(n, 1.f), *dx = tensor(n, 0.f),
          *db = tensor(C, 0.f);
    int rc = 0;
    rc |= gg_bias_act_f32(x, b, nullptr, nullptr, nullptr, y, nullptr, 0, act, .2f, 1.4f, 1.5f, (int64_t)n, C, P, nullptr);
    rc |= gg_bias_act_noise_f32(x, b, noise, P, y, act, .2f, 1.4f, -1.f, (int64_t)n, C, P, nullptr);
    rc |= gg_bias_act_noise_f32(x, b, noise, 0, y, act, .2f, 1.4f, -1.f, (int64_t)n, C, P, nullptr);
    memset(db, 0, C * 4);
    rc |= gg_bias_act_f32(dy, b, x, y, nullptr, dx, db, 1, act, .2f, 1.4f, 1.5f, (int64_t)n, C, P, nullptr);
    rc |= gg_bias_act_f32(dy, b, x, y, dx, dx, db, 2, act, .2f, 1.4f, 1.5f, (int64_t)n, C, P, nullptr);
    rc |= gg_bias_act_f32(x, b, nullptr, nullptr, nullptr, y, nullptr, 0, act, .2f, 1.4f, -1.f, (int64_t)n, C, 1, nullptr);      // bias along the innermost dim
    rc |= gg_bias_act_f32(dy, b, x, y, nullptr, dx, db, 1, act, .2f, 1.4f, -1.f, (int64_t)n, C, 1, nullptr);
    double s = 0; for (size_t i = 0; i < n; ++i) s += y[i] + dx[i];
    printf("rc %d checksum %.5f\n", rc, s);
    free(x); free(b); free(noise); free(y); free(dy); free(dx); free(db);
    return rc;
}
'''

_built = {}


def _build(kind):
    if kind in _built:
        return _built[kind]
    d = tempfile.mkdtemp(prefix='bias_act_shim_')
    flags = ['g++', '-std=c++20', '-O1', '-pthread', '-w', '-I', os.path.join(ROOT, 'tests')]
    cpp = os.path.join(d, 'bias_act_shim.cpp')
    if kind == 'lib':
        open(cpp, 'w').write(_source())
        out = os.path.join(d, 'bias_act_shim.so')
        cmd = flags + ['-shared', '-fPIC', '-o', out, cpp]
    else:
        open(cpp, 'w').write(_source() + SAN_MAIN)
        out = os.path.join(d, 'bias_act_' + kind)
        cmd = flags + ['-g', '-fno-omit-frame-pointer', '-fsanitize=' + kind, '-o', out, cpp]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]
    _built[kind] = out
    return out


@pytest.fixture(scope='module')
def lib():
    so = ctypes.CDLL(_build('lib'))
    P, I64, I, F = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
    so.gg_bias_act_f32.restype = I
    so.gg_bias_act_f32.argtypes = [P] * 7 + [I, I, F, F, F, I64, I, I64, P]
    so.gg_bias_act_noise_f32.restype = I
    so.gg_bias_act_noise_f32.argtypes = [P] * 3 + [I64, P, I, F, F, F, I64, I, I64, P]
    so.shim_error.restype = ctypes.c_char_p
    return so


def _aligned(a):
    a = np.ascontiguousarray(a, np.float32)
    raw = np.zeros(a.size + 8, np.float32)
    skew = (-(raw.ctypes.data // 4)) % 4
    view = raw[skew: skew + a.size].reshape(a.shape)
    view[...] = a
    return view, raw


def _ptr(a):
    return None if a is None else a.ctypes.data


@pytest.mark.parametrize('act', sorted(ACTS))
@pytest.mark.parametrize('shape,dim', [((2, 8, 16, 16), 1), ((3, 5, 6, 7), 1), ((4, 12), 1), ((2, 6, 4, 32), 1)], ids=['uniform', 'scalar', 'fc', 'vec'])
def test_bias_act_source_forward_and_gradient_on_the_cpu(lib, act, shape, dim):
    """Forward, first-order gradient with the fused bias reduction, and (where the activation has one) the second-order kernel, for
    the three kernel variants: vector with a warp-uniform channel ([2,8,16,16]: 256-element planes), vector with per-lane channels
    ([2,6,4,32]), scalar ([3,5,6,7], and the [4,12] fully-connected case whose bias runs along the innermost dim)."""
    name = ACTS[act]
    alpha_, gain_, has2 = R.ACTIVATIONS[name][0], float(R.ACTIVATIONS[name][1]), R.ACTIVATIONS[name][4]
    g = torch.Generator().manual_seed(act * 10 + len(shape))
    xt, bt, dyt = torch.randn(shape, generator=g), torch.randn(shape[dim], generator=g) * 0.5, torch.randn(shape, generator=g)
    clamp = 1.2
    n, C = xt.numel(), shape[dim]
    step = int(np.prod(shape[dim + 1:]))
    x, _a = _aligned(xt.numpy()); b, _b = _aligned(bt.numpy()); dy, _c = _aligned(dyt.numpy())
    y, _d = _aligned(np.full(shape, np.nan)); dx, _e = _aligned(np.full(shape, np.nan)); db, _f = _aligned(np.zeros(C))
    assert lib.gg_bias_act_f32(_ptr(x), _ptr(b), None, None, None, _ptr(y), None, 0, act, alpha_, gain_, clamp, n, C, step, None) == 0, lib.shim_error()
    want = R.bias_act(xt.double(), bt.double(), dim=dim, act=name, clamp=clamp).numpy()
    assert np.abs(y - want).max() <= 2e-6 * max(1.0, np.abs(want).max())
    assert lib.gg_bias_act_f32(_ptr(dy), _ptr(b), _ptr(x), _ptr(y), None, _ptr(dx), _ptr(db), 1, act, alpha_, gain_, clamp, n, C, step, None) == 0
    xr = xt.double() + bt.double().reshape([-1 if i == dim else 1 for i in range(len(shape))])
    want_dx = R.bias_act_grad_formula(1, name, dyt.double(), xr, torch.from_numpy(y.astype(np.float64)), None, alpha_, gain_, clamp).numpy()
    assert np.abs(dx - want_dx).max() <= 2e-6 * max(1.0, np.abs(want_dx).max())
    want_db = want_dx.sum(axis=tuple(i for i in range(len(shape)) if i != dim))
    assert np.abs(db - want_db).max() <= 1e-5 * max(1.0, np.abs(want_db).max())
    if has2:
        d2, _g = _aligned(np.full(shape, np.nan))
        assert lib.gg_bias_act_f32(_ptr(dy), _ptr(b), _ptr(x), _ptr(y), _ptr(dx), _ptr(d2), None, 2, act, alpha_, gain_, clamp, n, C, step, None) == 0
        want_d2 = R.bias_act_grad_formula(2, name, dyt.double(), xr, torch.from_numpy(y.astype(np.float64)), torch.from_numpy(dx.astype(np.float64)),
                                          alpha_, gain_, clamp).numpy()
        assert np.abs(d2 - want_d2).max() <= 5e-6 * max(1.0, np.abs(want_d2).max())


@pytest.mark.parametrize('shared', [True, False])
@pytest.mark.parametrize('shape', [(2, 8, 16, 16), (3, 5, 6, 7)], ids=['vec', 'scalar'])
def test_bias_act_noise_source_on_the_cpu(lib, shape, shared):
    N, C, H, W = shape
    g = torch.Generator().manual_seed(H)
    xt, bt = torch.randn(shape, generator=g), torch.randn(C, generator=g) * 0.5
    nt = torch.randn((1 if shared else N, 1, H, W), generator=g) * 0.3
    x, _a = _aligned(xt.numpy()); b, _b = _aligned(bt.numpy()); nz, _c = _aligned(nt.numpy()); y, _d = _aligned(np.full(shape, np.nan))
    assert lib.gg_bias_act_noise_f32(_ptr(x), _ptr(b), _ptr(nz), 0 if shared else H * W, _ptr(y), 3, 0.2, 1.4, -1.0, xt.numel(), C, H * W, None) == 0
    want = R.bias_act((xt + nt).double(), bt.double(), act='lrelu', gain=1.4).numpy()
    assert np.abs(y - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize('kind', ['thread', 'address'])
@pytest.mark.parametrize('args', [(2, 8, 256, 3), (2, 6, 128, 3), (3, 5, 42, 3), (1, 4, 128, 9)], ids=['uniform', 'vec', 'scalar', 'swish'])
def test_bias_act_translation_unit_under_sanitizers(kind, args):
    """Forward (+ per-sample and shared noise), gradient with the fused bias reduction (warp shuffles + atomics), second order, dense
    and bias-innermost layouts, with every CUDA thread a real thread: ThreadSanitizer sees the shuffle exchange and the atomics,
    AddressSanitizer every global access against exact-size tensors."""
    try:
        exe = _build(kind)
    except AssertionError as e:
        pytest.skip(f'g++ -fsanitize={kind} is not available: ' + str(e)[-200:])
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=0 exitcode=66', ASAN_OPTIONS='detect_leaks=0')
    res = subprocess.run([exe] + [str(v) for v in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=900)
    if 'FATAL: ThreadSanitizer' in res.stdout and 'data race' not in res.stdout:
        pytest.skip('ThreadSanitizer cannot run in this container')
    assert 'data race' not in res.stdout and 'AddressSanitizer' not in res.stdout, res.stdout[-4000:]
    assert res.returncode == 0 and 'rc 0' in res.stdout, res.stdout[-2000:]
