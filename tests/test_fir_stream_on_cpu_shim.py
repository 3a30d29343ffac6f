"""The SOURCE of the unit-rate FIR kernel family fir_stream<0, PX0, STG> (ga-gan_b200/csrc/upfirdn2d.cu: StreamP, fir_march, fir_stream,
launch_stream2 / launch_stream -- kernel AND launch code, unmodified) compiled with g++ against tests/cuda_cpu_shim.h and executed on
the CPU, one std::thread per CUDA thread, with real barriers for __syncthreads / __syncwarp.

Why: the staged variant (rows that are not 16-byte multiples travel through a per-warp shared-memory exchange) ships with thresholds of
256 / 384 columns; its GPU tests at those widths (tests/test_gpu_ops.py::test_upfirdn2d_wide_unaligned_rows_vs_oracle) were written
after the round's GPU budget was spent.  Here the very same source runs those shapes -- rows spanning several warps, warps spanning
rows and planes, parked lanes of the grid's last warp, vector and scalar stores -- against the oracle, and the build with
-fsanitize=thread checks the __syncwarp discipline of the exchange buffers (a missing barrier is a data race between the lane threads).
tests/test_fir_stage_model.py is the numpy model of the same index arithmetic; this file executes the code itself."""
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

DRIVER = r'''
extern "C" int run_fir_stream(const float* x, const float* f, float* y, int N, int C, int inH, int inW, int padx0, int pady0, int flip,
                              float gain, int outH, int outW) {
    // gg_upfirdn2d_f32, unit-rate branch (upfirdn2d.cu): StreamP sp{...}; if (stream_ok(sp, false)) return launch_stream<0>(sp, st);
    StreamP sp{x, f, y, N, C, inH, inW, padx0, pady0, flip, gain, outH, outW, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (padx0 < 0 || padx0 > 3) return -2;
    return launch_stream<0>(sp, nullptr);
}
'''


def _source():
    src = open(os.path.join(PKG, 'csrc', 'upfirdn2d.cu')).read()
    start = src.index('struct StreamP {')
    end = src.index('// fir_resample2<UP, PX0>: the true 2x cases of the 4x4 filter')
    end = src.rindex('// ----', start, end)
    body = src[start:end]
    body, n = re.subn(r'fir_stream<([^;]*?)><<<\(unsigned\)blocks, 128, 0, st>>>\(p\);', r'SHIM_LAUNCH((fir_stream<\1>), (unsigned)blocks, 128, p);', body)
    assert n == 2, 'expected the two launches of launch_stream2'
    return '#include "cuda_cpu_shim.h"\n' + body + DRIVER


_built = {}


def _build(tsan, sanitizer='thread'):
    key = (tsan, sanitizer)
    if key not in _built:
        _built[key] = _build_once(tsan, sanitizer)
    return _built[key]


def _build_once(tsan, sanitizer):
    d = tempfile.mkdtemp(prefix='fir_shim_')
    cpp = os.path.join(d, 'fir_stream_shim.cpp')
    with open(cpp, 'w') as f:
        f.write(_source())
    out = os.path.join(d, 'fir_stream_' + sanitizer if tsan else 'fir_stream_shim.so')
    flags = ['-std=c++20', '-O1', '-pthread', '-w', '-I', os.path.join(ROOT, 'tests')]
    if tsan:
        cpp_main = os.path.join(d, 'main.cpp')
        with open(cpp_main, 'w') as f:
            f.write(_source() + TSAN_MAIN)
        cmd = ['g++'] + flags + ['-g', '-fno-omit-frame-pointer', '-fsanitize=' + sanitizer, '-o', out, cpp_main]
    else:
        cmd = ['g++'] + flags + ['-shared', '-fPIC', '-o', out, cpp]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]
    return out


@pytest.fixture(scope='module')
def shim_lib():
    lib = ctypes.CDLL(_build(tsan=False))
    lib.run_fir_stream.restype = ctypes.c_int
    lib.run_fir_stream.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int] * 7 + [ctypes.c_float, ctypes.c_int, ctypes.c_int]
    lib.shim_error.restype = ctypes.c_char_p
    lib.shim_blocks.restype = ctypes.c_long
    return lib


def _aligned(shape, fill=None):
    n = int(np.prod(shape))
    raw = np.zeros(n + 8, np.float32)
    skew = (-(raw.ctypes.data // 4)) % 4
    view = raw[skew: skew + n].reshape(shape)
    if fill is not None:
        view[...] = fill
    return view, raw


CASES = [
    # name, N, C, inH, inW, padding [x0, x1, y0, y1], flip, gain
    ('filt_p1_257', 1, 2, 257, 257, [1, 1, 1, 1], False, 1.0),            # the four shapes of test_upfirdn2d_wide_unaligned_rows_vs_oracle
    ('filt_p2_384', 1, 2, 384, 384, [2, 2, 2, 2], False, 1.0),
    ('filt_both_odd_wide', 1, 1, 301, 515, [1, 1, 1, 1], True, 4.0),
    ('filt_p1_1025', 1, 1, 70, 1025, [1, 1, 2, 0], False, 1.0),
    ('aligned_both', 2, 1, 40, 264, [2, 1, 1, 2], False, 1.0),            # 16-byte rows on both sides: the plain vector kernel (STG = false)
    ('narrow_unaligned', 1, 3, 21, 131, [3, 0, 0, 3], True, 1.0),         # below the thresholds: per-lane scalar loads, padx0 = 3
    ('padx0_0_wide', 1, 1, 19, 390, [0, 3, 3, 0], False, 2.0),            # 390 -> 390 columns, both unaligned, padx0 = 0
    ('tiny', 1, 1, 5, 9, [1, 1, 1, 1], False, 1.0),
]


@pytest.mark.parametrize('case', CASES, ids=lambda c: c[0])
def test_fir_stream_source_on_the_cpu(shim_lib, case):
    name, N, C, H, W, pad, flip, gain = case
    g = torch.Generator().manual_seed(H * 1000 + W)
    xt = torch.randn(N, C, H, W, generator=g)
    f = R.setup_filter([1, 3, 3, 1])
    want = R.upfirdn2d(xt.double(), f, padding=pad, flip_filter=flip, gain=gain).numpy()
    outH, outW = want.shape[2:]
    x, _x = _aligned((N, C, H, W), xt.numpy())
    fa, _f = _aligned((4, 4), f.numpy())
    y, _y = _aligned((N, C, outH, outW), np.nan)
    rc = shim_lib.run_fir_stream(x.ctypes.data, fa.ctypes.data, y.ctypes.data, N, C, H, W, pad[0], pad[2], int(flip), gain, outH, outW)
    assert rc == 0, shim_lib.shim_error()
    assert not np.isnan(y).any(), f'{name}: some output element was never written'
    assert np.abs(y - want).max() <= 2e-6 * max(1.0, np.abs(want).max()), name
    # nothing written outside the output tensor (the exchange stores go through per-lane row pointers)
    assert (_y[:((-(_y.ctypes.data // 4)) % 4)] == 0).all() and (_y[((-(_y.ctypes.data // 4)) % 4) + y.size:] == 0).all()


TSAN_MAIN = r'''
#include <cstdlib>
int main(int argc, char** argv) {
    // argv: N C inH inW padx0 padx1 pady0 pady1 flip
    int a[9]; for (int i = 0; i < 9; ++i) a[i] = atoi(argv[1 + i]);
    const int N = a[0], C = a[1], H = a[2], W = a[3], outH = H + a[6] + a[7] - 3, outW = W + a[4] + a[5] - 3;
    // exact-size, 16-byte aligned heap tensors: under AddressSanitizer the red zones start right behind the last element
    const size_t nx = (size_t)N * C * H * W, ny = (size_t)N * C * outH * outW;
    float* xa = (float*)aligned_alloc(16, (nx * 4 + 15) / 16 * 16);
    float* ya = (float*)aligned_alloc(16, (ny * 4 + 15) / 16 * 16);
    float* f = (float*)aligned_alloc(16, 64);
    for (size_t i = 0; i < nx; ++i) xa[i] = (float)((i * 2654435761u) % 1000) / 500.f - 1.f;
    for (size_t i = 0; i < ny; ++i) ya[i] = 0.f;
    for (int i = 0; i < 16; ++i) f[i] = (float)((i % 4 == 0 || i % 4 == 3 ? 1 : 3) * (i / 4 == 0 || i / 4 == 3 ? 1 : 3)) / 64.f;
    int rc = run_fir_stream(xa, f, ya, N, C, H, W, a[4], a[6], a[8], 1.f, outH, outW);
    double s = 0; for (size_t i = 0; i < ny; ++i) s += ya[i];
    printf("rc %d checksum %.6f blocks %ld\n", rc, s, shim_blocks());
    free(xa); free(ya); free(f);
    return rc;
}
'''


@pytest.mark.parametrize('shape', [
    (1, 1, 12, 259, 1, 1, 1, 1, 0),      # 259 -> 258 columns, both unaligned: input window + output exchange on one buffer, 2 warps per row
    (1, 2, 10, 388, 0, 3, 2, 1, 1),      # 388 -> 388 columns, both aligned: the plain vector kernel (no exchange; the block barrier only)
    (1, 1, 9, 392, 1, 0, 3, 0, 0),       # aligned input (392), 390 unaligned output columns: output exchange only, lanes on different rows
], ids=['in+out', 'aligned', 'out-only'])
def test_fir_stream_exchange_buffers_are_race_free_under_thread_sanitizer(shape):
    """The staged kernel's shared-memory discipline (write, __syncwarp, read, __syncwarp -- for the input window and again for the output
    exchange that reuses the same buffer) with every lane a real thread, under ThreadSanitizer: any access pair that is not ordered by
    a barrier is reported and fails the test."""
    try:
        exe = _build(tsan=True)
    except AssertionError as e:                                 # no libtsan on this box
        pytest.skip('g++ -fsanitize=thread is not available: ' + str(e)[-200:])
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=0 report_signal_unsafe=0 exitcode=66')
    res = subprocess.run([exe] + [str(v) for v in shape], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
    if 'FATAL: ThreadSanitizer' in res.stdout and 'data race' not in res.stdout:
        pytest.skip('ThreadSanitizer cannot run in this container: ' + res.stdout[-200:])
    assert 'data race' not in res.stdout, res.stdout[-4000:]
    assert res.returncode == 0 and 'rc 0' in res.stdout, res.stdout[-2000:]


@pytest.mark.parametrize('shape', [
    (1, 1, 12, 259, 1, 1, 1, 1, 0), (1, 2, 10, 388, 0, 3, 2, 1, 1), (1, 1, 9, 392, 1, 0, 3, 0, 0), (2, 1, 7, 131, 3, 0, 0, 3, 1),
    (1, 1, 35, 1025, 1, 1, 2, 0, 0), (1, 3, 4, 8, 2, 1, 1, 2, 0), (1, 1, 33, 257, 1, 1, 1, 1, 0),
], ids=['in+out', 'aligned', 'out-only', 'narrow', 'w1025', 'tiny-aligned', 'w257'])
def test_fir_stream_touches_no_byte_outside_its_tensors_under_address_sanitizer(shape):
    """The memcheck of this kernel family: input, filter and output are exact-size heap blocks, every global load and store of every
    thread is instrumented -- the guarded halo loads, the 128-bit loads at the row ends, the parked lanes of the last warp, the
    per-lane row pointers of the output exchange."""
    try:
        exe = _build(tsan=True, sanitizer='address')
    except AssertionError as e:
        pytest.skip('g++ -fsanitize=address is not available: ' + str(e)[-200:])
    res = subprocess.run([exe] + [str(v) for v in shape], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         env=dict(os.environ, ASAN_OPTIONS='detect_leaks=0'), timeout=600)
    assert 'AddressSanitizer' not in res.stdout, res.stdout[-4000:]
    assert res.returncode == 0 and 'rc 0' in res.stdout, res.stdout[-2000:]


def test_thread_sanitizer_sees_a_removed_syncwarp():
    """The checker checks: with the barrier between the coalesced window write and the per-lane read taken out of the source, the same
    run must report a data race."""
    old = "            for (int i = 0; i < 9; ++i) sb.v[32 * i + lane] = buf[i];\n            __syncwarp();"
    src = _source() + TSAN_MAIN
    assert src.count(old) == 1
    d = tempfile.mkdtemp(prefix='fir_shim_mut_')
    with open(os.path.join(d, 'm.cpp'), 'w') as f:
        f.write(src.replace(old, old[:old.rindex('\n')]))
    exe = os.path.join(d, 'm')
    res = subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-w', '-g', '-fsanitize=thread', '-I', os.path.join(ROOT, 'tests'), '-o', exe,
                          os.path.join(d, 'm.cpp')], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        pytest.skip('g++ -fsanitize=thread is not available')
    run = subprocess.run([exe] + '1 1 12 259 1 1 1 1 0'.split(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         env=dict(os.environ, TSAN_OPTIONS='halt_on_error=1 exitcode=66'), timeout=600)
    if 'FATAL: ThreadSanitizer' in run.stdout and 'data race' not in run.stdout:
        pytest.skip('ThreadSanitizer cannot run in this container')
    assert 'data race' in run.stdout


def test_address_sanitizer_sees_a_widened_guard():
    """The checker checks: with the column guard of the coalesced input window widened by one (`col <= inW`), the run at a width whose
    last row ends exactly at the end of the tensor must report a heap-buffer-overflow."""
    old = "in[i] = (row_ok && col >= 0 && col < p.inW) ? __ldg(roww + col) : 0.f;"
    src = _source() + TSAN_MAIN
    assert src.count(old) == 1
    d = tempfile.mkdtemp(prefix='fir_shim_mut_')
    with open(os.path.join(d, 'm.cpp'), 'w') as f:
        f.write(src.replace(old, old.replace('col < p.inW', 'col <= p.inW')))
    exe = os.path.join(d, 'm')
    res = subprocess.run(['g++', '-std=c++20', '-O1', '-pthread', '-w', '-g', '-fsanitize=address', '-I', os.path.join(ROOT, 'tests'), '-o', exe,
                          os.path.join(d, 'm.cpp')], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if res.returncode != 0:
        pytest.skip('g++ -fsanitize=address is not available')
    run = subprocess.run([exe] + '1 1 12 259 1 1 1 1 0'.split(), stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                         env=dict(os.environ, ASAN_OPTIONS='detect_leaks=0'), timeout=600)
    assert 'heap-buffer-overflow' in run.stdout, run.stdout[-2000:]
