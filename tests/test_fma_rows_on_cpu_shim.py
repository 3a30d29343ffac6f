"""The SOURCE of gg_fma_rows_f32 (ga-gan_b200/csrc/reduce.cu: kernel template + host launch code) compiled with g++ against a
minimal CUDA execution-model shim and executed on the CPU: blockIdx / threadIdx loops around the unmodified kernel text, the
`<<<grid, threads>>>` launch rewritten into a call of that loop.  The kernel has no shared memory, no barrier and no warp
primitive, so running its threads one after the other is a faithful execution.  This pins the launch arithmetic (grid / block
shape, the vector / scalar choice, the shared-plane stride) and the kernel's index arithmetic of a kernel that was written after
the round's GPU budget was spent; the GPU test of the same entry point is tests/test_gpu_ops.py::test_fma_vs_oracle."""
import ctypes
import os
import re
import subprocess
import tempfile

import numpy as np
import pytest

from tests.util import PKG

SHIM = r'''
#include <cmath>
#include <cstdint>
#include <cstdio>
struct dim3 { unsigned x, y, z; dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {} };
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float a, float b, float c, float d) { float4 v = {a, b, c, d}; return v; }
static dim3 blockIdx, blockDim, threadIdx, gridDim;
#define __global__
#define __restrict__
#define __launch_bounds__(...)
template <class T> static inline T __ldg(const T* p) { return *p; }
typedef void* gg_stream_t;
typedef void* cudaStream_t;
#define GG_API
#define GG_OK 0
#define GG_EINVAL (-1)
#define GG_NUM_SMS 148
static char g_err[256];
#define GG_REQUIRE(cond, ...) do { if (!(cond)) { snprintf(g_err, sizeof g_err, __VA_ARGS__); return GG_EINVAL; } } while (0)
namespace gg { static inline int check_launch(const char*) { return GG_OK; } }
static long g_blocks = 0, g_threads_per_block = 0;
// LAUNCH(kernel, grid, threads, args...): every block, every thread, one after the other
#define LAUNCH(kernel, grid, threads, ...) do { dim3 g_ = (grid); dim3 b_ = dim3(threads); gridDim = g_; blockDim = b_;              \
    g_blocks = (long)g_.x * g_.y * g_.z; g_threads_per_block = b_.x;                                                                  \
    for (unsigned by = 0; by < g_.y; ++by) for (unsigned bx = 0; bx < g_.x; ++bx) for (unsigned tx = 0; tx < b_.x; ++tx) {            \
        blockIdx = dim3(bx, by, 0); threadIdx = dim3(tx, 0, 0); kernel(__VA_ARGS__); } } while (0)
extern "C" long shim_blocks() { return g_blocks; }
extern "C" long shim_threads() { return g_threads_per_block; }
extern "C" const char* shim_error() { return g_err; }
'''


def _extract():
    src = open(os.path.join(PKG, 'csrc', 'reduce.cu')).read()
    start = src.index('// one row (= one (sample, channel) plane) per blockIdx.x')
    end = src.index('// y[r, p] = s1[r] * x1[r, p] + s2[r] * x2[r, p]')
    body = src[start:end]
    assert 'fma_rows_kernel' in body and 'gg_fma_rows_f32' in body
    body = body.replace('}  // namespace', '')                                         # the kernel lived in an anonymous namespace
    body, n = re.subn(r'(fma_rows_kernel<\w+>)<<<grid, threads, 0, st>>>\(', r'LAUNCH(\1, grid, threads, ', body)
    assert n == 2, 'expected the two launches of gg_fma_rows_f32'
    return body


@pytest.fixture(scope='module')
def shim_lib():
    with tempfile.TemporaryDirectory() as d:
        cpp = os.path.join(d, 'fma_rows_shim.cpp')
        with open(cpp, 'w') as f:
            f.write(SHIM + _extract())
        so = os.path.join(d, 'fma_rows_shim.so')
        res = subprocess.run(['g++', '-std=c++17', '-O1', '-shared', '-fPIC', '-o', so, cpp], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert res.returncode == 0, res.stdout
        lib = ctypes.CDLL(so)
        lib.gg_fma_rows_f32.restype = ctypes.c_int
        lib.gg_fma_rows_f32.argtypes = [ctypes.c_void_p] * 3 + [ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
        lib.shim_error.restype = ctypes.c_char_p
        lib.shim_blocks.restype = ctypes.c_long
        lib.shim_threads.restype = ctypes.c_long
        yield lib


def _aligned(shape, rng, offset_floats=0):
    """float32 array whose data pointer is 16-byte aligned (plus `offset_floats` floats)."""
    n = int(np.prod(shape))
    raw = np.zeros(n + 8, np.float32)
    skew = (-(raw.ctypes.data // 4)) % 4
    view = raw[skew + offset_floats: skew + offset_floats + n]
    view[:] = rng.standard_normal(n).astype(np.float32)
    return view.reshape(shape), raw


@pytest.mark.parametrize('N,C,H,W,shared,offset', [
    (3, 5, 8, 12, True, 0), (3, 5, 8, 12, False, 0),          # vectorised (P % 4 == 0, aligned), one plane / one per sample
    (2, 4, 5, 7, True, 0), (2, 4, 5, 7, False, 0),            # scalar: P = 35
    (2, 3, 8, 8, False, 1),                                   # P % 4 == 0 but x is not 16-byte aligned -> scalar
    (1, 2, 40, 40, True, 0),                                  # P4 = 400 > 256: two column chunks per row
    (2, 2, 1, 1, False, 0),                                   # one pixel per plane
])
def test_fma_rows_source_on_the_cpu(shim_lib, N, C, H, W, shared, offset):
    rng = np.random.default_rng(N * 100 + W)
    P = H * W
    x, _x = _aligned((N, C, P), rng, offset)
    s, _s = _aligned((N, C), rng)
    z, _z = _aligned((1 if shared else N, P), rng)
    y, _y = _aligned((N, C, P), rng)
    y[:] = np.nan
    rc = shim_lib.gg_fma_rows_f32(x.ctypes.data, s.ctypes.data, z.ctypes.data, 0 if shared else P, y.ctypes.data, N * C, C, P, None)
    assert rc == 0, shim_lib.shim_error()
    want = x.astype(np.float64) * s[:, :, None] + (z[0][None, None, :] if shared else z[:, None, :])
    assert not np.isnan(y).any(), 'some element was never written'
    assert np.abs(y - want).max() <= 1e-6 * max(1.0, np.abs(want).max())
    vec = P % 4 == 0 and offset == 0
    pv = P // 4 if vec else P
    threads = 256 if pv >= 256 else (pv + 31) // 32 * 32
    assert shim_lib.shim_threads() == threads and threads % 32 == 0 and 32 <= threads <= 256
    assert shim_lib.shim_blocks() == N * C * min(-(-pv // threads), 4096)


def test_fma_rows_source_rejects_bad_arguments(shim_lib):
    a, _ = _aligned((16,), np.random.default_rng(0))
    p = a.ctypes.data
    assert shim_lib.gg_fma_rows_f32(p, p, p, 3, p, 4, 2, 4, None) == -1 and b'stride' in shim_lib.shim_error()
    assert shim_lib.gg_fma_rows_f32(p, p, p, 0, p, 5, 2, 4, None) == -1 and b'bad extents' in shim_lib.shim_error()
    assert shim_lib.gg_fma_rows_f32(None, p, p, 0, p, 4, 2, 4, None) == -1 and b'null' in shim_lib.shim_error()
    assert shim_lib.gg_fma_rows_f32(p, p, p, 0, p, 0, 2, 4, None) == 0                       # empty: nothing launched, no error
