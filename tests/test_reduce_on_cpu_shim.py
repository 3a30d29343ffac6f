"""The whole translation unit ga-gan_b200/csrc/reduce.cu -- gg_chan_dot_f32, gg_chan_dot_preact_f32, gg_scale_rows_f32,
gg_axpby_rows_f32, gg_fma_rows_f32: entry points, argument checks, launch arithmetic and kernels, unmodified -- compiled with g++
against tests/cuda_cpu_shim.h and executed on the CPU (one std::thread per CUDA thread; warp_sum's shuffles, the block reduction
through shared memory and the cross-CTA atomics emulated), against float64 numpy, and under ThreadSanitizer / AddressSanitizer with
exact-size tensors: the CPU stand-in for `compute-sanitizer` racecheck / memcheck, which is closed on the GPU pool (DESIGN.md
section 2).  What these kernels compute: the style / demodulation-coefficient gradients of modulated_conv2d
(reference training/networks.py:642-651 under autograd) and the per-(sample, channel) scalings of its second-order pass."""
import ctypes
import os

import numpy as np
import pytest

from tests import cpu_shim as S

SAN_MAIN = r'''
#include <cstdlib>
static float* tensor(size_t n, float scale) {            // exact-size, 16-byte aligned: the sanitizer's red zone starts behind element n-1
    float* p = (float*)aligned_alloc(16, (n * 4 + 15) / 16 * 16);
    for (size_t i = 0; i < n; ++i) p[i] = scale * ((float)((i * 2654435761u) % 2001) / 1000.f - 1.f);
    return p;
}
int main(int argc, char** argv) {
    // argv: N C P skew  -- every entry point of the unit on [N*C, P] tensors; skew = 1 moves the operands off 16-byte alignment (scalar variants)
    const int N = atoi(argv[1]), C = atoi(argv[2]), P = atoi(argv[3]), skew = atoi(argv[4]);
    const size_t rows = (size_t)N * C, n = rows * P;
    float *a0 = tensor(n + skew, 1.f), *b0 = tensor(n + skew, .7f), *y0 = tensor(n + skew, 0.f), *s1 = tensor(rows, 1.3f), *s2 = tensor(rows, .4f),
          *out = tensor(rows, 0.f), *bias = tensor(C, .2f), *noise = tensor((size_t)N * P + skew, .1f);
    float *a = a0 + skew, *b = b0 + skew, *y = y0 + skew;
    int rc = 0;
    rc |= gg_chan_dot_f32(a, b, out, rows, P, nullptr);
    rc |= gg_scale_rows_f32(a, s1, y, rows, P, nullptr);
    rc |= gg_axpby_rows_f32(a, s1, b, s2, y, rows, P, nullptr);
    rc |= gg_fma_rows_f32(a, s1, noise + skew, P, y, rows, C, P, nullptr);
    rc |= gg_fma_rows_f32(a, s1, noise + skew, 0, y, rows, C, P, nullptr);
    if (!skew && P % 4 == 0) {
        rc |= gg_chan_dot_preact_f32(a, b, bias, noise, P, out, N, C, P, 3, .2f, 1.4f, nullptr);
        rc |= gg_chan_dot_preact_f32(a, b, nullptr, noise, 0, out, N, C, P, 1, 0.f, 1.f, nullptr);
    }
    double s = 0; for (size_t i = 0; i < n; ++i) s += y[i]; for (size_t i = 0; i < rows; ++i) s += out[i];
    printf("rc %d checksum %.5f\n", rc, s);
    free(a0); free(b0); free(y0); free(s1); free(s2); free(out); free(bias); free(noise);
    return rc;
}
'''


def _source():
    return S.translate_unit(open(os.path.join(S.CSRC, 'reduce.cu')).read(), expect_launches=8)


@pytest.fixture(scope='module', autouse=True)
def _prebuilt():
    S.build_all('reduce_unit', _source(), SAN_MAIN)


@pytest.fixture(scope='module')
def lib():
    so = S.load(S.build('reduce_unit', _source(), 'lib'))
    P, I64, I, F = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float
    so.gg_chan_dot_f32.restype = I
    so.gg_chan_dot_f32.argtypes = [P, P, P, I64, I64, P]
    so.gg_chan_dot_preact_f32.restype = I
    so.gg_chan_dot_preact_f32.argtypes = [P, P, P, P, I64, P, I, I, I64, I, F, F, P]
    so.gg_scale_rows_f32.restype = I
    so.gg_scale_rows_f32.argtypes = [P, P, P, I64, I64, P]
    so.gg_axpby_rows_f32.restype = I
    so.gg_axpby_rows_f32.argtypes = [P, P, P, P, P, I64, I64, P]
    return so


def _t(rng, shape, skew=0):
    """float32 tensor whose first element sits `skew` floats behind a 16-byte boundary."""
    n = int(np.prod(shape))
    base, keep = S.aligned(np.zeros(n + 4))
    v = base[skew: skew + n].reshape(shape)
    v[...] = rng.standard_normal(shape)
    return v, keep


def _p(a):
    return None if a is None else a.ctypes.data


# rows, P, skew: vector path with one chunk / several chunks (atomics; the four-deep unrolled loop needs > 768 float4 per CTA),
# scalar path by misalignment and by P % 4 != 0, a row count over the grid.y limit is exercised by the host loop test below
@pytest.mark.parametrize('rows,P,skew', [(6, 256, 0), (2, 16384, 0), (3, 8200, 0), (5, 1021, 0), (4, 1024, 1), (2, 12291, 0), (7, 4, 0), (3, 1, 0)],
                         ids=['vec', 'vec-4deep-atomics', 'vec-atomics-ragged', 'scalar-odd', 'scalar-unaligned', 'scalar-atomics', 'one-float4', 'P1'])
def test_chan_dot_source_on_the_cpu(lib, rows, P, skew):
    rng = np.random.default_rng(rows * 1000 + P)
    a, _a = _t(rng, (rows, P), skew)
    b, _b = _t(rng, (rows, P), skew)
    out = np.full(rows, np.nan, np.float32)
    lib.shim_reset()
    assert lib.gg_chan_dot_f32(_p(a), _p(b), _p(out), rows, P, None) == 0, lib.shim_error()
    want = (a.astype(np.float64) * b.astype(np.float64)).sum(1)
    scale = np.sqrt((a.astype(np.float64) ** 2).sum(1) * (b.astype(np.float64) ** 2).sum(1))
    assert np.abs(out - want).max() <= 2e-6 * scale.max()
    chunks = max(1, min(-(-6 * 148 // rows), -(-P // 4096)))
    assert lib.shim_blocks_since_reset() == chunks * rows and lib.shim_threads() == 256


def test_chan_dot_source_degenerate_extents_and_checks(lib):
    out = np.full(3, np.nan, np.float32)
    a = np.ones((3, 8), np.float32)
    assert lib.gg_chan_dot_f32(_p(a), _p(a), _p(out), 3, 0, None) == 0 and (out == 0).all()        # empty planes: zeros, no launch
    out[:] = 7
    assert lib.gg_chan_dot_f32(_p(a), _p(a), _p(out), 0, 8, None) == 0 and (out == 7).all()        # no rows: nothing written
    assert lib.gg_chan_dot_f32(None, _p(a), _p(out), 3, 8, None) == -1 and b'null' in lib.shim_error()
    assert lib.gg_chan_dot_f32(_p(a), _p(a), _p(out), 1 << 31, 8, None) == -1 and b'too large' in lib.shim_error()
    for act, alpha in ((2, 0.0), (3, 0.0), (9, 0.2)):                                                # only invertible activations
        assert lib.gg_chan_dot_preact_f32(_p(a), _p(a), None, None, 0, _p(out), 1, 3, 8, act, alpha, 1.0, None) == -1 and b'invertible' in lib.shim_error()
    assert lib.gg_chan_dot_preact_f32(_p(a), _p(a), None, None, 0, _p(out), 1, 3, 6, 1, 0.0, 1.0, None) == -1       # planes of 6: not float4 groups
    assert lib.gg_chan_dot_preact_f32(_p(a), _p(a), None, _p(a), 5, _p(out), 1, 3, 8, 1, 0.0, 1.0, None) == -1      # noise stride is 0 or P


@pytest.mark.parametrize('N,C,P,act,noise,bias', [(2, 3, 64, 3, 'sample', True), (2, 3, 64, 1, 'shared', True), (1, 5, 16, 3, None, False),
                                                  (1, 2, 16384, 3, 'shared', True), (3, 2, 8, 1, None, True)],
                         ids=['lrelu-noise-per-sample', 'linear-shared-noise', 'plain', 'atomics', 'bias-only'])
def test_chan_dot_preact_source_on_the_cpu(lib, N, C, P, act, noise, bias):
    """out[n,c] = sum_p ds * (act^-1(y) - bias[c] - noise[n,p]): the demodulation gradient reconstructed from the SAVED OUTPUT of a
    fused conv + bias + noise + activation."""
    rng = np.random.default_rng(N * 100 + C * 10 + P)
    alpha, gain = 0.2, float(np.sqrt(2))
    pre = rng.standard_normal((N, C, P))
    bz = rng.standard_normal(C) * 0.5 if bias else np.zeros(C)
    nz = None if noise is None else rng.standard_normal((N, P) if noise == 'sample' else (1, P)) * 0.3
    full = pre + bz[None, :, None] + (0 if nz is None else nz[:, None, :])
    yv = (np.where(full > 0, full, alpha * full) if act == 3 else full) * gain
    ds, _a = _t(rng, (N, C, P))
    y, _b = S.aligned(yv)
    bt = S.aligned(bz)[0] if bias else None
    nt = None if nz is None else S.aligned(nz)[0]
    out = np.full(N * C, np.nan, np.float32)
    assert lib.gg_chan_dot_preact_f32(_p(ds), _p(y), _p(bt), _p(nt), 0 if noise != 'sample' else P, _p(out), N, C, P, act, alpha, gain, None) == 0, lib.shim_error()
    want = (ds.astype(np.float64) * pre).sum(2).reshape(-1)
    scale = np.sqrt((ds.astype(np.float64) ** 2).sum(2) * (pre ** 2).sum(2)).max()
    assert np.abs(out - want).max() <= 5e-6 * scale


@pytest.mark.parametrize('rows,P,skew', [(6, 64, 0), (5, 7, 0), (3, 64, 1), (1, 5000, 0), (300, 4, 0)], ids=['vec', 'scalar-odd', 'scalar-unaligned', 'one-row', 'many-rows'])
def test_scale_rows_and_axpby_rows_source_on_the_cpu(lib, rows, P, skew):
    rng = np.random.default_rng(rows * 31 + P)
    x1, _a = _t(rng, (rows, P), skew)
    x2, _b = _t(rng, (rows, P), skew)
    s1 = rng.standard_normal(rows).astype(np.float32)
    s2 = rng.standard_normal(rows).astype(np.float32)
    y, _c = _t(rng, (rows, P), skew)
    assert lib.gg_scale_rows_f32(_p(x1), _p(s1), _p(y), rows, P, None) == 0, lib.shim_error()
    assert np.array_equal(y, s1[:, None] * x1)                                    # one fp32 product per element: exact
    assert lib.gg_axpby_rows_f32(_p(x1), _p(s1), _p(x2), _p(s2), _p(y), rows, P, None) == 0, lib.shim_error()
    want = s1[:, None].astype(np.float64) * x1 + (s2[:, None] * x2).astype(np.float64)       # fmaf(s1, x1, fl(s2 * x2))
    assert np.abs(y - want).max() <= 1.5e-7 * np.abs(want).max() + 1e-30
    assert lib.gg_scale_rows_f32(_p(x1), _p(s1), _p(y), 0, P, None) == 0 and lib.gg_axpby_rows_f32(_p(x1), _p(s1), _p(x2), _p(s2), _p(y), rows, 0, None) == 0
    assert lib.gg_scale_rows_f32(_p(x1), None, _p(y), rows, P, None) == -1


@pytest.mark.parametrize('kind', ['thread', 'address'])
@pytest.mark.parametrize('args', [(2, 3, 64, 0), (1, 2, 16384, 0), (1, 1, 20000, 0), (2, 2, 37, 0), (1, 3, 64, 1)], ids=['vec', 'atomics', 'ragged-last-chunk', 'scalar-odd', 'scalar-unaligned'])
def test_reduce_translation_unit_under_sanitizers(kind, args):
    """ThreadSanitizer: the shared-memory block reduction (red[8] between __syncthreads) and the cross-CTA float atomics are race-free.
    AddressSanitizer: with tensors of exactly rows x P floats no kernel touches a byte outside them -- incl. the four-deep unrolled
    loop's look-ahead guard (`q + 3*256 < q1`) and the ragged last chunk."""
    exe = S.build('reduce_unit', _source(), kind, SAN_MAIN)
    out = S.run_sanitized(exe, args)
    if out is None:
        pytest.skip('the sanitizer runtime cannot start in this container')
    assert out.startswith('rc 0 checksum')


@pytest.mark.parametrize('kind,old,new,args,report', [
    ('address', 'for (; q + 3 * 256 < q1; q += 4 * 256) {', 'for (; q < q1; q += 4 * 256) {', (1, 1, 20000, 0), 'AddressSanitizer'),
    ('thread', "red[threadIdx.x >> 5] = s;\n    __syncthreads();", "red[threadIdx.x >> 5] = s;", (2, 3, 64, 0), 'data race'),
], ids=['look-ahead-guard-removed', 'barrier-removed'])
def test_the_sanitizer_runs_do_report_a_broken_kernel(kind, old, new, args, report):
    """Mutation check of the two runs above: the unrolled loop without its look-ahead guard reads past the tensor (AddressSanitizer
    reports it), the block reduction without its barrier is a data race (ThreadSanitizer reports it)."""
    src = _source()
    assert src.count(old) >= 1
    exe = S.build('reduce_mutant_' + kind, src.replace(old, new, 1), kind, SAN_MAIN)
    import subprocess
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=1 exitcode=66 history_size=7', ASAN_OPTIONS='detect_leaks=0')
    for attempt in range(4):                # whether ThreadSanitizer sees the two unordered accesses meet depends on the schedule
        res = subprocess.run([exe] + [str(v) for v in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=600)
        if 'FATAL: ThreadSanitizer' in res.stdout and 'data race' not in res.stdout:
            pytest.skip('the sanitizer runtime cannot start in this container')
        if report in res.stdout and res.returncode != 0:
            return
    assert False, res.stdout[-2000:]
