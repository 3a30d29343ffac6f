"""Helpers for executing CUDA translation units on the CPU through tests/cuda_cpu_shim.h (test infrastructure)."""
import os
import re
import subprocess
import tempfile

import numpy as np

from tests.util import PKG, ROOT

CSRC = os.path.join(PKG, 'csrc')


def device_helpers():
    """gg::warp_sum / floordiv / posmod of common.cuh, verbatim, inside `namespace gg`."""
    common = open(os.path.join(CSRC, 'common.cuh')).read()
    text = common[common.index('__device__ __forceinline__ float warp_sum(float v) {'):common.index('}  // namespace gg')]
    return 'namespace gg {\n' + text + '}\n'


def translate(cu_text, expect_launches):
    """A .cu file's text after its `#include "common.cuh"`, with every `kernel<<<grid, block, smem, st>>>(args);` rewritten into
    SHIM_LAUNCH((kernel), grid, block, args); and dynamic shared memory bound to the shim's buffer.  Nothing else is touched."""
    body = cu_text[cu_text.index('#include "common.cuh"') + len('#include "common.cuh"'):]
    body, n = re.subn(r'(\w+(?:<[^<>;]*>)?)<<<(\(unsigned\)\w+), (\d+), [^;]*?, st>>>\(([^;]*?)\);', r'SHIM_LAUNCH((\1), \2, \3, \4);', body)
    assert n == expect_launches, f'expected {expect_launches} kernel launches, rewrote {n}'
    body = re.sub(r'extern __shared__ float (\w+)\[\];', r'float* \1 = shim_dynamic_smem;', body)
    assert '<<<' not in body
    return '#include "cuda_cpu_shim.h"\n' + device_helpers() + body


_built = {}


def build(name, source, kind, main=''):
    """kind: 'lib' (shared object for ctypes) or a -fsanitize= value ('thread', 'address': an executable from source + main)."""
    key = (name, kind)
    if key in _built:
        return _built[key]
    d = tempfile.mkdtemp(prefix=name + '_shim_')
    cpp = os.path.join(d, name + '.cpp')
    flags = ['g++', '-std=c++20', '-O1', '-pthread', '-w', '-I', os.path.join(ROOT, 'tests')]
    if kind == 'lib':
        open(cpp, 'w').write(source)
        out = os.path.join(d, name + '.so')
        cmd = flags + ['-shared', '-fPIC', '-o', out, cpp]
    else:
        open(cpp, 'w').write(source + main)
        out = os.path.join(d, name + '_' + kind)
        cmd = flags + ['-g', '-fno-omit-frame-pointer', '-fsanitize=' + kind, '-o', out, cpp]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]
    _built[key] = out
    return out


def run_sanitized(exe, args, timeout=900):
    """Run a sanitizer build; returns its output, or None if the sanitizer cannot run in this container."""
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=0 exitcode=66', ASAN_OPTIONS='detect_leaks=0')
    res = subprocess.run([exe] + [str(v) for v in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=timeout)
    if 'FATAL: ThreadSanitizer' in res.stdout and 'data race' not in res.stdout:
        return None
    assert 'data race' not in res.stdout and 'AddressSanitizer' not in res.stdout, res.stdout[-4000:]
    assert res.returncode == 0, res.stdout[-2000:]
    return res.stdout


def aligned(a):
    """(16-byte aligned float32 copy of `a`, the owning buffer -- keep it alive)."""
    a = np.ascontiguousarray(a, np.float32)
    raw = np.zeros(a.size + 8, np.float32)
    skew = (-(raw.ctypes.data // 4)) % 4
    view = raw[skew: skew + a.size].reshape(a.shape)
    view[...] = a
    return view, raw
