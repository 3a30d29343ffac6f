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


def _compile(name, source, kind, main):
    """One g++ run; the result is kept on disk under a name derived from the source text, so a repeated session compiles nothing."""
    import hashlib
    text = source if kind == 'lib' else source + main
    hdrs = b''.join(open(os.path.join(ROOT, 'tests', h), 'rb').read() for h in ('cuda_cpu_shim.h', 'tc_cpu_shim.h'))
    d = os.path.join(tempfile.gettempdir(), 'gagan_shim_' + hashlib.sha256(text.encode() + hdrs + kind.encode()).hexdigest()[:20])
    out = os.path.join(d, name + ('.so' if kind == 'lib' else '_' + kind))
    if os.path.isfile(out):
        return out
    os.makedirs(d, exist_ok=True)
    cpp = os.path.join(d, name + '.cpp')
    open(cpp, 'w').write(text)
    flags = ['g++', '-std=c++20', '-O1', '-pthread', '-w', '-I', os.path.join(ROOT, 'tests')]
    tmp = out + '.tmp%d' % os.getpid()
    cmd = flags + (['-shared', '-fPIC'] if kind == 'lib' else ['-g', '-fno-omit-frame-pointer', '-fsanitize=' + kind]) + ['-o', tmp, cpp]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-3000:]
    os.replace(tmp, out)
    return out


def build(name, source, kind, main=''):
    """kind: 'lib' (shared object for ctypes) or a -fsanitize= value ('thread', 'address': an executable from source + main)."""
    key = (name, kind, hash(source), hash(main) if kind != 'lib' else 0)
    if key not in _built:
        _built[key] = _compile(name, source, kind, main)
    return _built[key]


def build_all(name, source, main):
    """The three builds of a unit (library, ThreadSanitizer, AddressSanitizer executables) side by side: the compiles dominate the
    wall time of these tests."""
    from concurrent.futures import ThreadPoolExecutor
    with ThreadPoolExecutor(max_workers=3) as ex:
        list(ex.map(lambda kind: build(name, source, kind, main), ('lib', 'thread', 'address')))


def run_sanitized(exe, args, timeout=900):
    """Run a sanitizer build; returns its output, or None if the sanitizer cannot run in this container."""
    env = dict(os.environ, TSAN_OPTIONS='halt_on_error=0 exitcode=66 history_size=7', ASAN_OPTIONS='detect_leaks=0')
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


def _balanced(text, i, open_ch, close_ch):
    """Index just behind the bracket that closes the one at text[i]."""
    assert text[i] == open_ch
    depth = 0
    for j in range(i, len(text)):
        if text[j] == open_ch:
            depth += 1
        elif text[j] == close_ch:
            depth -= 1
            if depth == 0:
                return j + 1
    raise AssertionError('unbalanced ' + open_ch)


def _split_top(text):
    """Split at commas that are outside every bracket."""
    parts, depth, cur = [], 0, ''
    for ch in text:
        if ch in '([{':
            depth += 1
        elif ch in ')]}':
            depth -= 1
        if ch == ',' and depth == 0:
            parts.append(cur.strip())
            cur = ''
        else:
            cur += ch
    parts.append(cur.strip())
    return parts


def translate_unit(cu_text, expect_launches, helpers=True):
    """A whole .cu translation unit for the shim: everything behind its `#include "common.cuh"`, with EVERY
    `kernel<T...><<<grid, block, smem, stream>>>(args...)` rewritten into SHIM_LAUNCH((kernel<T...>), grid, block, args...) by a
    bracket-matching pass (any grid / block expression, any argument list), `extern __shared__` bound to the shim's buffer.
    Nothing else of the source is touched: entry points, argument checks, launch arithmetic and kernels are the shipped text."""
    body = cu_text[cu_text.index('#include "common.cuh"') + len('#include "common.cuh"'):]
    out, pos, n = '', 0, 0
    while True:
        k = body.find('<<<', pos)
        if k < 0:
            break
        # the kernel expression in front of <<<: an identifier, optionally with one template argument list
        j = k
        if body[j - 1] == '>':
            depth = 0
            while True:
                j -= 1
                if body[j] == '>':
                    depth += 1
                elif body[j] == '<':
                    depth -= 1
                    if depth == 0:
                        break
        m = re.search(r'[\w:]+$', body[:j])
        start = m.start()
        kernel = body[start:k]
        e = body.index('>>>', k)
        cfg = _split_top(body[k + 3:e])
        assert len(cfg) in (2, 3, 4), cfg
        a0 = e + 3
        assert body[a0] == '(', body[a0:a0 + 20]
        a1 = _balanced(body, a0, '(', ')')
        out += body[pos:start] + f'SHIM_LAUNCH(({kernel}), {cfg[0]}, {cfg[1]}, {body[a0 + 1:a1 - 1]})'
        pos = a1
        n += 1
    out += body[pos:]
    assert n == expect_launches, f'expected {expect_launches} kernel launches, rewrote {n}'
    out = re.sub(r'extern __shared__ float (\w+)\[\];', r'float* \1 = shim_dynamic_smem;', out)
    return '#include "cuda_cpu_shim.h"\n' + (device_helpers() if helpers else '') + out


def load(so):
    import ctypes
    lib = ctypes.CDLL(so)
    lib.shim_blocks.restype = ctypes.c_long
    lib.shim_blocks_since_reset.restype = ctypes.c_long
    lib.shim_threads.restype = ctypes.c_long
    lib.shim_error.restype = ctypes.c_char_p
    return lib


def fptr(a):
    import ctypes
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float)) if a is not None else None


def tc_common_tail():
    """The PTX-free part of csrc/tc_common.cuh, verbatim: instruction descriptor, the 3xTF32 split, rz_compensation, the fused epilogue."""
    t = open(os.path.join(CSRC, 'tc_common.cuh')).read()
    tail = t[t.index('// UMMA shared-memory descriptor, SWIZZLE_NONE'):t.index("// host: point this translation unit's watchdog")]
    assert 'asm' not in tail
    return 'namespace ggtc {\n' + tail + '}  // namespace ggtc\n'


def translate_tc_unit(cu_text, expect_launches):
    """A tcgen05 translation unit for tests/tc_cpu_shim.h: everything behind its `#include "tc_common.cuh"`.  Rewritten: the `<<<>>>`
    launches (the dynamic shared-memory size becomes the size of the CTA's heap block), `extern __shared__`, and the unit's own inline
    PTX -- `setmaxnreg` (dropped: register allocation has no CPU counterpart), `bar.sync id, n` (the shim's named barrier) and
    `ld.shared.v4.f32` (a plain 16-byte read).  Everything else -- warp roles, pipelines, descriptors, index arithmetic, host code -- is
    the shipped text."""
    body = cu_text[cu_text.index('#include "tc_common.cuh"') + len('#include "tc_common.cuh"'):]
    body = re.sub(r'asm volatile\("setmaxnreg\.[a-z]+\.sync\.aligned\.u32 \d+;"\);', '', body)
    body = re.sub(r'asm volatile\("bar\.sync %0, %1;" ::"r"\((\w+)\), "r"\((\w+)\) : "memory"\);', r'shim_named_barrier(\1, \2);', body)
    body = re.sub(r'asm volatile\("ld\.shared\.v4\.f32 \{%0, %1, %2, %3\}, \[%4\];" : "=f"\((\w+)\.x\), "=f"\(\1\.y\), "=f"\(\1\.z\), "=f"\(\1\.w\) : "r"\(smem_u32\((\w+)\)\)\);',
                  r'\1 = *reinterpret_cast<const float4*>(\2);', body)
    assert 'asm' not in body, body[body.index('asm') - 200: body.index('asm') + 200]
    out, pos, n = '', 0, 0
    while True:
        k = body.find('<<<', pos)
        if k < 0:
            break
        j = k
        if body[j - 1] == '>':
            j = _rfind_template_open(body, j)
        start = re.search(r'[\w:]+$', body[:j]).start()
        kernel = body[start:k]
        e = body.index('>>>', k)
        cfg = _split_top(body[k + 3:e])
        a0 = e + 3
        a1 = _balanced(body, a0, '(', ')')
        launch = f'SHIM_LAUNCH(({kernel}), {cfg[0]}, {cfg[1]}, {body[a0 + 1:a1 - 1]})'
        if len(cfg) > 2 and cfg[2] != '0':
            launch = f'(shim_set_smem({cfg[2]}), {launch})'
        out += body[pos:start] + launch
        pos = a1
        n += 1
    out += body[pos:]
    assert n == expect_launches, f'expected {expect_launches} kernel launches, rewrote {n}'
    out, m = re.subn(r'extern __shared__ __align__\(\d+\) uint8_t (\w+)\[\];', r'uint8_t* \1 = shim_tc_smem;', out)
    assert m >= 1
    out = out.replace('#include <cuda.h>', '')
    return '#include "tc_cpu_shim.h"\n' + device_helpers() + tc_common_tail() + out


def _rfind_template_open(body, j):
    depth = 0
    while True:
        j -= 1
        if body[j] == '>':
            depth += 1
        elif body[j] == '<':
            depth -= 1
            if depth == 0:
                return j


FULL = os.environ.get('GG_SANITIZE_ALL') == '1'


def subset(params, default_ids, id_of=lambda p: p[0]):
    """pytest params for a list of cases: the ones named in `default_ids` always run, the others only with GG_SANITIZE_ALL=1 (the whole
    matrix takes ~10 minutes of host time; the default CPU suite keeps one representative per kernel and mechanism)."""
    import pytest
    out = []
    for p in params:
        i = id_of(p)
        marks = [] if (FULL or i in default_ids) else [pytest.mark.skip(reason='runs with GG_SANITIZE_ALL=1')]
        out.append(pytest.param(*p, id=i, marks=marks) if isinstance(p, tuple) else pytest.param(p, id=i, marks=marks))
    return out


def mutant_is_reported(exe, args, attempts=({}, {'SHIM_JITTER': '100'}, {'SHIM_SLOW_WARPS': '0-3:200'}, {'SHIM_JITTER': '500'}, {'SHIM_SLOW_WARPS': '4-19:200'})):
    """Run a mutant (one wait / barrier removed) under ThreadSanitizer until one run reports it: a data race, or the model's abort on an
    over-arrival / deadlock.  Whether the two unordered accesses of a removed wait meet inside the sanitizer's history depends on the
    schedule, so the later attempts perturb it (tests/tc_cpu_shim.h: SHIM_JITTER, SHIM_SLOW_WARPS).  Returns (reported, last output);
    reported is None when the sanitizer runtime cannot start in this container."""
    out = ''
    for knobs in attempts:
        env = dict(os.environ, TSAN_OPTIONS='halt_on_error=1 exitcode=66 history_size=7', SHIM_WAIT_TIMEOUT_S='8', **knobs)
        res = subprocess.run([exe] + [str(v) for v in args], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, timeout=900)
        out = res.stdout
        if 'FATAL: ThreadSanitizer' in out and 'data race' not in out:
            return None, out
        if res.returncode != 0 and ('data race' in out or 'TC SHIM ABORT' in out):
            return True, out
    return False, out
