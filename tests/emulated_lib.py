"""TEST INFRASTRUCTURE: libgagan_b200 built FOR THE CPU -- every translation unit of ga-gan_b200/csrc (api.cu with its argument checks and
kernel-family dispatch, the SIMT kernels, the four tcgen05 kernels), unmodified except for the launch syntax and the units' own inline
PTX (tests/cpu_shim.py), compiled with g++ against tests/cuda_cpu_shim.h / tests/tc_cpu_shim.h and linked into one shared object that
exports the SAME C ABI as the product library (include/gagan_b200.h).  `pointers` are host pointers, the stream argument is ignored.

Use: `GG_DRYRUN=emu python -m pytest tests -m gpu -k ...` (tests/conftest.py) binds the product's own ctypes wrappers
(ga-gan_b200/torch_utils/custom_ops.py::_Plugin -- argument checks, output allocation, precision bookkeeping) to this library and runs
the GPU suite's python paths on CPU tensors: operator layer -> wrappers -> C ABI -> dispatch -> kernel source on the hardware model.
It is slow (one std::thread per CUDA thread) and meant for the small shapes; it measures nothing."""
import hashlib
import os
import subprocess
import tempfile
from concurrent.futures import ThreadPoolExecutor

from tests import cpu_shim as S
from tests.util import ROOT

SIMT_UNITS = {'bias_act.cu': 3, 'reduce.cu': 8, 'conv_thin.cu': 3, 'conv_simt.cu': 2, 'upfirdn2d.cu': 8}
TC_UNITS = {'conv_tc.cu': 2, 'conv_march.cu': 2, 'wgrad_tma.cu': 1, 'wgrad_tc.cu': 1}
NUM_SMS = 6          # persistent kernels start min(tiles, NUM_SMS) CTAs of 384 threads each

API_STUBS = r'''
#define cudaHostAllocMapped 0
#define cudaHostAllocPortable 0
#define cudaDevAttrComputeCapabilityMajor 75
#define cudaDevAttrComputeCapabilityMinor 76
static inline cudaError_t cudaHostAlloc(void** p, size_t, int) { *p = nullptr; return 1; }       // no watchdog record on the CPU
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
extern "C" {       // the prototypes of include/gagan_b200.h that api.cu itself calls
int gg_bias_act_f32(const float* x, const float* b, const float* xref, const float* yref, const float* dy, float* y, float* dbias, int grad, int act, float alpha,
                    float gain, float clamp, int64_t sizeX, int sizeB, int64_t stepB, gg_stream_t stream);
int gg_bias_act_noise_f32(const float* x, const float* b, const float* noise, int64_t noise_batch_stride, float* y, int act, float alpha, float gain, float clamp,
                          int64_t sizeX, int sizeB, int64_t stepB, gg_stream_t stream);
int gg_conv2d_wgrad_pm_f32(const float* a, const float* b, float* dw, int N, int A, int HA, int WA, int B, int HB, int WB, int KH, int KW, int stride, int pad_y, int pad_x,
                           int flip_w, int out_layout, const float* a_scale, const float* b_scale, int prec, int* used_prec, int pm_dim, unsigned pm_dead, gg_stream_t stream);
}
static inline cudaError_t cudaDeviceGetAttribute(int* v, int attr, int) { *v = attr == cudaDevAttrComputeCapabilityMajor ? 10 : 0; return cudaSuccess; }   // "sm_100"
'''


def _unit_source(name):
    text = open(os.path.join(S.CSRC, name)).read()
    head = f'#define SHIM_MULTI_UNIT 1\n#define GG_NUM_SMS {NUM_SMS}\n'
    if name in SIMT_UNITS:
        return head + S.translate_unit(text, SIMT_UNITS[name])
    if name in TC_UNITS:
        return head + S.translate_tc_unit(text, TC_UNITS[name])
    assert name == 'api.cu'
    body = text[text.index('#include "tc_common.cuh"') + len('#include "tc_common.cuh"'):]
    assert '<<<' not in body and 'asm' not in body
    return head + '#include "tc_cpu_shim.h"\n' + API_STUBS + S.tc_common_tail() + body


def build():
    """Path of the CPU build of the library (cached by the hash of its sources)."""
    units = sorted(list(SIMT_UNITS) + list(TC_UNITS) + ['api.cu'])
    sources = {u: _unit_source(u) for u in units}
    h = hashlib.sha256()
    for u in units:
        h.update(sources[u].encode())
    for hdr in ('cuda_cpu_shim.h', 'tc_cpu_shim.h'):
        h.update(open(os.path.join(ROOT, 'tests', hdr), 'rb').read())
    d = os.path.join(tempfile.gettempdir(), 'gagan_emulated_' + h.hexdigest()[:16])
    so = os.path.join(d, 'libgagan_b200_emulated.so')
    if os.path.isfile(so):
        return so
    os.makedirs(d, exist_ok=True)
    flags = ['g++', '-std=c++20', '-O2', '-pthread', '-w', '-fPIC', '-I', os.path.join(ROOT, 'tests')]

    def compile_unit(u):
        cpp, obj = os.path.join(d, u[:-3] + '.cpp'), os.path.join(d, u[:-3] + '.o')
        open(cpp, 'w').write(sources[u])
        res = subprocess.run(flags + ['-c', '-o', obj, cpp], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        assert res.returncode == 0, f'{u}:\n' + res.stdout[-4000:]
        return obj

    with ThreadPoolExecutor(max_workers=min(10, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_unit, units))
    tmp = so + '.tmp'
    res = subprocess.run(['g++', '-shared', '-pthread', '-o', tmp] + objs, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0, res.stdout[-4000:]
    os.replace(tmp, so)
    return so


def bind():
    """Point the product's ctypes layer at the CPU build and relax what cannot hold for CPU tensors (is_cuda, the stream handle, the
    device context manager); returns the _Plugin object the operator modules are given."""
    import contextlib
    import torch
    from torch_utils import custom_ops
    custom_ops.LIB_PATH = build()
    custom_ops._lib = None
    lib = custom_ops.load_library()            # declares every prototype of include/gagan_b200.h on the emulated library

    def require(t, name):
        if not isinstance(t, torch.Tensor):
            raise RuntimeError(f'{name} must be a tensor')
        if t.dtype != torch.float32:
            raise RuntimeError(f'{name} must be float32 (this build serves the fp32 path; fp16/fp64 are out of scope)')
    custom_ops._require_cuda = require
    custom_ops._check_device = lambda t: None
    custom_ops._stream = lambda t: None
    real_device = torch.cuda.device
    torch.cuda.device = lambda dev: contextlib.nullcontext() if torch.device(dev).type == 'cpu' else real_device(dev)
    return custom_ops._Plugin(lib)
