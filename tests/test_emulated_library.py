"""The product's operator layer on a CPU BUILD OF THE LIBRARY, in the default CPU suite: a child process runs a slice of the GPU suite
(`tests/test_gpu_ops.py`) with GG_DRYRUN=emu, i.e. torch_utils.ops.* -> the product's own ctypes wrappers (custom_ops._Plugin) ->
the C ABI of include/gagan_b200.h -> api.cu's checks and kernel-family dispatch -> every kernel's shipped source, the tcgen05 ones on
the model of mbarriers / TMA / tensor memory / tcgen05.mma (tests/emulated_lib.py, tests/tc_cpu_shim.h).  No stand-in kernel, no oracle
in the path under test.  The slice holds the reference's own golden vectors for all five operators (tests/golden/*.npz: upfirdn2d,
bias_act with first and second order, conv2d_resample incl. grouped and up / down, modulated_conv2d fused / non-fused / up / ToRGB),
the fp32-faithfulness and zero-block-skipping checks of the tensor-core path, the per-sample scales, float16 modulated_conv2d, the fma
entry point and the operators' error behaviour."""
import os
import re
import subprocess
import sys

from tests.util import ROOT

SLICE = ('fma_vs_oracle or upfirdn2d_golden or bias_act_golden or modulated_conv2d_golden or conv2d_resample_golden or conv2d_tc_zero_block or '
         'upfirdn2d_errors or bias_act_empty_and_errors or conv2d_scales_fused or conv2d_tc_is_fp32_faithful or float16_modulated')


def test_gpu_suite_slice_passes_on_the_cpu_build_of_the_library():
    env = dict(os.environ, GG_DRYRUN='emu')
    env.pop('GG_SANITIZE_ALL', None)
    res = subprocess.run([sys.executable, '-m', 'pytest', os.path.join(ROOT, 'tests', 'test_gpu_ops.py'), '-m', 'gpu', '-q', '-p', 'no:cacheprovider', '-k', SLICE],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env, cwd=ROOT, timeout=1500)
    tail = res.stdout[-3000:]
    assert res.returncode == 0, tail
    m = re.search(r'(\d+) passed', tail)
    assert m and int(m.group(1)) == 17 and 'failed' not in tail and 'skipped' not in tail, tail


def test_the_cpu_build_exports_the_abi_of_the_header():
    """Every symbol the product's loader declares (custom_ops._SIGNATURES == include/gagan_b200.h) exists in the CPU build, the reported
    version and device check are those of api.cu, and argument validation is api.cu's own (no kernel runs for a rejected call)."""
    import ctypes
    from tests import emulated_lib
    from torch_utils import custom_ops
    lib = ctypes.CDLL(emulated_lib.build())
    for name in custom_ops.EXPORTED_SYMBOLS:
        assert hasattr(lib, name), name
    assert lib.gg_version() == 100 and lib.gg_device_ok() == 1
    lib.gg_last_error.restype = ctypes.c_char_p
    P = ctypes.c_void_p
    lib.gg_chan_dot_f32.argtypes = [P, P, P, ctypes.c_int64, ctypes.c_int64, P]
    assert lib.gg_chan_dot_f32(None, None, None, 1, 1, None) == -1 and b'null pointer' in lib.gg_last_error()
