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


def test_conv2d_dispatch_of_api_cu_on_the_cpu_build():
    """gg_conv2d_f32 through plain ctypes on the CPU build: which kernel family AUTO picks (used_prec), the refusals, and that the two
    tensor-core families (row-marching / tile) give the same numbers on a shape both serve -- api.cu:100-143 executed, not restated."""
    import ctypes
    import numpy as np
    import torch
    import torch.nn.functional as F
    from tests import cpu_shim as S, emulated_lib
    lib = ctypes.CDLL(emulated_lib.build())
    P, I32 = ctypes.c_void_p, ctypes.c_int
    lib.gg_conv2d_f32.restype = I32
    lib.gg_conv2d_f32.argtypes = [P, P, P] + [I32] * 14 + [P, P, I32, ctypes.POINTER(I32), P]
    lib.gg_last_error.restype = ctypes.c_char_p
    lib.gg_set_conv_kernel_family.restype = I32
    lib.gg_set_conv_kernel_family.argtypes = [I32]

    def conv(N, I, H, W, O, K, pad, stride=1, prec=-1, skew=0, OH=None, OW=None, seed=0):
        g = torch.Generator().manual_seed(seed)
        x, w = torch.randn(N, I, H, W, generator=g), torch.randn(O, I, K, K, generator=g)
        want = F.conv2d(x.double(), w.double(), stride=stride, padding=pad).numpy()
        base = S.aligned(np.zeros(x.numel() + 4))[0]
        xs = base[skew: skew + x.numel()].reshape(x.shape)
        xs[...] = x.numpy()
        ws = S.aligned(w.numpy())[0]
        y = S.aligned(np.full(want.shape, np.nan))[0]
        used = I32(-99)
        rc = lib.gg_conv2d_f32(xs.ctypes.data, ws.ctypes.data, y.ctypes.data, N, I, H, W, O, K, K, OH or want.shape[2], OW or want.shape[3], stride, pad, pad, 0, 0, None, None,
                               prec, ctypes.byref(used), None)
        err = float(np.abs(y - want).max() / np.abs(want).max()) if rc == 0 else None
        return rc, used.value, err, y

    for args, kw, want_prec in [((1, 16, 8, 8, 16, 3, 1), {}, 3),              # tcgen05 tile kernel
                                ((1, 16, 8, 64, 32, 3, 1), {}, 3),             # row-marching kernel (OW >= 64, O <= 64)
                                ((2, 32, 8, 8, 3, 1, 0), {}, 0),               # ToRGB: the HBM-streaming thin kernel
                                ((1, 3, 8, 8, 16, 1, 0), {}, 0),               # fromRGB
                                ((1, 16, 8, 10, 16, 3, 1), {}, 0),             # W % 4 != 0: no TMA rows -> exact FFMA kernel
                                ((1, 16, 8, 8, 16, 3, 1), {'skew': 1}, 0),     # unaligned base pointer -> FFMA under AUTO
                                ((1, 8, 8, 8, 16, 3, 1), {}, 0),               # < 16 input channels
                                ((1, 16, 9, 9, 16, 3, 1), {'stride': 2}, 0),   # strided: the generic FFMA path
                                ((1, 16, 8, 8, 16, 3, 1), {'prec': 0}, 0),     # explicit FFMA
                                ((1, 16, 8, 8, 16, 3, 1), {'prec': 1}, 1)]:    # explicit one-product mode
        rc, used, err, _ = conv(*args, **kw)
        assert rc == 0 and used == want_prec, (args, kw, rc, used, lib.gg_last_error())
        assert err <= (2e-3 if want_prec == 1 else 5e-6), (args, kw, err)
    rc, _, _, _ = conv(1, 16, 8, 10, 16, 3, 1, prec=3)                          # an explicit tensor-core request for a shape that path does not serve
    assert rc == -3 and b'not served by the tcgen05 path' in lib.gg_last_error()
    rc, _, _, _ = conv(1, 16, 8, 8, 16, 3, 1, prec=99)
    assert rc == -1 and b'unknown precision mode' in lib.gg_last_error()
    rc, _, _, _ = conv(1, 16, 9, 9, 16, 3, 1, stride=2, OH=6)
    assert rc == -1 and b'output size mismatch' in lib.gg_last_error()
    _, _, _, y_march = conv(2, 32, 12, 64, 48, 3, 1, seed=3)
    assert lib.gg_set_conv_kernel_family(0) == 1
    try:
        _, used, _, y_tile = conv(2, 32, 12, 64, 48, 3, 1, seed=3)
    finally:
        lib.gg_set_conv_kernel_family(1)
    assert used == 3 and np.abs(y_march - y_tile).max() <= 2e-6 * np.abs(y_tile).max() and not np.array_equal(y_march, y_tile)      # two kernels, one answer
