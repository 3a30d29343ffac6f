import os
import sys
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import tests.util  # noqa: E402,F401  (installs the drop-in: gagan_b200.install(baseline/_ref/DissimilarDomains))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (sm_100a) device; run with -m gpu on the GPU box')


@pytest.fixture(scope='session')
def device():
    import torch
    if os.environ.get('GG_DRYRUN'):
        # Authoring aid for a box without a GPU: `GG_DRYRUN=1 pytest tests -m gpu` runs the PYTHON paths of the GPU tests on CPU
        # tensors with every kernel replaced by the torch stand-in (tests/fake_plugin.py).  Tests that probe kernel-level behaviour
        # (argument validation of the C ABI wrappers, precision modes, launch counters, torch.cuda calls) fail in this mode by
        # design; everything else must pass -- it catches host-layer and test-code regressions before GPU time is spent.
        # GG_DRYRUN=2 (3, 4, ...): the stand-in additionally sums its channels in a permuted order (seed = the value): the same
        # mathematics with another fp32 rounding, which is what the GPU kernels are to a tolerance written on the CPU.
        # GG_DRYRUN=emu: no stand-in at all -- the product's own ctypes wrappers on a CPU BUILD OF THE LIBRARY (tests/emulated_lib.py:
        # api.cu's dispatch and every kernel's source, the tcgen05 ones on the hardware model of tests/tc_cpu_shim.h).  Slow; for the
        # small shapes.  Kernel-level behaviour (argument validation, precision modes, launch counters) is real in this mode.
        from tests.fake_plugin import FakePlugin
        from torch_utils.ops import conv2d_gradfix as cg, bias_act as BA, upfirdn2d as U, fma as FM
        from torch_utils import custom_ops
        if os.environ['GG_DRYRUN'] == 'emu':
            from tests import emulated_lib
            fp = emulated_lib.bind()
        else:
            level = int(os.environ['GG_DRYRUN'])
            fp = FakePlugin(permute_seed=(level if level > 1 else None))
        cg._plugin = fp; BA._plugin = fp; U._plugin = fp
        for name in ('bias_act_plugin', 'upfirdn2d_plugin', 'conv2d_plugin'):
            custom_ops._cached_plugins[name] = fp
        def dtype_only(t):                                   # the entry checks minus "must be a CUDA tensor"
            if os.environ['GG_DRYRUN'] == 'emu' and t.dtype != torch.float32:
                raise RuntimeError('this build serves fp32 kernels; other dtypes are out of scope')
        for mod in (cg, U, BA):
            mod._check_input = dtype_only
        FM._on_device = lambda a: True
        return torch.device('cpu')
    if not torch.cuda.is_available():
        pytest.skip('no CUDA device')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    return torch.device('cuda:0')
