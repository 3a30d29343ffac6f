"""gagan_b200 -- the B200-native operator layer of GA-GAN's StyleGAN2 hot path, and how it is dropped into a reference checkout.

The directory is called `ga-gan_b200/` (not an identifier); `import gagan_b200` works through the loader `gagan_b200.py`
at the repository root, or load this file with importlib under the package name `gagan_b200`.

    import gagan_b200
    gagan_b200.install('/path/to/GA-GAN/DissimilarDomains')     # the user's own reference checkout
    from training import networks, loss, augment                # the REFERENCE's modules, unmodified, now running on
    G = networks.Generator(...)                                 # libgagan_b200.so (tcgen05 / TMA kernels, sm_100a)

`install()` is the whole drop-in: the reference binds its hot path by module path (training/networks.py:14-19,
training/loss.py:13, training/augment.py:14-16, training/training_loop.py:25-26), so the five operator modules
`torch_utils.ops.{bias_act, upfirdn2d, conv2d_gradfix, conv2d_resample, fma}`, `torch_utils.custom_ops` and the function
`training.networks.modulated_conv2d` are replaced by this package's; everything else of the checkout (module tree, loss,
ADA pipe, persistence, training_stats, dnnlib ...) is the reference's own code and is neither copied nor shadowed.
"""
import os
import sys
import types
import importlib

__version__ = '0.2'

OPS = ('bias_act', 'upfirdn2d', 'conv2d_gradfix', 'conv2d_resample', 'fma')
_state = dict(installed=False, reference_root=None, fused_callers=False)


def _import_or_stub(name):
    """The reference's package `name` if one is importable, else an empty stand-in package (operator-only use)."""
    try:
        return importlib.import_module(name)
    except ModuleNotFoundError:
        mod = types.ModuleType(name)
        mod.__path__ = []
        mod.__doc__ = 'stand-in created by gagan_b200.install(): only the B200 operator modules live here'
        sys.modules[name] = mod
        parent, _, leaf = name.rpartition('.')
        if parent:
            setattr(sys.modules[parent], leaf, mod)
        return mod


def install(reference_root=None, fused_callers=True):
    """Make `torch_utils.ops.*`, `torch_utils.custom_ops` and `training.networks.modulated_conv2d` resolve to this build.

    reference_root   directory holding the reference's `torch_utils/`, `training/`, `dnnlib/` (a DissimilarDomains
                     checkout).  It is put on sys.path.  None: only the operator modules are registered (under a
                     stand-in `torch_utils` package unless one is importable already).
    fused_callers    also swap in the fused `forward`s of this package's training/networks.py for SynthesisLayer and
                     Conv2dLayer (same results, fewer full-tensor passes) and restore the `img is None` guard that
                     the fork's SynthesisBlock.forward lost (networks.py:1058-1063).  With False
                     the reference's forwards run untouched on the replaced operators.

    Idempotent.  Modules that imported the reference's operator modules BEFORE this call are re-bound as well.
    Returns the reference's `training.networks` module (None without a reference).
    """
    from .torch_utils import custom_ops as my_custom_ops
    mine = {name: importlib.import_module(f'{__name__}.torch_utils.ops.{name}') for name in OPS}

    if reference_root is not None:
        root = os.path.abspath(reference_root)
        for pkg in ('torch_utils', 'training', 'dnnlib'):
            if not os.path.isdir(os.path.join(root, pkg)):
                raise RuntimeError(f'gagan_b200.install: {root} is not a DissimilarDomains checkout (no {pkg}/)')
        if root not in sys.path:
            sys.path.insert(0, root)
        _state['reference_root'] = root

    _import_or_stub('torch_utils')
    ops_pkg = _import_or_stub('torch_utils.ops')

    swapped = {}                                            # id(reference module) -> replacement
    def register(fullname, mod):
        old = sys.modules.get(fullname)
        if old is not None and old is not mod:
            swapped[id(old)] = mod
        sys.modules[fullname] = mod
        parent, _, leaf = fullname.rpartition('.')
        setattr(sys.modules[parent], leaf, mod)

    for name, mod in mine.items():
        register('torch_utils.ops.' + name, mod)
    register('torch_utils.custom_ops', my_custom_ops)
    if swapped:                                             # `from torch_utils.ops import bias_act` executed earlier somewhere
        for m in list(sys.modules.values()):
            d = getattr(m, '__dict__', None)
            if not isinstance(m, types.ModuleType) or d is None:
                continue
            for key, val in list(d.items()):
                if isinstance(val, types.ModuleType) and id(val) in swapped:
                    d[key] = swapped[id(val)]

    ref_networks = None
    if _state['reference_root'] is not None:
        ref_networks = importlib.import_module('training.networks')
        src = os.path.abspath(getattr(ref_networks, '__file__', ''))
        if not src.startswith(_state['reference_root']):
            raise RuntimeError(f'gagan_b200.install: `training.networks` resolves to {src}, not to the checkout '
                               f'{_state["reference_root"]} (another `training` package is ahead of it on sys.path)')
        from .training import networks as my_networks
        my_networks.attach(ref_networks, fused_callers=fused_callers)
        _state['fused_callers'] = bool(fused_callers)
    _state['installed'] = True
    return ref_networks


def install_rosinality(model_module, fused_layers=True):
    """Bind this build's operators into a rosinality-style StyleGAN2 module -- GA-GAN's second code base,
    SimilarDomains/gan_models/StyleGAN2/model.py (`ModulatedConv2d` :176-275, `op.upfirdn2d`, `op.fused_leaky_relu`); see
    rosinality.py.  (`SimilarDomains/gan_models/StyleGAN2/nvidia.py` needs nothing of its own: it imports `torch_utils.ops.*`
    by module path, :16-21, which `install()` already serves.)"""
    from .rosinality import install_rosinality as bind
    return bind(model_module, fused_layers=fused_layers)


def installed():
    return dict(_state)
