"""TEST INFRASTRUCTURE ONLY -- the live reference (CPU, impl='ref') under private module names.

`load()` imports the oracle's byte-identical copy of the reference's hot-path packages
(oracle/_ref/DissimilarDomains/{torch_utils,training,dnnlib}, made by tools/vendor_reference.py, git-ignored) WITHOUT
leaving anything in `sys.modules`: a process that has the product installed (`gagan_b200.install`, which owns the names
`torch_utils.ops.*`) can hold the unmodified reference next to it and compare the two on the same inputs.  On CPU tensors
every reference op takes its own `impl='ref'` branch (upfirdn2d.py:172, bias_act.py:120), i.e. plain torch ops -- this is
the ground truth the parity tests and bench.py's `cpu_baseline` / `--impl reference` legs use.

The only deviation from the checked-in reference is applied from OUTSIDE, as SURVEY.md section 0.2 prescribes: the
`img is None` guard that the fork de-indented out of SynthesisBlock.forward (networks.py:1058-1063) is restored by letting
`misc.assert_shape` / `upfirdn2d.upsample2d` pass None through.  No file is edited.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; the product never does.
"""
import os
import sys
import types
import importlib
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_ROOT = os.path.join(ROOT, 'oracle', '_ref', 'DissimilarDomains')
_TOPS = ('torch_utils', 'training', 'dnnlib')
_cache = None


def available():
    return os.path.isdir(os.path.join(REF_ROOT, 'training'))


def load():
    """Namespace with the reference's modules: .networks .loss .augment .misc .training_stats .upfirdn2d .bias_act
    .conv2d_resample .conv2d_gradfix .fma .grid_sample_gradfix .dnnlib  (imported once, cached)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError(f'{REF_ROOT} is missing: run tools/vendor_reference.py where /root/reference exists')

    def is_ours(name):
        return name.split('.')[0] in _TOPS

    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if is_ours(k)}
    sys.path.insert(0, REF_ROOT)
    try:
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ns = types.SimpleNamespace()
            ns.dnnlib = importlib.import_module('dnnlib')
            ns.misc = importlib.import_module('torch_utils.misc')
            ns.persistence = importlib.import_module('torch_utils.persistence')
            # persistent_class pickles its constructor arguments as a sanity check and looks `torch_utils.persistence` up by NAME
            # in sys.modules -- the one name this private copy must not occupy.  The check guards snapshot pickling only.
            ns.persistence._check_pickleable = lambda obj: None
            ns.training_stats = importlib.import_module('torch_utils.training_stats')
            for op in ('upfirdn2d', 'bias_act', 'conv2d_resample', 'conv2d_gradfix', 'fma', 'grid_sample_gradfix'):
                setattr(ns, op, importlib.import_module('torch_utils.ops.' + op))
            # the guard of SURVEY.md section 0.2, from outside
            plain_assert, plain_up = ns.misc.assert_shape, ns.upfirdn2d.upsample2d
            ns.misc.assert_shape = lambda t, s: None if t is None else plain_assert(t, s)
            ns.upfirdn2d.upsample2d = lambda x, f, **kw: None if x is None else plain_up(x, f, **kw)
            ns.networks = importlib.import_module('training.networks')
            ns.loss = importlib.import_module('training.loss')
            ns.augment = importlib.import_module('training.augment')
        for mod in (ns.networks, ns.loss, ns.misc, ns.bias_act):
            assert os.path.abspath(mod.__file__).startswith(REF_ROOT), mod.__file__
    finally:
        sys.path.remove(REF_ROOT)
        for k in [k for k in sys.modules if is_ours(k)]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cache = ns
    return ns


SD_ROOT = os.path.join(ROOT, 'oracle', '_ref', 'SimilarDomains')
_sd_cache = None


def rosinality_available():
    return os.path.isfile(os.path.join(SD_ROOT, 'gan_models', 'StyleGAN2', 'model.py'))


def load_rosinality():
    """The reference's second StyleGAN2 implementation (SimilarDomains/gan_models/StyleGAN2/model.py with the torch-native op/ files
    its op/__init__.py:1-6 selects), unmodified, as a private module object: nothing stays in `sys.modules`, so the copy of the same
    file that a test binds the product into (gagan_b200.install_rosinality) is a different object."""
    global _sd_cache
    if _sd_cache is not None:
        return _sd_cache
    if not rosinality_available():
        raise RuntimeError(f'{SD_ROOT} is missing: run tools/vendor_reference.py where /root/reference exists')
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k.split('.')[0] == 'gan_models'}
    sys.path.insert(0, SD_ROOT)
    try:
        mod = importlib.import_module('gan_models.StyleGAN2.model')
        assert os.path.abspath(mod.__file__).startswith(SD_ROOT), mod.__file__
    finally:
        sys.path.remove(SD_ROOT)
        for k in [k for k in sys.modules if k.split('.')[0] == 'gan_models']:
            del sys.modules[k]
        sys.modules.update(saved)
    _sd_cache = mod
    return mod


def load_ga_operators():
    """GA/crossover_mutation.py of the reference (gaussian_crossover, simulated_binary_crossover, dynamic_mutation), loaded from
    oracle/_ref/GA by file path -- the GA package itself cannot be imported (GA/__init__.py:5)."""
    import importlib.util
    path = os.path.join(ROOT, 'oracle', '_ref', 'GA', 'crossover_mutation.py')
    if not os.path.isfile(path):
        return None
    spec = importlib.util.spec_from_file_location('_oracle_ref_ga_crossover_mutation', path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
