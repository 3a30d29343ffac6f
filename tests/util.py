"""Shared test helpers (import-safe on a box with no GPU and no /root/reference)."""
import os
import sys
import contextlib
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'ga-gan_b200')
GOLDEN = os.path.join(ROOT, 'tests', 'golden')
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# The drop-in under test: the reference checkout that travels with the repo (tools/vendor_reference.py -> baseline/_ref,
# git-ignored, byte-identical to /root/reference) with this build's operators installed into it.  Without the checkout only
# the operator modules are registered and every test that needs the reference's networks skips.
CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
HAVE_CHECKOUT = os.path.isdir(os.path.join(CHECKOUT, 'training'))

import gagan_b200  # noqa: E402

gagan_b200.install(CHECKOUT if HAVE_CHECKOUT else None)


def reference_networks():
    """The reference's own `training.networks` module running on the library (skips if the checkout did not travel)."""
    import pytest
    if not HAVE_CHECKOUT:
        pytest.skip('baseline/_ref/DissimilarDomains is absent (run tools/vendor_reference.py where /root/reference exists)')
    import training.networks
    return training.networks


SD_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'SimilarDomains')


def rosinality_model(fused_layers=True):
    """SimilarDomains/gan_models/StyleGAN2/model.py from the checkout that travels with the repo, with this build bound into it
    (gagan_b200.install_rosinality); skips if the file did not travel."""
    import pytest
    if not os.path.isfile(os.path.join(SD_CHECKOUT, 'gan_models', 'StyleGAN2', 'model.py')):
        pytest.skip('baseline/_ref/SimilarDomains is absent (run tools/vendor_reference.py where /root/reference exists)')
    if SD_CHECKOUT not in sys.path:
        sys.path.insert(0, SD_CHECKOUT)
    import importlib
    model = importlib.import_module('gan_models.StyleGAN2.model')
    return gagan_b200.install_rosinality(model, fused_layers=fused_layers)


def quiet(fn, *args, **kwargs):
    """Call `fn` with stdout swallowed (the reference prints one line per layer when it registers domain modulation)."""
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*args, **kwargs)

TOL = 1e-3   # BASELINE.json north_star: max relative error 1e-3 vs the reference in fp32 / TF32 off


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'), allow_pickle=False)
    return {k: d[k] for k in d.files}


def t(a, device='cpu', requires_grad=False):
    if a is None:
        return None
    return torch.as_tensor(np.asarray(a), dtype=torch.float32).to(device).requires_grad_(requires_grad)


def max_rel_err(a, b):
    """max|a-b| / max|b| (the north star's metric)."""
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(b).detach().double().cpu()
    assert a.shape == b.shape, (a.shape, b.shape)
    d = b.abs().max().item()
    e = (a - b).abs().max().item() if a.numel() else 0.0
    return e if d == 0 else e / d


def assert_close(a, b, tol=TOL, what=''):
    e = max_rel_err(a, b)
    assert e <= tol, f'{what}: max-rel-err {e:.3e} > {tol:.1e}'
    return e


@contextlib.contextmanager
def patched_randn(seed):
    """Route torch.randn / torch.randn_like through ONE seeded CPU generator.

    The reference (on CPU, when goldens are made), the oracle and the product
    (on the GPU) then see identical noise as long as they draw in the same order.
    """
    g = torch.Generator().manual_seed(seed)
    orig_randn, orig_like = torch.randn, torch.randn_like

    def randn(*size, device=None, dtype=None, generator=None, **_):
        if len(size) == 1 and isinstance(size[0], (list, tuple, torch.Size)):
            size = tuple(size[0])
        out = orig_randn(tuple(int(s) for s in size), generator=g, dtype=dtype or torch.float32)
        return out.to(device) if device is not None else out

    def randn_like(x, **_):
        return randn(*x.shape, device=x.device, dtype=x.dtype)

    torch.randn, torch.randn_like = randn, randn_like
    try:
        yield
    finally:
        torch.randn, torch.randn_like = orig_randn, orig_like


@contextlib.contextmanager
def patched_rand(seed):
    """patched_randn plus torch.rand through the same seeded CPU generator (the ADA pipe draws its apply / skip masks, flips and
    angles with torch.rand on the device)."""
    orig_rand = torch.rand
    with patched_randn(seed):
        g = torch.Generator().manual_seed(seed + 1)

        def rand(*size, device=None, dtype=None, generator=None, **_):
            if len(size) == 1 and isinstance(size[0], (list, tuple, torch.Size)):
                size = tuple(size[0])
            out = orig_rand(tuple(int(s_) for s_ in size), generator=g, dtype=dtype or torch.float32)
            return out.to(device) if device is not None else out
        torch.rand = rand
        try:
            yield
        finally:
            torch.rand = orig_rand
