"""Two helpers the operator modules share (kept here so that they do not depend on the reference's `torch_utils.misc`)."""
import functools
import torch


def check_dims(t, dims, what='tensor'):
    """Raise unless `t` has exactly `len(dims)` dimensions whose sizes equal the non-None entries of `dims`."""
    got = tuple(int(s) for s in t.shape)
    ok = len(got) == len(dims) and all(d is None or int(d) == g for d, g in zip(dims, got))
    if not ok:
        raise AssertionError(f'{what}: shape {list(got)} does not match {list(dims)}')


def scoped(fn):
    """Run `fn` inside a torch profiler range named after it (the reference labels its ops the same way, so traces
    of the two line up: `modulated_conv2d`, `conv2d_resample`, ...)."""
    @functools.wraps(fn)
    def run(*args, **kwargs):
        with torch.autograd.profiler.record_function(fn.__name__):
            return fn(*args, **kwargs)
    return run
