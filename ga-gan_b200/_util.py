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


def to_f32(v):
    """float16 tensors (also inside an `epilogue` dict) -> float32 copies; everything else unchanged."""
    if isinstance(v, torch.Tensor):
        return v.float() if v.dtype == torch.float16 else v
    if isinstance(v, dict):
        return {k: to_f32(u) for k, u in v.items()}
    return v


def fp16_storage(lead):
    """The mixed-precision entry of a public operator (the reference's `num_fp16_res` / `conv_clamp` path:
    networks.py:994,1031-1035, bias_act.py:88-122, upfirdn2d.py:130-174, conv2d_resample.py:59-154).

    The native kernels of this build are fp32 (include/gagan_b200.h: every entry point is `_f32`).  When the operator's leading
    tensor `lead` is float16, the call is evaluated on fp32 copies of its float16 arguments and the result is rounded to float16
    once, at the operator's boundary -- the same places where the reference's fp16 tensors live, with fp32 (3xTF32) arithmetic in
    between, so a network built with `num_fp16_res > 0` runs unchanged and no worse than on the reference's fp16 cuDNN path.  The
    casts are differentiable torch ops: gradients arrive and leave as float16, as they do in the reference.  This is fp16
    *storage*, not an fp16 tensor-core path: it costs two extra cast passes per operator instead of halving the traffic
    (DESIGN.md section 7)."""
    def deco(fn):
        @functools.wraps(fn)
        def run(*args, **kwargs):
            x = args[0] if args else kwargs.get(lead)
            if not (isinstance(x, torch.Tensor) and x.dtype == torch.float16):
                return fn(*args, **kwargs)
            y = fn(*[to_f32(a) for a in args], **{k: to_f32(v) for k, v in kwargs.items()})
            return y.to(torch.float16) if isinstance(y, torch.Tensor) else y
        return run
    return deco
