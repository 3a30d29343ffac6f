"""`fma(a, b, c) = a * b + c` with broadcast-aware hand-written gradients.

API of the reference's `torch_utils/ops/fma.py:15-58`.  The only caller is the non-fused
`modulated_conv2d` (training/networks.py:648); in this build the demodulation scale lives in the conv
epilogue where possible, and this op remains for API compatibility and for the noise add.
"""
import torch


def fma(a, b, c):  # => a * b + c
    return _FusedMultiplyAdd.apply(a, b, c)


class _FusedMultiplyAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):  # pylint: disable=arguments-differ
        out = torch.addcmul(c, a, b)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):  # pylint: disable=arguments-differ
        a, b = ctx.saved_tensors
        da = _unbroadcast(dout * b, a.shape) if ctx.needs_input_grad[0] else None
        db = _unbroadcast(dout * a, b.shape) if ctx.needs_input_grad[1] else None
        dc = _unbroadcast(dout, ctx.c_shape) if ctx.needs_input_grad[2] else None
        return da, db, dc


def _unbroadcast(x, shape):
    """Sum `x` over the dimensions along which `shape` was broadcast (fma.py:49-58)."""
    extra_dims = x.ndim - len(shape)
    assert extra_dims >= 0
    dim = [i for i in range(x.ndim) if x.shape[i] > 1 and (i < extra_dims or shape[i - extra_dims] == 1)]
    if len(dim):
        x = x.sum(dim=dim, keepdim=True)
    if extra_dims:
        x = x.reshape(-1, *x.shape[extra_dims + 1:])
    assert x.shape == shape
    return x
