"""`fma(a, b, c) = a * b + c` with broadcast-aware hand-written gradients.

API of the reference's `torch_utils/ops/fma.py:15-58`.  Its one call site is the non-fused `modulated_conv2d`
(training/networks.py:648): `fma(x [N,O,H,W], dcoefs [N,O,1,1], noise [N,1,H,W] | [H,W])`.  This build's own
`modulated_conv2d` carries both terms inside the convolution kernel and never calls it; callers that keep the
reference's formulation (e.g. SimilarDomains/gan_models/StyleGAN2/nvidia.py:86, which binds `torch_utils.ops.fma` by
module path) get that shape from one pass of `gg_fma_rows_f32` (include/gagan_b200.h).  Any other broadcast pattern is the
reference's own expression, `torch.addcmul` (fma.py:23).  The gradients are differentiable torch expressions in both
cases, as in the reference (fma.py:28-45), so the op stays closed under differentiation.
"""
import torch

from .. import custom_ops
from ..._util import fp16_storage


def _on_device(a):
    return a.is_cuda and a.dtype == torch.float32


def _rows_shape(a, b, c):
    """The call shape of networks.py:648 on fp32 CUDA tensors: a [N,C,H,W], b [N,C,1,1], c one plane per sample or per batch."""
    if not (isinstance(a, torch.Tensor) and _on_device(a) and a.ndim == 4 and a.numel() > 0):
        return False
    if not (isinstance(b, torch.Tensor) and isinstance(c, torch.Tensor) and b.dtype == c.dtype == a.dtype and b.device == c.device == a.device):
        return False
    N, C, H, W = a.shape
    return tuple(b.shape) == (N, C, 1, 1) and tuple(c.shape) in ((N, 1, H, W), (1, 1, H, W), (H, W))


@fp16_storage('a')
def fma(a, b, c):  # => a * b + c
    if _rows_shape(a, b, c):
        return _FusedMultiplyAddRows.apply(a, b, c)
    return _FusedMultiplyAdd.apply(a, b, c)


class _FusedMultiplyAddRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, c):  # pylint: disable=arguments-differ
        out = custom_ops.get_plugin('conv2d_plugin').fma_rows(a, b.reshape(b.shape[0], b.shape[1]), c)
        ctx.save_for_backward(a, b)
        ctx.c_shape = c.shape
        return out

    @staticmethod
    def backward(ctx, dout):  # pylint: disable=arguments-differ
        return _FusedMultiplyAdd.backward(ctx, dout)


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
