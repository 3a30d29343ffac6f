"""Fused bias + activation (+ gain, clamp) with first- and second-order gradients, on sm_100a.

Same API as the reference's `torch_utils/ops/bias_act.py` (`activation_funcs` :23-60,
`bias_act(x, b, dim, act, alpha, gain, clamp, impl)` :88-122) and the same closed autograd
structure (`BiasActCuda` :181-214, `BiasActCudaGrad` :217-245), over the C-ABI kernel
`gg_bias_act_f32` (include/gagan_b200.h).  Differences, all deliberate:

  * the bias gradient is produced by the SAME kernel launch as dx (warp-shuffle reduction +
    atomics) instead of a second full-tensor `dx.sum(...)` pass (reference :211-212, :243);
  * fp32 kernels; float16 tensors are served with fp32 arithmetic and a float16 result
    (`_util.fp16_storage`); `impl='cuda'` on a non-CUDA tensor or with a missing library raises --
    there is no fallback; `impl='ref'` is not part of the product (the restatement lives in
    oracle/, which only tests and the CPU-baseline leg of bench.py may import).
"""
import numpy as np
import torch

from .. import custom_ops
from ..._util import fp16_storage

# ----------------------------------------------------------------------------

class _Act:
    """One row of `activation_funcs`: attribute access (`.def_gain`, `.cuda_idx`, ...) as the callers expect."""
    __slots__ = ('func', 'def_alpha', 'def_gain', 'cuda_idx', 'ref', 'has_2nd_grad')

    def __init__(self, func, def_alpha, def_gain, cuda_idx, ref, has_2nd_grad):
        self.func, self.def_alpha, self.def_gain = func, def_alpha, def_gain
        self.cuda_idx, self.ref, self.has_2nd_grad = cuda_idx, ref, has_2nd_grad

    def __getitem__(self, key):
        return getattr(self, key)


activation_funcs = {
    'linear':   _Act(func=lambda x, **_: x,                                          def_alpha=0,   def_gain=1,          cuda_idx=1, ref='',  has_2nd_grad=False),
    'relu':     _Act(func=lambda x, **_: torch.nn.functional.relu(x),                def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=2, ref='y', has_2nd_grad=False),
    'lrelu':    _Act(func=lambda x, alpha, **_: torch.nn.functional.leaky_relu(x, alpha), def_alpha=0.2, def_gain=np.sqrt(2), cuda_idx=3, ref='y', has_2nd_grad=False),
    'tanh':     _Act(func=lambda x, **_: torch.tanh(x),                              def_alpha=0,   def_gain=1,          cuda_idx=4, ref='y', has_2nd_grad=True),
    'sigmoid':  _Act(func=lambda x, **_: torch.sigmoid(x),                           def_alpha=0,   def_gain=1,          cuda_idx=5, ref='y', has_2nd_grad=True),
    'elu':      _Act(func=lambda x, **_: torch.nn.functional.elu(x),                 def_alpha=0,   def_gain=1,          cuda_idx=6, ref='y', has_2nd_grad=True),
    'selu':     _Act(func=lambda x, **_: torch.nn.functional.selu(x),                def_alpha=0,   def_gain=1,          cuda_idx=7, ref='y', has_2nd_grad=True),
    'softplus': _Act(func=lambda x, **_: torch.nn.functional.softplus(x),            def_alpha=0,   def_gain=1,          cuda_idx=8, ref='y', has_2nd_grad=True),
    'swish':    _Act(func=lambda x, **_: torch.sigmoid(x) * x,                       def_alpha=0,   def_gain=np.sqrt(2), cuda_idx=9, ref='x', has_2nd_grad=True),
}

# ----------------------------------------------------------------------------

_plugin = None
_null_tensor = torch.empty([0])


def _init():
    """Bind the native library (bias_act.py:70-83).  Raises instead of falling back."""
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('bias_act_plugin', sources=['bias_act.cu'])
    return True


def _null_like(x):
    return torch.empty([0], dtype=x.dtype, device=x.device)


# ----------------------------------------------------------------------------

def _check_input(x):
    if x.device.type != 'cuda':
        raise RuntimeError('bias_act: the B200 build has no CPU path; x must be a CUDA tensor')
    if x.dtype != torch.float32:                        # (float16 never arrives here: fp16_storage)
        raise RuntimeError('bias_act: this build serves fp32 kernels (float16 tensors through them); other dtypes are out of scope')
    _init()


@fp16_storage('x')
def bias_act(x, b=None, dim=1, act='linear', alpha=None, gain=None, clamp=None, impl='cuda', noise=None):
    r"""Fused bias and activation function: `clamp(act(x + b) * gain)`.

    Args / semantics identical to the reference (bias_act.py:88-122).  Supports first and second
    order gradients, not third.  float16 tensors: fp32 arithmetic, float16 result (`_util.fp16_storage`).
    """
    assert isinstance(x, torch.Tensor)
    assert impl in ['ref', 'cuda']
    if impl != 'cuda':
        raise RuntimeError("bias_act: impl='ref' is not shipped in the B200 build (see oracle/ops_ref.py for the CPU restatement)")
    _check_input(x)
    if noise is not None:
        # extension: per-pixel noise ([H,W] or [N,1,H,W]) added before the activation inside the same kernel -- the
        # `x.add_(noise)` / fma pass of the SynthesisLayer (networks.py:648-653)
        if dim != 1 or x.ndim != 4:
            raise RuntimeError('bias_act(noise=...) needs an NCHW tensor with dim == 1')
        spec = activation_funcs[act]
        if 'x' in spec.ref or spec.has_2nd_grad:
            # these activations differentiate through the saved INPUT (swish: act'(x + b)): it must include the noise, so the
            # add is made explicit and the regular path runs (lrelu / linear / relu, what the networks use, stay fused)
            return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x + noise.to(x.dtype), b)
        return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b, noise)
    return _bias_act_cuda(dim=dim, act=act, alpha=alpha, gain=gain, clamp=clamp).apply(x, b)


# ----------------------------------------------------------------------------

_bias_act_cuda_cache = dict()


def _bias_act_cuda(dim=1, act='linear', alpha=None, gain=None, clamp=None):
    """Autograd Function factory (bias_act.py:165-251), cached per argument tuple."""
    assert clamp is None or clamp >= 0
    spec = activation_funcs[act]
    alpha = float(alpha if alpha is not None else spec.def_alpha)
    gain = float(gain if gain is not None else spec.def_gain)
    clamp = float(clamp if clamp is not None else -1)

    key = (dim, act, alpha, gain, clamp)
    if key in _bias_act_cuda_cache:
        return _bias_act_cuda_cache[key]

    is_identity = (act == 'linear' and gain == 1 and clamp < 0)

    def bcast(v, ndim):
        return v.reshape([-1 if i == dim else 1 for i in range(ndim)])

    # Forward op.
    class BiasActCuda(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, b, noise=None):  # pylint: disable=arguments-differ
            ctx.memory_format = torch.channels_last if x.ndim > 2 and x.stride()[1] == 1 else torch.contiguous_format
            ctx.noise_shape = tuple(noise.shape) if noise is not None else None
            if noise is not None:
                ctx.memory_format = torch.contiguous_format
            x = x.contiguous(memory_format=ctx.memory_format)
            b = b.contiguous() if b is not None else _null_like(x)
            null = _null_like(x)
            y = x
            if noise is not None and x.numel() != 0:
                y = _plugin.bias_act_noise(x, b, noise.to(x.dtype), spec.cuda_idx, alpha, gain, clamp)
            elif (not is_identity or b.numel() != 0) and x.numel() != 0:
                y = _plugin.bias_act(x, b, null, null, null, 0, dim, spec.cuda_idx, alpha, gain, clamp)
            ctx.save_for_backward(
                x if 'x' in spec.ref or spec.has_2nd_grad else null,
                b if 'x' in spec.ref or spec.has_2nd_grad else null,
                y if ('y' in spec.ref or clamp >= 0) else null)   # the clamp mask needs y even for 'linear'
            return y

        @staticmethod
        def backward(ctx, dy):  # pylint: disable=arguments-differ
            dy = dy.contiguous(memory_format=ctx.memory_format)
            x, b, y = ctx.saved_tensors
            dx = None
            db = None
            want_dn = ctx.noise_shape is not None and len(ctx.needs_input_grad) > 2 and ctx.needs_input_grad[2]
            if ctx.needs_input_grad[0] or ctx.needs_input_grad[1] or want_dn:
                want_db = bool(ctx.needs_input_grad[1])
                if is_identity:
                    dx = dy
                    if want_db:
                        db = dx.sum([i for i in range(dx.ndim) if i != dim])
                else:
                    dx, db = BiasActCudaGrad.apply(dy, x, b, y, want_db)
                    if not want_db:
                        db = None
            if ctx.noise_shape is None:
                return dx, db
            dn = None
            if want_dn:                                   # the noise enters like a bias that varies per pixel instead of per channel
                dn = dx.sum(dim=1, keepdim=True)          # [N,1,H,W]
                if int(np.prod(ctx.noise_shape)) != dn.numel():   # one plane shared by the batch: [H,W] or [1,1,H,W]
                    dn = dn.sum(dim=0)
                dn = dn.reshape(ctx.noise_shape)
            return (dx if ctx.needs_input_grad[0] else None), db, dn

    # Backward op: (dx, db) = grad-1 kernel with the bias reduction fused in.
    class BiasActCudaGrad(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x, b, y, want_db):  # pylint: disable=arguments-differ
            ctx.memory_format = torch.channels_last if dy.ndim > 2 and dy.stride()[1] == 1 else torch.contiguous_format
            null = _null_like(dy)
            db = torch.zeros([dy.shape[dim]], dtype=dy.dtype, device=dy.device) if want_db else None
            if dy.numel() != 0:
                dx = _plugin.bias_act(dy, b, x, y, null, 1, dim, spec.cuda_idx, alpha, gain, clamp, dbias=db)
            else:
                dx = torch.empty_like(dy)
            ctx.save_for_backward(dy if spec.has_2nd_grad else null, x, b, y)
            ctx.want_db = want_db
            ctx.shape = tuple(dy.shape)
            ctx.set_materialize_grads(False)
            if db is None:
                db = null
                ctx.mark_non_differentiable(db)
            return dx, db

        @staticmethod
        def backward(ctx, d_dx, d_db):  # pylint: disable=arguments-differ
            dy, x, b, y = ctx.saved_tensors
            d_dy = None
            d_x = None
            d_b = None
            d_y = None
            # (dx, db) is linear in dy:  dx = dy * g,  db = sum(dx)  =>  cotangent of dy = g * (d_dx + bcast(d_db))
            tot = None
            if d_dx is not None:
                tot = d_dx.contiguous(memory_format=ctx.memory_format)
            if ctx.want_db and d_db is not None:
                e = bcast(d_db, len(ctx.shape)).expand(ctx.shape)
                tot = (tot + e) if tot is not None else e.contiguous(memory_format=ctx.memory_format)
            if tot is None:
                return None, None, None, None, None

            if ctx.needs_input_grad[0]:
                d_dy, _ = BiasActCudaGrad.apply(tot, x, b, y, False)

            if spec.has_2nd_grad and (ctx.needs_input_grad[1] or ctx.needs_input_grad[2]):
                null = _null_like(tot)
                want = bool(ctx.needs_input_grad[2])
                d_b_buf = torch.zeros([tot.shape[dim]], dtype=tot.dtype, device=tot.device) if want else None
                d_x = _plugin.bias_act(tot, b, x, y, dy, 2, dim, spec.cuda_idx, alpha, gain, clamp, dbias=d_b_buf)
                if want:
                    d_b = d_b_buf

            return d_dy, d_x, d_b, d_y, None

    BiasActCuda.Grad = BiasActCudaGrad          # the closed backward op, also used by the fused conv + bias_act Function (conv2d_gradfix)
    BiasActCuda.is_identity = is_identity
    _bias_act_cuda_cache[key] = BiasActCuda
    return BiasActCuda

# ----------------------------------------------------------------------------
