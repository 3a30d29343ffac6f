"""`conv2d` / `conv_transpose2d` closed under differentiation, on the library's own sm_100a kernels.

API of the reference's `torch_utils/ops/conv2d_gradfix.py`: module switches `enabled`,
`weight_gradients_disabled`, `no_weight_gradients()` (:22-32), `conv2d(...)` (:37-46),
`conv_transpose2d(...)` (:49-58).  The reference wraps cuDNN (`F.conv2d`, `F.conv_transpose2d`,
`aten::cudnn_convolution_backward_weight`, :141-146,178-188) in two autograd Functions whose backward
passes are built from each other, so that gradients of any order exist (:126-212).  This module keeps
exactly that algebra -- forward conv, data gradient = the transposed op, weight gradient = its own op
whose backward is again {conv, transposed conv} -- but every node runs `gg_conv2d_f32` /
`gg_conv2d_wgrad_f32` (tcgen05 implicit GEMM with 3xTF32 operands where the shape is eligible, exact
fp32 FFMA otherwise; see include/gagan_b200.h).

Unlike the reference (:61-81) the custom op is not limited to torch 1.7-1.9 and is always used for
CUDA tensors; `enabled` is kept for API compatibility.  CPU tensors raise: there is no fallback.
"""
import contextlib
import numpy as np
import torch

from .. import custom_ops
from ..._util import fp16_storage

# pylint: disable=redefined-builtin

enabled = False                     # kept for API compatibility (training_loop.py:209 sets it to True)
weight_gradients_disabled = False   # forcefully skip weight gradients (loss.py:98,143 via no_weight_gradients())


@contextlib.contextmanager
def no_weight_gradients():
    global weight_gradients_disabled
    old = weight_gradients_disabled
    weight_gradients_disabled = True
    yield
    weight_gradients_disabled = old


_plugin = None


def _init():
    global _plugin
    if _plugin is None:
        _plugin = custom_ops.get_plugin('conv2d_plugin', sources=['conv_tc.cu', 'conv_simt.cu'])
    return True


def _check_input(input):
    assert isinstance(input, torch.Tensor)
    if input.device.type != 'cuda':
        raise RuntimeError('conv2d_gradfix: the B200 build has no CPU path; input must be a CUDA tensor')
    if input.dtype != torch.float32:
        raise RuntimeError('conv2d_gradfix: this build serves fp32 only')
    _init()


def _is_s1(weight, stride, padding, dilation, groups, output_padding=0):
    k = (int(weight.shape[2]), int(weight.shape[3]))
    pad = _tuple_of_ints(padding, 2)
    return (_tuple_of_ints(stride, 2) == (1, 1) and _tuple_of_ints(dilation, 2) == (1, 1) and groups == 1
            and _tuple_of_ints(output_padding, 2) == (0, 0) and pad[0] <= k[0] - 1 and pad[1] <= k[1] - 1)


@fp16_storage('input')
def conv2d(input, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    _check_input(input)
    if _is_s1(weight, stride, padding, dilation, groups):
        y = conv2d_s1(input, weight, padding=padding)
        return y if bias is None else y + bias.reshape(1, -1, 1, 1)
    return _conv2d_gradfix(transpose=False, weight_shape=weight.shape, stride=stride, padding=padding, output_padding=0,
                           dilation=dilation, groups=groups).apply(input, weight, bias)


@fp16_storage('input')
def conv_transpose2d(input, weight, bias=None, stride=1, padding=0, output_padding=0, groups=1, dilation=1):
    _check_input(input)
    if _is_s1(weight, stride, padding, dilation, groups, output_padding):
        # stride-1 transposed conv == correlation with the transposed, flipped kernel and padding k-1-p
        pad = _tuple_of_ints(padding, 2)
        y = conv2d_s1(input, weight, padding=(weight.shape[2] - 1 - pad[0], weight.shape[3] - 1 - pad[1]), io=True, flip=True)
        return y if bias is None else y + bias.reshape(1, -1, 1, 1)
    return _conv2d_gradfix(transpose=True, weight_shape=weight.shape, stride=stride, padding=padding,
                           output_padding=output_padding, groups=groups, dilation=dilation).apply(input, weight, bias)


def _tuple_of_ints(xs, ndim):
    xs = tuple(xs) if isinstance(xs, (tuple, list)) else (xs,) * ndim
    assert len(xs) == ndim
    assert all(isinstance(x, int) for x in xs)
    return xs


# ----------------------------------------------------------------------------
# Kernel-level helpers: one group at a time (the C ABI serves groups == 1).


def _run_conv(x, w, transpose, stride, padding, output_padding, groups):
    if groups == 1:
        return _plugin.conv2d(x, w, stride=stride, padding=padding, transposed=transpose, output_padding=output_padding)
    xs = x.chunk(groups, dim=1)
    ws = w.chunk(groups, dim=0)
    return torch.cat([_plugin.conv2d(xg, wg, stride=stride, padding=padding, transposed=transpose,
                                     output_padding=output_padding) for xg, wg in zip(xs, ws)], dim=1)


def _run_wgrad(grad_output, input, weight_shape, transpose, stride, padding, groups):
    kh, kw = weight_shape[2], weight_shape[3]
    # y = conv(x, w):   dw[o,i] = sum dy[o] * x[i]      -> a = x,  b = dy
    # y = convT(x, w):  x' = conv(y', w) is its adjoint  -> a = dy, b = x  (dw comes out as [I,O,kh,kw])
    a, b = (input, grad_output) if not transpose else (grad_output, input)
    if groups == 1:
        return _plugin.conv2d_wgrad(a, b, (kh, kw), stride=stride, padding=padding)
    a_s = a.chunk(groups, dim=1)
    b_s = b.chunk(groups, dim=1)
    return torch.cat([_plugin.conv2d_wgrad(ag, bg, (kh, kw), stride=stride, padding=padding) for ag, bg in zip(a_s, b_s)], dim=0)


# ----------------------------------------------------------------------------
# The stride-1 primitive every tensor-core convolution goes through.
#
#   y[n,o,Y,X] = sum_{i,a,b} x[n,i,Y-py+a,X-px+b] * v[o,i,a,b],   Y < OH, X < OW, x == 0 outside its extent
#
# where v is `w` read as [out,in,kh,kw] (io=False) or [in,out,kh,kw] (io=True), spatially flipped when flip=True.
# The output size is a free parameter (cropping / extending into the zero region), which makes the op closed under
# differentiation without any tensor copies:  d/dx = the same op with (io, flip) toggled, pad -> k-1-pad and
# out_hw = x's extent;  d/dw = `gg_conv2d_wgrad_f32`, whose own gradients are again this op (conv2d_gradfix.py:126-212
# is the template).  conv2d_resample's phase-major up/down paths and the plain stride-1 layers all use it.

_conv2d_s1_cache = dict()

pad_unaligned_rows = False          # module switch.  The tensor-core kernels read their input through TMA, whose row pitch must be a multiple
                                    # of 16 bytes: an input whose width is not a multiple of 4 (cropped / odd-size images; never the networks'
                                    # power-of-two maps) falls to the exact FFMA kernel at 7-15 TFLOP/s (conv_tc.cu::conv2d_tc_eligible).  With
                                    # stride 1 the output extent is a free parameter, so zero columns appended to the input rows change nothing:
                                    # True pads such inputs to the next multiple of 4 (one extra pass over x) and keeps them on tcgen05.  Pinned
                                    # on the CPU (tests/test_autograd_algebra.py); its GPU test was written after the round's GPU budget was
                                    # spent -- default off until that test has run on a B200.


def _tma_rows(x, w_shape):
    """x with its rows zero-padded to a multiple of 4 floats when `pad_unaligned_rows` is on and the tensor-core path would otherwise
    lose the call (>= 16 channels on both sides, square kernel of at most 3, maps of at least 64 columns)."""
    W = int(x.shape[3])
    if not pad_unaligned_rows or W % 4 == 0 or W < 64:
        return x
    if min(int(w_shape[0]), int(w_shape[1])) < 16 or int(w_shape[2]) != int(w_shape[3]) or int(w_shape[2]) > 3:
        return x
    return torch.nn.functional.pad(x, (0, 4 - W % 4))


FUSABLE_ACTS = ('linear', 'relu', 'lrelu')          # the activations the convolution kernels apply in their store loop
fuse_epilogue = False                               # module switch.  Measured on the 1024^2 step: neutral (47.2 vs 47.3 img/s) -- the activation
                                                    # arithmetic lands in the consumer warps' store loop, the critical path of the small-K layers, and
                                                    # costs what the separate HBM-speed bias_act pass cost; it saves one activation tensor per layer


def conv2d_s1(x, w, padding=(0, 0), out_hw=None, io=False, flip=False, live=1.0, in_scale=None, out_scale=None, pm=None, epilogue=None):
    """`in_scale [N,I]` / `out_scale [N,O]`: per-sample channel scales folded into the kernel's operand conversion and
    epilogue -- y = out_scale * conv(in_scale * x, w) -- i.e. the style modulation / demodulation of modulated_conv2d
    (networks.py:642,648-651) without their full-tensor multiply passes.

    `epilogue = dict(bias, noise, act, alpha, gain, clamp)`: additionally y = bias_act(y + noise, bias, act, alpha, gain, clamp)
    inside the same kernel launch (see _conv_act_s1)."""
    _check_input(x)
    kh, kw = int(w.shape[2]), int(w.shape[3])
    padding = _tuple_of_ints(padding, 2)
    if out_hw is None:
        out_hw = (x.shape[2] + 2 * padding[0] - kh + 1, x.shape[3] + 2 * padding[1] - kw + 1)
    key = (tuple(w.shape), padding, (int(out_hw[0]), int(out_hw[1])), bool(io), bool(flip), float(live), (tuple(pm) if pm else None))
    if epilogue is not None:
        from . import bias_act as _ba
        act = epilogue.get('act', 'linear')
        spec = _ba.activation_funcs[act]
        alpha = float(epilogue['alpha'] if epilogue.get('alpha') is not None else spec.def_alpha)
        gain = float(epilogue['gain'] if epilogue.get('gain') is not None else spec.def_gain)
        clamp = float(epilogue['clamp'] if epilogue.get('clamp') is not None else -1)
        bias, noise = epilogue.get('bias'), epilogue.get('noise')
        if act in FUSABLE_ACTS and fuse_scales is not False and fuse_epilogue:
            fn = _conv_act_s1(*key, in_scale is not None, out_scale is not None, act, alpha, gain, clamp)
            return fn.apply(x, w, in_scale, out_scale, bias, noise)
        y = conv2d_s1(x, w, padding=padding, out_hw=out_hw, io=io, flip=flip, live=live, in_scale=in_scale, out_scale=out_scale, pm=pm)
        return _ba.bias_act(y, bias, act=act, alpha=alpha, gain=gain, clamp=(clamp if clamp >= 0 else None), noise=noise)
    if in_scale is None and out_scale is None:
        return _conv2d_s1(*key).apply(x, w)
    if fuse_scales is False:            # debug switch: explicit multiply passes around the unscaled kernel
        y = _conv2d_s1(*key).apply(x * in_scale[:, :, None, None] if in_scale is not None else x, w)
        return y * out_scale[:, :, None, None] if out_scale is not None else y
    return _scaled_conv2d_s1(*key, in_scale is not None, out_scale is not None).apply(x, w, in_scale, out_scale)


fuse_scales = True
closed_scaled_backward = True       # gradients of gradients of the scaled convolution stay on the fused kernels (below); False = the
                                    # round-1 formulation that spells the partial derivatives out with broadcast multiplies (A/B switch)
_scaled_conv2d_s1_cache = dict()


class _ChanDot(torch.autograd.Function):
    """out[n,c] = sum_p p[n,c,:] * q[n,c,:] (gg_chan_dot_f32: one pass over both operands, no product tensor).  Its backward is
    two broadcast multiplies written with torch ops, so the node is differentiable to any order."""
    @staticmethod
    def forward(ctx, p, q):
        ctx.save_for_backward(p, q)
        return _plugin.chan_dot(p, q)

    @staticmethod
    def backward(ctx, g):
        p, q = ctx.saved_tensors
        g = g[:, :, None, None]
        return (g * q if ctx.needs_input_grad[0] else None), (g * p if ctx.needs_input_grad[1] else None)


def _safe(s):
    return torch.where(s == 0, torch.ones_like(s), s)


def _scaled_conv2d_s1(weight_shape, padding, out_hw, io, flip, live, pm, has_a, has_b):
    """y = b * conv(a * x, w) with the per-sample scales inside the kernels, differentiable twice ON THE FUSED KERNELS.

    Returned object: `.apply(x, w, a, b)` and `.Wgrad` (the differentiable weight gradient).  Autograd sees two nodes,

        yh = Core(x, w, a, b0)       b0 = b.detach(): one launch, y-hat = b0 * conv(a * x, w)
        y  = OutScale(yh, b)         no launch (y is yh): carries the dependence on b,  d/db = sum_p dy * yh / b0

    so that the demodulation-coefficient gradient d/db is a function of the graph node yh and of nothing else.  In the second-order
    pass (path-length regularisation, loss.py:96-109) the gradient that d/db sends back therefore arrives at Core TOGETHER with
    everything else that flows into yh, and one data-gradient + one weight-gradient launch serve both -- what the reference gets from
    autograd by keeping the unscaled convolution output as a node of its own (networks.py:646-651).  Every first-order partial is a node
    with EXPLICIT derivatives:

        Core.backward (create_graph)   (dx, da) = DxDa(dy, w, a, x; b0),  dw = ScaledWgradS1(dy, x, a, b0)
        DxDa.backward                  the two incoming gradients are combined in one pass (a * gg_dx + g_da * x), then one forward
                                       convolution (-> d/d dy) and one weight gradient; da does not depend on a, so nothing flows there
        OutScale.backward              g_yh = dy (identity; its b-dependence is the node RatioScale), g_b = DemodDot(dy, yh; b0)

    Writing d/da and d/db as quotients (chan_dot(x, dx) / a, chan_dot(dy, y) / b) and letting autograd differentiate those instead
    produces the missing a- / b-dependence as the difference of two large terms: 1.6e-3 on a second-order style gradient (measured).
    (Re-running the forward and calling autograd.grad on it would be wrong as well: out_scale = dcoefs is itself a function of
    in_scale = styles, and the total derivative would count that path twice.)

    Third and higher orders: the second-order nodes hand over to autograd over the spelled-out expression (`partial_grads`).  The whole
    calculus is pinned on the CPU against autograd over y = b * conv(a * x, w) in fp64, up to third order, with the kernels replaced by
    a torch stand-in (tests/test_autograd_algebra.py), and on the GPU by tests/test_gpu_ops.py."""
    key = (weight_shape, padding, out_hw, io, flip, live, pm, has_a, has_b)
    if key in _scaled_conv2d_s1_cache:
        return _scaled_conv2d_s1_cache[key]
    kh, kw = weight_shape[2], weight_shape[3]
    dpad = (kh - 1 - padding[0], kw - 1 - padding[1])
    wg_kw = dict(stride=1, padding=padding, flip_w=flip, out_layout=(1 if io else 0), flop_scale=live, pm=pm)

    def kernel(x, w, a, b, pad, hw, io_, flip_):
        x = _tma_rows(x, w.shape)
        if not io_:
            return _plugin.conv2d(x, w, stride=1, padding=pad, transposed=False, flip_w=flip_, out_hw=hw, flop_scale=live,
                                  in_scale=a, out_scale=b)
        return _plugin.conv2d(x, w, stride=1, padding=(kh - 1 - pad[0], kw - 1 - pad[1]), transposed=True, flip_w=(not flip_),
                              out_hw=hw, flop_scale=live, in_scale=a, out_scale=b)

    def hw_of(t):
        return (int(t.shape[2]), int(t.shape[3]))

    def unscaled_or_scaled(pad, hw, io_, flip_, t, w_, s_in):
        # conv_{pad,hw,io_,flip_}(s_in * t, w_) as a differentiable node (the spelled-out routes below)
        if s_in is None:
            return _conv2d_s1(weight_shape, pad, hw, io_, flip_, live, pm).apply(t, w_)
        return _scaled_conv2d_s1(weight_shape, pad, hw, io_, flip_, live, pm, True, False).apply(t, w_, s_in, None)

    def partial_grads(outputs_fn, grad_outputs, inputs, need):
        # third and higher order: the vector-Jacobian product of the spelled-out expressions, by autograd.  The incoming gradients go in
        # as `grad_outputs` (cotangents), NOT as factors of a scalar: they depend on the inputs themselves through this very node, and
        # differentiating a scalar that contains them would count that path again (3x on a squared-gradient penalty).
        with torch.enable_grad():
            outs = outputs_fn()
            pairs = [(o, go) for o, go in zip(outs, grad_outputs) if go is not None]
            idx = [i for i, t in enumerate(inputs) if need[i] and t is not None and t.requires_grad]
            got = torch.autograd.grad([o for o, _ in pairs], [inputs[i] for i in idx], [go for _, go in pairs],
                                      create_graph=True, allow_unused=True) if (idx and pairs) else ()
        out = [None] * len(inputs)
        for i, v in zip(idx, got):
            out[i] = v
        return out

    class Core(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, a, b0):
            assert tuple(w.shape) == weight_shape
            a_ = a.contiguous() if a is not None else None
            b_ = b0.contiguous() if b0 is not None else None
            ctx.save_for_backward(x, w, a, b0)
            return kernel(x, w, a_, b_, padding, out_hw, io, flip)

        @staticmethod
        def backward(ctx, dy):
            x, w, a, b0 = ctx.saved_tensors
            need = ctx.needs_input_grad
            dx = dw = da = None
            want_da = a is not None and need[2]
            if torch.is_grad_enabled() and closed_scaled_backward:
                if need[0] or want_da:
                    dx, da = DxDa.apply(dy, w, a, x, b0)
                    if not need[0]:
                        dx = None
                    if not want_da:
                        da = None
                if need[1] and not weight_gradients_disabled:
                    dw = ScaledWgradS1.apply(dy, x, a, b0)
                return dx, dw, da, None
            if torch.is_grad_enabled():
                # closed_scaled_backward = False: the round-1 formulation -- the partial derivatives spelled out with broadcast multiplies
                # around the closed, unscaled primitives
                conv = _conv2d_s1(weight_shape, padding, out_hw, io, flip, live, pm)
                conv_t = _conv2d_s1(weight_shape, dpad, hw_of(x), not io, not flip, live, pm)
                xa = x * a[:, :, None, None] if a is not None else x
                dyb = dy * b0[:, :, None, None] if b0 is not None else dy
                if need[0] or want_da:
                    u = conv_t.apply(dyb, w)
                    if need[0]:
                        dx = u * a[:, :, None, None] if a is not None else u
                    if want_da:
                        da = (x * u).sum([2, 3])
                if need[1] and not weight_gradients_disabled:
                    dw = conv.Wgrad.apply(dyb, xa)
                return dx, dw, da, None
            dy = dy.contiguous()
            if need[0] or want_da:
                # d/dx = a * convT(b * dy): the same kernel with the scales swapped
                dx = kernel(dy, w, b0, a, dpad, hw_of(x), not io, not flip)
                if want_da:
                    # d/da[n,i] = sum_p x * convT(b*dy) = sum_p x * dx / a      (a == 0 has measure zero; its gradient reads as 0)
                    da = _plugin.chan_dot(x, dx) / _safe(a)
                if not need[0]:
                    dx = None
            if need[1] and not weight_gradients_disabled:
                dw = _plugin.conv2d_wgrad(x, dy, (kh, kw), a_scale=a, b_scale=b0, **wg_kw)
            return dx, dw, da, None

    class DxDa(torch.autograd.Function):
        """(dx, da) = (a * u, sum_p x * u),  u = convT(b0 * dy, w): the data and style gradients of Core as ONE node."""
        @staticmethod
        def forward(ctx, dy, w, a, x, b0):
            ctx.set_materialize_grads(False)
            dy = dy.contiguous()
            dxf = kernel(dy, w, b0, a, dpad, hw_of(x), not io, not flip)
            da = _plugin.chan_dot(x, dxf) / _safe(a) if a is not None else None
            ctx.save_for_backward(dy, w, a, x, b0, dxf)
            return dxf, da

        @staticmethod
        def backward(ctx, gg, g):
            dy, w, a, x, b0, dxf = ctx.saved_tensors
            need = ctx.needs_input_grad
            if gg is None and g is None:
                return None, None, None, None, None
            if torch.is_grad_enabled():
                def outputs():
                    u = unscaled_or_scaled(dpad, hw_of(x), not io, not flip, dy, w, b0)
                    return [u * a[:, :, None, None] if a is not None else u, (x * u).sum([2, 3])]
                g_dy, g_w, g_a, g_x = partial_grads(outputs, [gg, g], [dy, w, a, x], need)
                return g_dy, g_w, g_a, g_x, None
            g_dy = g_w = g_a = g_x = None
            # e = the gradient that arrives at u, as a conv INPUT:  a * gg (through dx)  +  g * x (through da)
            if gg is not None:
                gg = gg.contiguous()
            if g is not None:
                g = g.contiguous()
            e, e_scale = None, None
            if gg is not None and g is not None:
                e = _plugin.axpby_rows(gg, a, x, g)
            elif gg is not None:
                e, e_scale = gg, a                                   # (a may be None: plain)
            else:
                e, e_scale = x, g
            if need[0]:
                g_dy = kernel(e, w, e_scale, b0, padding, hw_of(dy), io, flip)          # b0 * conv(e, w)
            if need[1] and not weight_gradients_disabled:
                g_w = _plugin.conv2d_wgrad(e, dy, (kh, kw), a_scale=e_scale, b_scale=b0, **wg_kw)
            if a is not None and need[2] and gg is not None:
                g_a = _plugin.chan_dot(gg, dxf) / _safe(a)          # sum_p gg * u      (da itself does not depend on a)
            if need[3] and g is not None:
                g_x = _plugin.scale_rows(dxf, g / _safe(a))         # g * u
            return g_dy, g_w, g_a, g_x, None

    class OutScale(torch.autograd.Function):
        """y = yh * (b / b0): numerically yh itself (no launch); the node that owns the dependence on the output scale b."""
        @staticmethod
        def forward(ctx, yh, b):
            ctx.save_for_backward(yh, b)
            return yh.view_as(yh)

        @staticmethod
        def backward(ctx, dy):
            yh, b = ctx.saved_tensors
            need = ctx.needs_input_grad
            if torch.is_grad_enabled():
                b0 = b.detach()
                return (RatioScale.apply(dy, b, b0) if need[0] else None), (DemodDot.apply(dy, yh, b0) if need[1] else None)
            return (dy if need[0] else None), (_plugin.chan_dot(dy, yh) / _safe(b) if need[1] else None)

    class RatioScale(torch.autograd.Function):
        """g_yh = dy * (b / b0): numerically dy itself (no launch); d/db = sum_p gg * dy / b0."""
        @staticmethod
        def forward(ctx, dy, b, b0):
            ctx.save_for_backward(dy, b0)
            return dy.view_as(dy)

        @staticmethod
        def backward(ctx, gg):
            dy, b0 = ctx.saved_tensors
            need = ctx.needs_input_grad
            if torch.is_grad_enabled():
                return (gg if need[0] else None), ((gg * dy).sum([2, 3]) / _safe(b0) if need[1] else None), None
            return (gg if need[0] else None), (_plugin.chan_dot(gg, dy) / _safe(b0) if need[1] else None), None

    class DemodDot(torch.autograd.Function):
        """db[n,o] = sum_p dy * yh / b0 = sum_p dy * conv(a * x, w): a function of the nodes dy and yh only."""
        @staticmethod
        def forward(ctx, dy, yh, b0):
            ctx.save_for_backward(dy, yh, b0)
            return _plugin.chan_dot(dy, yh) / _safe(b0)

        @staticmethod
        def backward(ctx, g):
            dy, yh, b0 = ctx.saved_tensors
            need = ctx.needs_input_grad
            s = g / _safe(b0)
            if torch.is_grad_enabled():
                s4 = s[:, :, None, None]
                return (s4 * yh if need[0] else None), (s4 * dy if need[1] else None), None
            return (_plugin.scale_rows(yh, s) if need[0] else None), (_plugin.scale_rows(dy, s) if need[1] else None), None

    class ScaledWgradS1(torch.autograd.Function):
        """dw = wgrad(a * x, b * dy) with the scales inside the kernel; gradients w.r.t. (dy, x, a, b) for an incoming ggw are the two
        scaled convolutions with ggw as the weight, and their channel dots (dw is bilinear in (a x, b dy): genuine dependences)."""
        @staticmethod
        def forward(ctx, dy, x, a, b):
            ctx.save_for_backward(dy, x, a, b)
            return _plugin.conv2d_wgrad(x, dy.contiguous(), (kh, kw), a_scale=(a.contiguous() if a is not None else None),
                                        b_scale=(b.contiguous() if b is not None else None), **wg_kw)

        @staticmethod
        def backward(ctx, ggw):
            dy, x, a, b = ctx.saved_tensors
            need = ctx.needs_input_grad
            g_dy = g_x = g_a = g_b = None
            if need[0] or (b is not None and need[3]):
                F = _scaled_conv2d_s1(weight_shape, padding, hw_of(dy), io, flip, live, pm, a is not None, b is not None)
                t = F.apply(x, ggw, a, b)                           # b * conv(a * x, ggw)
                if need[0]:
                    g_dy = t
                if b is not None and need[3]:
                    g_b = _ChanDot.apply(dy, t) / _safe(b)
            if need[1] or (a is not None and need[2]):
                T = _scaled_conv2d_s1(weight_shape, dpad, hw_of(x), not io, not flip, live, pm, b is not None, a is not None)
                t = T.apply(dy, ggw, b, a)                          # a * convT(b * dy, ggw)
                if need[1]:
                    g_x = t
                if a is not None and need[2]:
                    g_a = _ChanDot.apply(x, t) / _safe(a)
            return g_dy, g_x, g_a, g_b

    class ScaledConvS1:
        Wgrad = ScaledWgradS1

        @staticmethod
        def apply(x, w, a, b):
            if b is None:
                return Core.apply(x, w, a, None)
            return OutScale.apply(Core.apply(x, w, a, b.detach()), b)

    _scaled_conv2d_s1_cache[key] = ScaledConvS1
    return ScaledConvS1


_conv_act_s1_cache = dict()


def _conv_act_s1(weight_shape, padding, out_hw, io, flip, live, pm, has_a, has_b, act, alpha, gain, clamp):
    """y = clamp(act(b * conv(a * x, w) + noise + bias) * gain) as ONE kernel launch (gg_conv2d_act_f32): the SynthesisLayer /
    Conv2dLayer pattern `bias_act(modulated_conv2d(...))` (networks.py:904-921, 752-760) when the convolution is the layer's last
    operator.  Backward = the bias_act gradient kernel on the saved OUTPUT (lrelu / relu / linear need nothing else), then the
    ScaledConvS1 gradients; the pre-activation tensor is never stored -- the one place that needs it, the gradient of the output
    scale b, reconstructs it from y (gg_chan_dot_preact_f32).  Closed under double differentiation: with grad mode on inside
    backward every partial derivative is spelled out with differentiable ops over the closed primitives."""
    key = (weight_shape, padding, out_hw, io, flip, live, pm, has_a, has_b, act, alpha, gain, clamp)
    if key in _conv_act_s1_cache:
        return _conv_act_s1_cache[key]
    from . import bias_act as _ba
    _ba._init()                         # the backward below calls the bias_act gradient kernel directly
    spec = _ba.activation_funcs[act]
    BA = _ba._bias_act_cuda(dim=1, act=act, alpha=alpha, gain=gain, clamp=(clamp if clamp >= 0 else None))
    kh, kw = weight_shape[2], weight_shape[3]
    dpad = (kh - 1 - padding[0], kw - 1 - padding[1])
    invertible = clamp < 0 and gain != 0 and (act == 'linear' or (act == 'lrelu' and alpha != 0))

    def kernel(x, w, a, b, pad, hw, io_, flip_, epi=None):
        x = _tma_rows(x, w.shape)
        if not io_:
            return _plugin.conv2d(x, w, stride=1, padding=pad, transposed=False, flip_w=flip_, out_hw=hw, flop_scale=live,
                                  in_scale=a, out_scale=b, epilogue=epi)
        return _plugin.conv2d(x, w, stride=1, padding=(kh - 1 - pad[0], kw - 1 - pad[1]), transposed=True, flip_w=(not flip_),
                              out_hw=hw, flop_scale=live, in_scale=a, out_scale=b, epilogue=epi)

    def unbroadcast_noise(ds, shape):
        dn = ds.sum(dim=1, keepdim=True)                    # [N,1,H,W]
        if int(np.prod(shape)) != dn.numel():               # one plane shared by the batch: [H,W] or [1,1,H,W]
            dn = dn.sum(dim=0)
        return dn.reshape(shape)

    class ConvActS1(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w, a, b, bias, noise):
            assert tuple(w.shape) == weight_shape
            a_ = a.contiguous() if a is not None else None
            b_ = b.contiguous() if b is not None else None
            y = kernel(x, w, a_, b_, padding, out_hw, io, flip,
                       epi=(bias, noise.to(x.dtype) if noise is not None else None, spec.cuda_idx, alpha, gain, clamp))
            ctx.save_for_backward(x, w, a, b, bias, noise, y)
            return y

        @staticmethod
        def backward(ctx, dy):
            x, w, a, b, bias, noise, y = ctx.saved_tensors
            need = ctx.needs_input_grad
            null = torch.empty([0], dtype=dy.dtype, device=dy.device)
            dy = dy.contiguous()
            # 1. through bias / noise / activation: the closed bias_act gradient op (for these activations it only reads y)
            want_db = bool(bias is not None and need[4])
            if BA.is_identity:
                ds = dy
                dbias = ds.sum([0, 2, 3]) if want_db else None
            else:
                ds, dbias = BA.Grad.apply(dy, null, null, y, want_db)
                if not want_db:
                    dbias = None
            dnoise = unbroadcast_noise(ds, noise.shape) if (noise is not None and need[5]) else None
            dx = dw = da = db = None
            # 2. through the scaled convolution (same algebra as ScaledConvS1.backward with dy := ds)
            if torch.is_grad_enabled():
                conv = _conv2d_s1(weight_shape, padding, out_hw, io, flip, live, pm)
                conv_t = _conv2d_s1(weight_shape, dpad, (x.shape[2], x.shape[3]), not io, not flip, live, pm)
                xa = x * a[:, :, None, None] if a is not None else x
                dsb = ds * b[:, :, None, None] if b is not None else ds
                if need[0] or (a is not None and need[2]):
                    u = conv_t.apply(dsb, w)
                    if need[0]:
                        dx = u * a[:, :, None, None] if a is not None else u
                    if a is not None and need[2]:
                        da = (x * u).sum([2, 3])
                if need[1] and not weight_gradients_disabled:
                    dw = conv.Wgrad.apply(dsb, xa)
                if b is not None and need[3]:
                    db = (ds * conv.apply(xa, w)).sum([2, 3])
                return dx, dw, da, db, dbias, dnoise
            ds = ds.contiguous()
            if need[0] or (a is not None and need[2]):
                dx = kernel(ds, w, b, a, dpad, (x.shape[2], x.shape[3]), not io, not flip)
                if a is not None and need[2]:
                    da = _plugin.chan_dot(x, dx) / torch.where(a == 0, torch.ones_like(a), a)
                if not need[0]:
                    dx = None
            if need[1] and not weight_gradients_disabled:
                dw = _plugin.conv2d_wgrad(x, ds, (kh, kw), stride=1, padding=padding, flip_w=flip, out_layout=(1 if io else 0),
                                          flop_scale=live, a_scale=a, b_scale=b, pm=pm)
            if b is not None and need[3]:
                # d/db[n,o] = sum_p ds * conv(a*x, w) = sum_p ds * (pre - bias - noise) / b,   pre = act^-1(y / gain)
                if invertible and (y.shape[2] * y.shape[3]) % 4 == 0:
                    db = _plugin.chan_dot_preact(ds, y, bias, noise, spec.cuda_idx, alpha, gain) / torch.where(b == 0, torch.ones_like(b), b)
                else:                       # relu / clamped outputs cannot be inverted: one more (unscaled) convolution instead
                    db = _plugin.chan_dot(ds, kernel(x, w, a, None, padding, out_hw, io, flip))
            return dx, dw, da, db, dbias, dnoise

    _conv_act_s1_cache[key] = ConvActS1
    return ConvActS1


def _conv2d_s1(weight_shape, padding, out_hw, io, flip, live=1.0, pm=None):
    # `live` = fraction of structurally non-zero weight blocks (9/16 for the phase-major stride-2 forms); it only scales
    # the algorithmic-FLOP accounting of bench.py so that skipped zero blocks are never counted as work.
    # `pm` = (pm_dim, pm_dead): which entries of the weight are zero by construction (gg_conv2d_wgrad_pm_f32): the weight-gradient
    # kernel skips them.
    key = (weight_shape, padding, out_hw, io, flip, live, pm)
    if key in _conv2d_s1_cache:
        return _conv2d_s1_cache[key]
    kh, kw = weight_shape[2], weight_shape[3]
    assert 0 <= padding[0] <= kh - 1 and 0 <= padding[1] <= kw - 1, 'conv2d_s1: padding must be within [0, k-1]'
    assert out_hw[0] >= 1 and out_hw[1] >= 1
    dpad = (kh - 1 - padding[0], kw - 1 - padding[1])

    def run(x, w):
        x = _tma_rows(x, w.shape)
        if not io:
            return _plugin.conv2d(x, w, stride=1, padding=padding, transposed=False, flip_w=flip, out_hw=out_hw, flop_scale=live)
        return _plugin.conv2d(x, w, stride=1, padding=dpad, transposed=True, flip_w=(not flip), out_hw=out_hw, flop_scale=live)

    class ConvS1(torch.autograd.Function):
        @staticmethod
        def forward(ctx, x, w):
            assert tuple(w.shape) == weight_shape
            ctx.save_for_backward(x, w)
            return run(x, w)

        @staticmethod
        def backward(ctx, dy):
            x, w = ctx.saved_tensors
            dx = dw = None
            if ctx.needs_input_grad[0]:
                dx = _conv2d_s1(weight_shape, dpad, (x.shape[2], x.shape[3]), not io, not flip, live, pm).apply(dy, w)
            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                dw = WgradS1.apply(dy, x)
            return dx, dw

    class WgradS1(torch.autograd.Function):
        @staticmethod
        def forward(ctx, dy, x):
            ctx.save_for_backward(dy, x)
            return _plugin.conv2d_wgrad(x, dy, (kh, kw), stride=1, padding=padding, flip_w=flip, out_layout=(1 if io else 0), flop_scale=live, pm=pm)

        @staticmethod
        def backward(ctx, ggw):
            dy, x = ctx.saved_tensors
            g_dy = g_x = None
            if ctx.needs_input_grad[0]:
                g_dy = _conv2d_s1(weight_shape, padding, (dy.shape[2], dy.shape[3]), io, flip, live, pm).apply(x, ggw)
            if ctx.needs_input_grad[1]:
                g_x = _conv2d_s1(weight_shape, dpad, (x.shape[2], x.shape[3]), not io, not flip, live, pm).apply(dy, ggw)
            return g_dy, g_x

    ConvS1.Wgrad = WgradS1
    _conv2d_s1_cache[key] = ConvS1
    return ConvS1


# ----------------------------------------------------------------------------

_conv2d_gradfix_cache = dict()


def _conv2d_gradfix(transpose, weight_shape, stride, padding, output_padding, dilation, groups):
    ndim = 2
    weight_shape = tuple(weight_shape)
    stride = _tuple_of_ints(stride, ndim)
    padding = _tuple_of_ints(padding, ndim)
    output_padding = _tuple_of_ints(output_padding, ndim)
    dilation = _tuple_of_ints(dilation, ndim)

    key = (transpose, weight_shape, stride, padding, output_padding, dilation, groups)
    if key in _conv2d_gradfix_cache:
        return _conv2d_gradfix_cache[key]

    # Validate arguments (conv2d_gradfix.py:110-119).
    assert groups >= 1
    assert len(weight_shape) == ndim + 2
    assert all(stride[i] >= 1 for i in range(ndim))
    assert all(padding[i] >= 0 for i in range(ndim))
    assert all(dilation[i] >= 0 for i in range(ndim))
    if not transpose:
        assert all(output_padding[i] == 0 for i in range(ndim))
    else:
        assert all(0 <= output_padding[i] < max(stride[i], dilation[i]) for i in range(ndim))
    if dilation != (1, 1):
        raise RuntimeError('conv2d_gradfix: dilation != 1 is not served by the B200 build (the networks never use it)')
    if stride[0] != stride[1]:
        raise RuntimeError('conv2d_gradfix: anisotropic stride is not served by the B200 build')

    common_kwargs = dict(stride=stride, padding=padding, dilation=dilation, groups=groups)

    def calc_output_padding(input_shape, output_shape):
        # conv2d_gradfix.py:124-133
        if transpose:
            return [0, 0]
        return [
            input_shape[i + 2]
            - (output_shape[i + 2] - 1) * stride[i]
            - (1 - 2 * padding[i])
            - dilation[i] * (weight_shape[i + 2] - 1)
            for i in range(ndim)
        ]

    class Conv2d(torch.autograd.Function):
        @staticmethod
        def forward(ctx, input, weight, bias):
            assert weight.shape == weight_shape
            output = _run_conv(input, weight, transpose, stride[0], padding, output_padding, groups)
            if bias is not None:
                output = output + bias.reshape(1, -1, 1, 1)
            ctx.save_for_backward(input, weight)
            return output

        @staticmethod
        def backward(ctx, grad_output):
            input, weight = ctx.saved_tensors
            grad_input = None
            grad_weight = None
            grad_bias = None

            if ctx.needs_input_grad[0]:
                p = calc_output_padding(input_shape=input.shape, output_shape=grad_output.shape)
                grad_input = _conv2d_gradfix(transpose=(not transpose), weight_shape=weight_shape, output_padding=p,
                                             **common_kwargs).apply(grad_output, weight, None)
                assert grad_input.shape == input.shape

            if ctx.needs_input_grad[1] and not weight_gradients_disabled:
                grad_weight = Conv2dGradWeight.apply(grad_output, input)
                assert grad_weight.shape == weight_shape

            if ctx.needs_input_grad[2]:
                grad_bias = grad_output.sum([0, 2, 3])

            return grad_input, grad_weight, grad_bias

    class Conv2dGradWeight(torch.autograd.Function):
        @staticmethod
        def forward(ctx, grad_output, input):
            grad_weight = _run_wgrad(grad_output, input, weight_shape, transpose, stride[0], padding, groups)
            assert grad_weight.shape == weight_shape
            ctx.save_for_backward(grad_output, input)
            return grad_weight

        @staticmethod
        def backward(ctx, grad2_grad_weight):
            grad_output, input = ctx.saved_tensors
            grad2_grad_output = None
            grad2_input = None

            if ctx.needs_input_grad[0]:
                grad2_grad_output = Conv2d.apply(input, grad2_grad_weight, None)
                assert grad2_grad_output.shape == grad_output.shape

            if ctx.needs_input_grad[1]:
                p = calc_output_padding(input_shape=input.shape, output_shape=grad_output.shape)
                grad2_input = _conv2d_gradfix(transpose=(not transpose), weight_shape=weight_shape, output_padding=p,
                                              **common_kwargs).apply(grad_output, grad2_grad_weight, None)
                assert grad2_input.shape == input.shape

            return grad2_grad_output, grad2_input

    _conv2d_gradfix_cache[key] = Conv2d
    return Conv2d
