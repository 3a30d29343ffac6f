"""Plugin loader: binds libgagan_b200.so (C ABI, include/gagan_b200.h) and exposes it with the
call signatures of the reference's pybind modules.

Mirrors DissimilarDomains/torch_utils/custom_ops.py:23,47,50-136 (`verbosity`, `_cached_plugins`,
`get_plugin(module_name, sources, **build_kwargs)`), with two deliberate differences:

  * the native code is ONE prebuilt sm_100a shared library (built in-tree by
    `make -C ga-gan_b200/csrc`), not a per-op JIT build -- `sources`/`build_kwargs` are accepted
    and ignored;
  * a missing / unloadable library or a non-sm_100 device is a hard RuntimeError.  The reference
    swallows build failures and silently falls back to its slow `impl='ref'` path
    (bias_act.py:76-83, upfirdn2d.py:33-40); this build has no fallback of any kind.

The objects returned by `get_plugin('bias_act_plugin'|'upfirdn2d_plugin'|'conv2d_plugin')` take and
return torch tensors exactly like `_plugin.bias_act(...)` (bias_act.cpp:32) and
`_plugin.upfirdn2d(...)` (upfirdn2d.cpp:16); outputs are allocated by torch (caching allocator),
kernels run on torch's current stream of the tensor's device.
"""
import os
import ctypes
import subprocess
import threading

import torch

# ----------------------------------------------------------------------------
# Global options (custom_ops.py:23).

verbosity = 'brief'  # Verbosity level: 'none', 'brief', 'full'

# ----------------------------------------------------------------------------

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, 'libgagan_b200.so')
CSRC_DIR = os.path.join(_PKG_DIR, 'csrc')

_cached_plugins = dict()
_lib = None
_lib_lock = threading.Lock()

PREC_FP32_SIMT, PREC_TF32X1, PREC_TF32X3, PREC_AUTO, PREC_AUTO_FAST = 0, 1, 3, -1, -2
conv_precision = PREC_AUTO   # module-level switch, like conv2d_gradfix.enabled
conv_profile = None          # bench.py sets this to a list: (kind, flops, prec, start_event, end_event) per conv launch


def _prof_begin(t):
    if conv_profile is None:
        return None
    ev = torch.cuda.Event(enable_timing=True)
    ev.record(torch.cuda.current_stream(t.device))
    return ev


def _prof_end(t, ev0, kind, flops, prec):
    if ev0 is None:
        return
    ev1 = torch.cuda.Event(enable_timing=True)
    ev1.record(torch.cuda.current_stream(t.device))
    conv_profile.append((kind, flops, prec, ev0, ev1))

_c_float_p = ctypes.c_void_p
_SIGNATURES = {
    'gg_last_error': (ctypes.c_char_p, []),
    'gg_version': (ctypes.c_int, []),
    'gg_device_ok': (ctypes.c_int, []),
    'gg_launch_count': (ctypes.c_int64, []),
    'gg_set_conv_kernel_family': (ctypes.c_int, [ctypes.c_int]),
    'gg_watchdog_report': (ctypes.c_int, [ctypes.c_char_p, ctypes.c_int]),
    'gg_bias_act_f32': (ctypes.c_int, [_c_float_p] * 7 + [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                                          ctypes.c_float, ctypes.c_int64, ctypes.c_int, ctypes.c_int64,
                                                          ctypes.c_void_p]),
    'gg_bias_act_noise_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int64, _c_float_p, ctypes.c_int, ctypes.c_float, ctypes.c_float,
                                             ctypes.c_float, ctypes.c_int64, ctypes.c_int, ctypes.c_int64, ctypes.c_void_p]),
    'gg_upfirdn2d_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 15 + [ctypes.c_float, ctypes.c_int, ctypes.c_int,
                                                           ctypes.c_void_p]),
    'gg_fir4_pm_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 7 + [ctypes.c_float] + [ctypes.c_int] * 8 + [ctypes.c_void_p]),
    'gg_chan_dot_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    'gg_scale_rows_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    'gg_fma_rows_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int64, _c_float_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    'gg_axpby_rows_f32': (ctypes.c_int, [_c_float_p] * 5 + [ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]),
    'gg_conv2d_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 14 + [_c_float_p, _c_float_p, ctypes.c_int,
                                                        ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    'gg_conv2d_act_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 14 + [_c_float_p] * 4 + [ctypes.c_int64, ctypes.c_int, ctypes.c_float,
                                                            ctypes.c_float, ctypes.c_float, ctypes.c_int, ctypes.POINTER(ctypes.c_int),
                                                            ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    'gg_chan_dot_preact_f32': (ctypes.c_int, [_c_float_p] * 4 + [ctypes.c_int64, _c_float_p, ctypes.c_int, ctypes.c_int, ctypes.c_int64, ctypes.c_int,
                                                 ctypes.c_float, ctypes.c_float, ctypes.c_void_p]),
    'gg_conv2d_wgrad_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 14 + [_c_float_p, _c_float_p, ctypes.c_int,
                                                              ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    'gg_conv2d_wgrad_pm_f32': (ctypes.c_int, [_c_float_p] * 3 + [ctypes.c_int] * 14 + [_c_float_p, _c_float_p, ctypes.c_int,
                                                                 ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_uint,
                                                                 ctypes.c_void_p]),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES.keys())


def build_library(verbose=False):
    """Compile libgagan_b200.so for sm_100a with the in-tree Makefile (nvcc cross-compiles without a GPU)."""
    cmd = ['make', '-C', CSRC_DIR, '-j', str(min(8, os.cpu_count() or 1))]
    res = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or res.returncode != 0:
        print(res.stdout)
    if res.returncode != 0:
        raise RuntimeError(f'building libgagan_b200.so failed (exit {res.returncode}); see output above')
    return LIB_PATH


def load_library():
    """dlopen the C-ABI library and declare every prototype.  Works without a GPU (cudart is linked
    statically and initialises lazily); compute entry points need an sm_100 device."""
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.isfile(LIB_PATH):
            if verbosity != 'none':
                print(f'libgagan_b200.so not found, building it in {CSRC_DIR} ...', flush=True)
            build_library(verbose=(verbosity == 'full'))
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in _SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
        return lib


def set_conv_kernel_family(family):
    """1 = row-marching kernel for the <= 64-channel 3x3 layers (default), 0 = the tile kernel for everything; returns the old value."""
    return int(load_library().gg_set_conv_kernel_family(int(family)))


def watchdog_report():
    """One line describing the first mbarrier-wait watchdog expiry of this process ('' if none): readable even after the CUDA context faulted."""
    buf = ctypes.create_string_buffer(256)
    return buf.value.decode() if load_library().gg_watchdog_report(buf, 256) else ''


def launch_count():
    return int(load_library().gg_launch_count())


def _check(code, what):
    if code != 0:
        msg = load_library().gg_last_error()
        raise RuntimeError(f'{what}: {msg.decode() if msg else "error " + str(code)}')


def _ptr(t):
    return None if t is None or t.numel() == 0 else ctypes.c_void_p(t.data_ptr())


def _stream(t):
    return ctypes.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _require_cuda(t, name):
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f'{name} must reside on CUDA device')          # bias_act.cpp:35 / upfirdn2d.cpp:19
    if t.dtype != torch.float32:
        raise RuntimeError(f'{name} must be float32 (this build serves the fp32 path; fp16/fp64 are out of scope)')


_device_checked = set()


def _check_device(t):
    idx = t.device.index
    if idx not in _device_checked:
        with torch.cuda.device(t.device):
            if not load_library().gg_device_ok():
                raise RuntimeError('libgagan_b200.so targets sm_100a (B200) only; this device is not sm_100')
        _device_checked.add(idx)


class _Plugin:
    """Tensor-level entry points with the argument lists of the reference's pybind functions."""

    def __init__(self, lib):
        self._lib = lib

    # bias_act.cpp:32-90
    def bias_act(self, x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp, dbias=None):
        _require_cuda(x, 'x')
        _check_device(x)
        for name, t in (('b', b), ('xref', xref), ('yref', yref), ('dy', dy)):
            if t.numel() != 0 and (t.dtype != x.dtype or t.device != x.device):
                raise RuntimeError(f'{name} must have the same dtype and device as x')
        for name, t in (('xref', xref), ('yref', yref), ('dy', dy)):
            if t.numel() != 0 and (t.shape != x.shape or t.stride() != x.stride()):
                raise RuntimeError(f'{name} must have the same shape and layout as x')
        if x.numel() > 2 ** 31 - 1:
            raise RuntimeError('x is too large')
        if b.dim() != 1:
            raise RuntimeError('b must have rank 1')
        if b.numel() != 0 and not (0 <= dim < x.dim()):
            raise RuntimeError('dim is out of bounds')
        if b.numel() != 0 and b.numel() != x.shape[dim]:
            raise RuntimeError('b has wrong number of elements')
        if grad < 0:
            raise RuntimeError('grad must be non-negative')
        if not (x.is_contiguous() or (x.dim() == 4 and x.is_contiguous(memory_format=torch.channels_last))):
            raise RuntimeError('x must be non-overlapping and dense')
        if not b.is_contiguous():
            raise RuntimeError('b must be contiguous')
        y = torch.empty_like(x)
        has_b = b.numel() != 0 or dbias is not None
        size_b = x.shape[dim] if has_b else 1
        step_b = x.stride(dim) if has_b else 1
        with torch.cuda.device(x.device):
            _check(self._lib.gg_bias_act_f32(_ptr(x), _ptr(b), _ptr(xref), _ptr(yref), _ptr(dy), _ptr(y), _ptr(dbias),
                                            int(grad), int(act), float(alpha), float(gain), float(clamp),
                                            x.numel(), int(size_b), int(step_b), _stream(x)), 'bias_act')
        return y

    # forward bias_act with per-pixel noise (include/gagan_b200.h: gg_bias_act_noise_f32); x dense NCHW, dim == 1
    def bias_act_noise(self, x, b, noise, act, alpha, gain, clamp):
        _require_cuda(x, 'x')
        _check_device(x)
        if x.dim() != 4 or not x.is_contiguous():
            raise RuntimeError('bias_act(noise=...): x must be dense NCHW')
        N, C, H, W = x.shape
        noise = noise.contiguous()
        if noise.numel() == H * W:
            nbs = 0
        elif noise.numel() == N * H * W:
            nbs = H * W
        else:
            raise RuntimeError('bias_act(noise=...): noise must be [H,W] or [N,1,H,W]')
        if b.numel() != 0 and b.numel() != C:
            raise RuntimeError('b has wrong number of elements')
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _check(self._lib.gg_bias_act_noise_f32(_ptr(x), _ptr(b), _ptr(noise), nbs, _ptr(y), int(act), float(alpha), float(gain),
                                                  float(clamp), x.numel(), C, H * W, _stream(x)), 'bias_act')
        return y

    # upfirdn2d.cpp:16-94
    def upfirdn2d(self, x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        _require_cuda(x, 'x')
        _check_device(x)
        if f.device != x.device:
            raise RuntimeError('f must reside on the same device as x')
        if f.dtype != torch.float32:
            raise RuntimeError('f must be float32')
        if x.dim() != 4:
            raise RuntimeError('x must be rank 4')
        if f.dim() != 2:
            raise RuntimeError('f must be rank 2')
        if f.shape[0] < 1 or f.shape[1] < 1:
            raise RuntimeError('f must be at least 1x1')
        if upx < 1 or upy < 1:
            raise RuntimeError('upsampling factor must be at least 1')
        if downx < 1 or downy < 1:
            raise RuntimeError('downsampling factor must be at least 1')
        x = x.contiguous()
        f = f.contiguous()
        N, C, H, W = x.shape
        # C integer division truncates toward zero (upfirdn2d.cpp:32-33)
        out_w = int((W * upx + padx0 + padx1 - f.shape[1] + downx) / downx)
        out_h = int((H * upy + pady0 + pady1 - f.shape[0] + downy) / downy)
        if out_w < 1 or out_h < 1:
            raise RuntimeError('output must be at least 1x1')
        y = torch.empty([N, C, out_h, out_w], dtype=x.dtype, device=x.device)
        with torch.cuda.device(x.device):
            _check(self._lib.gg_upfirdn2d_f32(_ptr(x), _ptr(f), _ptr(y), N, C, H, W, f.shape[0], f.shape[1], int(upx), int(upy),
                                             int(downx), int(downy), int(padx0), int(padx1), int(pady0), int(pady1),
                                             1 if flip else 0, float(gain), out_h, out_w, _stream(x)), 'upfirdn2d')
        return y

    # out[n,c] = sum over pixels of a*b (include/gagan_b200.h: gg_chan_dot_f32)
    def chan_dot(self, a, b):
        _require_cuda(a, 'a')
        _require_cuda(b, 'b')
        _check_device(a)
        if a.shape != b.shape or a.dim() < 2:
            raise RuntimeError('chan_dot: a and b must have the same [N,C,...] shape')
        a = a.contiguous(); b = b.contiguous()
        rows = a.shape[0] * a.shape[1]
        out = torch.empty([a.shape[0], a.shape[1]], dtype=a.dtype, device=a.device)
        with torch.cuda.device(a.device):
            _check(self._lib.gg_chan_dot_f32(_ptr(a), _ptr(b), _ptr(out), rows, a.numel() // max(rows, 1), _stream(a)), 'chan_dot')
        return out

    # y[n,c,:,:] = s[n,c] * x[n,c,:,:] (include/gagan_b200.h: gg_scale_rows_f32)
    def scale_rows(self, x, s):
        _require_cuda(x, 'x')
        _require_cuda(s, 's')
        _check_device(x)
        if x.dim() < 2 or tuple(s.shape) != tuple(x.shape[:2]) or s.device != x.device:
            raise RuntimeError('scale_rows: x must be [N,C,...] and s [N,C] on the same device')
        x = x.contiguous(); s = s.contiguous()
        rows = x.shape[0] * x.shape[1]
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _check(self._lib.gg_scale_rows_f32(_ptr(x), _ptr(s), _ptr(y), rows, x.numel() // max(rows, 1), _stream(x)), 'scale_rows')
        return y

    # y[n,c,:,:] = s[n,c] * x[n,c,:,:] + z[n or 0,:,:] (include/gagan_b200.h: gg_fma_rows_f32); z is [H,W], [1,1,H,W] or [N,1,H,W]
    def fma_rows(self, x, s, z):
        for nm, t in (('x', x), ('s', s), ('z', z)):
            _require_cuda(t, nm)
        _check_device(x)
        if x.dim() != 4 or tuple(s.shape) != tuple(x.shape[:2]) or s.device != x.device or z.device != x.device:
            raise RuntimeError('fma_rows: x must be [N,C,H,W], s [N,C], on one device')
        N, C, H, W = x.shape
        if z.numel() == H * W:
            zbs = 0
        elif z.numel() == N * H * W:
            zbs = H * W
        else:
            raise RuntimeError('fma_rows: z must be [H,W], [1,1,H,W] or [N,1,H,W]')
        x = x.contiguous(); s = s.contiguous(); z = z.contiguous()
        y = torch.empty_like(x)
        with torch.cuda.device(x.device):
            _check(self._lib.gg_fma_rows_f32(_ptr(x), _ptr(s), _ptr(z), zbs, _ptr(y), N * C, C, H * W, _stream(x)), 'fma_rows')
        return y

    # y[n,c,:,:] = s1[n,c] * x1[n,c,:,:] + s2[n,c] * x2[n,c,:,:] (include/gagan_b200.h: gg_axpby_rows_f32)
    def axpby_rows(self, x1, s1, x2, s2):
        for nm, t in (('x1', x1), ('s1', s1), ('x2', x2), ('s2', s2)):
            _require_cuda(t, nm)
        _check_device(x1)
        if x1.shape != x2.shape or x1.dim() < 2 or tuple(s1.shape) != tuple(x1.shape[:2]) or s1.shape != s2.shape:
            raise RuntimeError('axpby_rows: x1, x2 must be [N,C,...] of one shape and s1, s2 [N,C]')
        x1 = x1.contiguous(); x2 = x2.contiguous(); s1 = s1.contiguous(); s2 = s2.contiguous()
        rows = x1.shape[0] * x1.shape[1]
        y = torch.empty_like(x1)
        with torch.cuda.device(x1.device):
            _check(self._lib.gg_axpby_rows_f32(_ptr(x1), _ptr(s1), _ptr(x2), _ptr(s2), _ptr(y), rows, x1.numel() // max(rows, 1), _stream(x1)),
                   'axpby_rows')
        return y

    # d out_scale of a convolution with a fused epilogue (include/gagan_b200.h: gg_chan_dot_preact_f32)
    def chan_dot_preact(self, ds, y, bias, noise, act_idx, alpha, gain):
        _require_cuda(ds, 'ds')
        _require_cuda(y, 'y')
        _check_device(ds)
        if ds.shape != y.shape or ds.dim() != 4:
            raise RuntimeError('chan_dot_preact: ds and y must have the same [N,C,H,W] shape')
        ds = ds.contiguous(); y = y.contiguous()
        N, C, H, W = ds.shape
        nbs = 0
        if noise is not None:
            noise = noise.contiguous()
            nbs = H * W if (noise.numel() == N * H * W and noise.ndim == 4) else 0
        out = torch.empty([N, C], dtype=ds.dtype, device=ds.device)
        with torch.cuda.device(ds.device):
            _check(self._lib.gg_chan_dot_preact_f32(_ptr(ds), _ptr(y), _ptr(bias.contiguous() if bias is not None else None), _ptr(noise), nbs,
                                                   _ptr(out), N, C, H * W, int(act_idx), float(alpha), float(gain), _stream(ds)), 'chan_dot_preact')
        return out

    # 4x4 FIR at unit rate with a phase-major side (include/gagan_b200.h: gg_fir4_pm_f32)
    def fir4_pm(self, x, f, padx0, pady0, flip, gain, in_hw, out_hw, in_pm=None, out_pm=None):
        """x: plain [N,C,H,W] or, with in_pm=(pmH,pmW), phase-major [N,4C,pmH,pmW] whose valid logical extent is in_hw.
        Returns plain [N,C,*out_hw] or, with out_pm=(pmH,pmW), phase-major [N,4C,pmH,pmW] (zero outside out_hw)."""
        _require_cuda(x, 'x')
        _check_device(x)
        if f.shape != (4, 4) or f.dtype != torch.float32 or f.device != x.device:
            raise RuntimeError('fir4_pm: f must be a 4x4 float32 filter on the device of x')
        x = x.contiguous(); f = f.contiguous()
        N = x.shape[0]
        C = x.shape[1] // 4 if in_pm is not None else x.shape[1]
        if in_pm is not None and tuple(x.shape[1:]) != (4 * C, in_pm[0], in_pm[1]):
            raise RuntimeError('fir4_pm: phase-major input has the wrong shape')
        if in_pm is None and tuple(x.shape[2:]) != tuple(in_hw):
            raise RuntimeError('fir4_pm: in_hw does not match x')
        shape = [N, 4 * C, out_pm[0], out_pm[1]] if out_pm is not None else [N, C, out_hw[0], out_hw[1]]
        y = torch.empty(shape, dtype=x.dtype, device=x.device)
        ip, op = (in_pm or (0, 0)), (out_pm or (0, 0))
        with torch.cuda.device(x.device):
            _check(self._lib.gg_fir4_pm_f32(_ptr(x), _ptr(f), _ptr(y), N, C, int(in_hw[0]), int(in_hw[1]), int(padx0), int(pady0),
                                           1 if flip else 0, float(gain), int(out_hw[0]), int(out_hw[1]),
                                           1 if in_pm is not None else 0, int(ip[0]), int(ip[1]),
                                           1 if out_pm is not None else 0, int(op[0]), int(op[1]), _stream(x)), 'fir4_pm')
        return y

    # replaces torch.nn.functional.conv2d / conv_transpose2d (conv2d_gradfix.py:141-146), groups == 1
    def conv2d(self, x, w, stride=1, padding=(0, 0), transposed=False, output_padding=(0, 0), flip_w=False,
               in_scale=None, out_scale=None, prec=None, out_hw=None, flop_scale=1.0, epilogue=None):
        """`epilogue=(bias, noise, act_idx, alpha, gain, clamp)`: the bias_act pass that follows the convolution, fused into the
        kernel's store loop (include/gagan_b200.h: gg_conv2d_act_f32); bias [O] / noise [OH,OW] or [N,1,OH,OW] may be None."""
        _require_cuda(x, 'input')
        _require_cuda(w, 'weight')
        _check_device(x)
        x = x.contiguous()
        w = w.contiguous()
        N, I, H, W = x.shape
        if not transposed:
            O, wi, KH, KW = w.shape
            OH = (H + 2 * padding[0] - KH) // stride + 1
            OW = (W + 2 * padding[1] - KW) // stride + 1
        else:
            wi, O, KH, KW = w.shape
            OH = (H - 1) * stride - 2 * padding[0] + KH + output_padding[0]
            OW = (W - 1) * stride - 2 * padding[1] + KW + output_padding[1]
        if out_hw is not None:          # stride-1 only: crop / extend the output into the zero region (include/gagan_b200.h)
            if stride != 1:
                raise RuntimeError('conv2d: out_hw needs stride 1')
            OH, OW = int(out_hw[0]), int(out_hw[1])
        if wi != I:
            raise RuntimeError(f'conv2d: weight expects {wi} input channels, input has {I}')
        if OH < 1 or OW < 1:
            raise RuntimeError('conv2d: output must be at least 1x1')
        if in_scale is not None:
            in_scale = in_scale.contiguous()
            assert in_scale.shape == (N, I) and in_scale.dtype == torch.float32
        if out_scale is not None:
            out_scale = out_scale.contiguous()
            assert out_scale.shape == (N, O) and out_scale.dtype == torch.float32
        y = torch.empty([N, O, OH, OW], dtype=x.dtype, device=x.device)
        used = ctypes.c_int(0)
        ev0 = _prof_begin(x)
        with torch.cuda.device(x.device):
            if epilogue is None:
                _check(self._lib.gg_conv2d_f32(_ptr(x), _ptr(w), _ptr(y), N, I, H, W, O, KH, KW, OH, OW, int(stride), int(padding[0]),
                                              int(padding[1]), 1 if transposed else 0, 1 if flip_w else 0, _ptr(in_scale),
                                              _ptr(out_scale), int(conv_precision if prec is None else prec), ctypes.byref(used),
                                              _stream(x)), 'conv2d')
            else:
                bias, noise, act_idx, alpha, gain, clamp = epilogue
                nbs = 0
                if bias is not None:
                    bias = bias.contiguous()
                    if tuple(bias.shape) != (O,) or bias.dtype != torch.float32 or bias.device != x.device:
                        raise RuntimeError(f'conv2d: the fused bias must be a float32 [{O}] tensor on {x.device}')
                if noise is not None:
                    noise = noise.contiguous()
                    if noise.dtype != torch.float32 or noise.device != x.device or noise.numel() not in (OH * OW, N * OH * OW):
                        raise RuntimeError('conv2d: the fused noise must be float32 [OH,OW] or [N,1,OH,OW] on the device of x')
                    nbs = OH * OW if (noise.numel() == N * OH * OW and noise.ndim == 4) else 0
                fused = ctypes.c_int(0)
                _check(self._lib.gg_conv2d_act_f32(_ptr(x), _ptr(w), _ptr(y), N, I, H, W, O, KH, KW, OH, OW, int(stride), int(padding[0]),
                                                  int(padding[1]), 1 if transposed else 0, 1 if flip_w else 0, _ptr(in_scale),
                                                  _ptr(out_scale), _ptr(bias), _ptr(noise), nbs, int(act_idx), float(alpha), float(gain),
                                                  float(clamp), int(conv_precision if prec is None else prec), ctypes.byref(used),
                                                  ctypes.byref(fused), _stream(x)), 'conv2d')
                self.last_conv_fused = fused.value
        self.last_conv_prec = used.value
        # algorithmic FLOPs: 2*N*O*I*kh*kw*Hout*Wout, transposed: *Hin*Win
        _prof_end(x, ev0, 'convT' if transposed else 'conv', flop_scale * 2.0 * N * O * I * KH * KW * (H * W if (transposed and stride != 1) else OH * OW), used.value)
        return y

    # replaces aten::cudnn_convolution(_transpose)_backward_weight (conv2d_gradfix.py:178-188)
    def conv2d_wgrad(self, a, b, kernel_size, stride=1, padding=(0, 0), flip_w=False, out_layout=0, a_scale=None,
                     b_scale=None, prec=None, flop_scale=1.0, pm=None):
        """`pm=(pm_dim, pm_dead)`: structural hint for the phase-major stride-2 weights (include/gagan_b200.h:
        gg_conv2d_wgrad_pm_f32) -- entries of dw that are zero by construction in the weight may be left zero."""
        _require_cuda(a, 'input')
        _require_cuda(b, 'grad_output')
        _check_device(a)
        a = a.contiguous()
        b = b.contiguous()
        N, A, HA, WA = a.shape
        N2, B, HB, WB = b.shape
        if N != N2:
            raise RuntimeError('conv2d_wgrad: batch size mismatch')
        KH, KW = kernel_size
        # per-sample channel scales: dense [N,A] / [N,B] fp32 on the operands' device (raw pointers go to the kernel)
        for nm, sc, ch in (('a_scale', a_scale, A), ('b_scale', b_scale, B)):
            if sc is not None and (tuple(sc.shape) != (N, ch) or sc.dtype != torch.float32 or sc.device != a.device):
                raise RuntimeError(f'conv2d_wgrad: {nm} must be a float32 [{N},{ch}] tensor on {a.device}')
        a_scale = a_scale.contiguous() if a_scale is not None else None
        b_scale = b_scale.contiguous() if b_scale is not None else None
        shape = [A, B, KH, KW] if out_layout else [B, A, KH, KW]
        dw = torch.empty(shape, dtype=a.dtype, device=a.device)
        used = ctypes.c_int(0)
        ev0 = _prof_begin(a)
        with torch.cuda.device(a.device):
            _check(self._lib.gg_conv2d_wgrad_pm_f32(_ptr(a), _ptr(b), _ptr(dw), N, A, HA, WA, B, HB, WB, KH, KW, int(stride),
                                                   int(padding[0]), int(padding[1]), 1 if flip_w else 0, int(out_layout),
                                                   _ptr(a_scale), _ptr(b_scale), int(conv_precision if prec is None else prec),
                                                   ctypes.byref(used), int(pm[0]) if pm else 0, int(pm[1]) if pm else 0,
                                                   _stream(a)), 'conv2d_wgrad')
        self.last_wgrad_prec = used.value
        _prof_end(a, ev0, 'wgrad', flop_scale * 2.0 * N * A * B * KH * KW * HB * WB, used.value)
        return dw


_KNOWN_PLUGINS = ('bias_act_plugin', 'upfirdn2d_plugin', 'conv2d_plugin')


def get_plugin(module_name, sources=None, **build_kwargs):
    """custom_ops.py:50-136.  Returns the (single, shared) plugin object for any of the known names."""
    assert verbosity in ['none', 'brief', 'full']
    if module_name in _cached_plugins:
        return _cached_plugins[module_name]
    if module_name not in _KNOWN_PLUGINS:
        raise RuntimeError(f'unknown plugin "{module_name}" (this build provides {_KNOWN_PLUGINS})')
    if verbosity == 'full':
        print(f'Setting up plugin "{module_name}" from {LIB_PATH} ...')
    plugin = _Plugin(load_library())
    _cached_plugins[module_name] = plugin
    return plugin
