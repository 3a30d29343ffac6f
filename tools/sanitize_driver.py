#!/usr/bin/env python
"""Small-shape driver for compute-sanitizer (memcheck / racecheck / synccheck / initcheck) over every kernel family of
libgagan_b200.so: the TMA / mbarrier / TMEM pipelines (conv_tc in all four tile configurations, wgrad_tma in both, the
global-load wgrad fallback), the FIR kernels (fir_stream phase-major in / out, fir_resample2 up / down, fir4_tile,
upfirdn2d_generic), bias_act (+noise, grad, fused db), chan_dot and the thin 1x1 kernels.  Shapes are small because the
sanitizer slows kernels by 10-100x; every result is also checked against the exact FFMA path so that a sanitizer run
is a correctness run too.     tools/sanitize.sh runs it under each tool and keeps the logs.
"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gagan_b200

gagan_b200.install(None)
from torch_utils import custom_ops
from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample

dev = torch.device('cuda:0')
torch.manual_seed(0)
plugin = custom_ops.get_plugin('conv2d_plugin')
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
worst = 0.0


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check(name, a, b, tol=2e-5):
    global worst
    e = rel(a.double(), b.double())
    worst = max(worst, e)
    print(f'{name:60s} max-rel-err {e:.2e}', flush=True)
    assert e <= tol, name


# conv_tc: <32,2> <64,2> <128,2> <256,1>, k = 3 / 2 / 1, scales, ragged edges
for (N, I, O, H, W, k) in [(2, 32, 32, 24, 20, 3), (1, 48, 64, 17, 36, 3), (2, 64, 128, 16, 16, 2), (1, 32, 272, 20, 12, 3), (2, 16, 16, 9, 8, 1)]:
    x = torch.randn(N, I, H, W, device=dev); w = torch.randn(O, I, k, k, device=dev) / np.sqrt(I * k * k)
    a = torch.rand(N, I, device=dev) + 0.5; b = torch.rand(N, O, device=dev) + 0.5
    p = (k // 2, k // 2)
    y = plugin.conv2d(x, w, padding=p, in_scale=a, out_scale=b, prec=custom_ops.PREC_TF32X3)
    y0 = plugin.conv2d(x, w, padding=p, in_scale=a, out_scale=b, prec=custom_ops.PREC_FP32_SIMT)
    check(f'conv_tc N{N} {I}->{O} {H}x{W} k{k} scaled', y, y0)
    dy = torch.randn_like(y)
    dw = plugin.conv2d_wgrad(x, dy, (k, k), padding=p, a_scale=a, b_scale=b, prec=custom_ops.PREC_AUTO)   # > 256 channels: SIMT
    dw0 = plugin.conv2d_wgrad(x, dy, (k, k), padding=p, a_scale=a, b_scale=b, prec=custom_ops.PREC_FP32_SIMT)
    check(f'wgrad_tma N{N} {I}->{O} {H}x{W} k{k} scaled', dw, dw0)

# wgrad global-load fallback (rows that are not 16-byte multiples)
x = torch.randn(2, 32, 13, 21, device=dev); dy = torch.randn(2, 32, 13, 21, device=dev)
check('wgrad_tc (unaligned rows)', plugin.conv2d_wgrad(x, dy, (3, 3), padding=(1, 1), prec=custom_ops.PREC_TF32X3),
      plugin.conv2d_wgrad(x, dy, (3, 3), padding=(1, 1), prec=custom_ops.PREC_FP32_SIMT))

# conv2d_resample: phase-major up / down (fir_stream in / out + conv_tc 2x2 + wgrad pm hint), skip path (fir_resample2), autograd
for (N, I, O, R, up, down, k) in [(2, 32, 32, 16, 2, 1, 3), (2, 32, 64, 32, 1, 2, 3), (2, 32, 32, 32, 1, 2, 1), (1, 64, 32, 8, 2, 1, 3)]:
    x = torch.randn(N, I, R, R, device=dev, requires_grad=True); w = (torch.randn(O, I, k, k, device=dev) / np.sqrt(I * k * k)).requires_grad_(True)
    outs = []
    for prec in (custom_ops.PREC_AUTO, custom_ops.PREC_FP32_SIMT):
        custom_ops.conv_precision = prec
        y = conv2d_resample.conv2d_resample(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=(up == 1))
        if prec == custom_ops.PREC_AUTO:
            dy = torch.randn_like(y)
        outs.append([y.detach()] + list(torch.autograd.grad(y, [x, w], dy)))
    custom_ops.conv_precision = custom_ops.PREC_AUTO
    for name, u, v in zip(('y', 'dx', 'dw'), *outs):
        check(f'conv2d_resample N{N} {I}->{O} r{R} up{up} down{down} k{k}: {name}', u, v)

# upfirdn2d family + bias_act
img = torch.randn(2, 3, 16, 16, device=dev, requires_grad=True)
u = upfirdn2d.upsample2d(img, f); d = upfirdn2d.downsample2d(u, f); fl = upfirdn2d.filter2d(img, f)
torch.autograd.grad((u.sum() + d.sum() + fl.sum()), img)
f12 = upfirdn2d.setup_filter(list(np.hanning(12))).to(dev)
upfirdn2d.upsample2d(img, f12, up=2); upfirdn2d.downsample2d(u, f12, down=2, flip_filter=True)
odd = torch.randn(1, 5, 19, 23, device=dev)
upfirdn2d.upfirdn2d(odd, f, up=2, down=1, padding=[2, 1, 2, 1], gain=4); upfirdn2d.upfirdn2d(odd, f, down=2, padding=[1, 1, 1, 1])
x = torch.randn(2, 32, 16, 16, device=dev, requires_grad=True); bb = torch.randn(32, device=dev, requires_grad=True)
nz = torch.randn(2, 1, 16, 16, device=dev)
y = bias_act.bias_act(x, bb, act='lrelu', noise=nz)
gx, gb = torch.autograd.grad(y.square().sum(), [x, bb], create_graph=True)
torch.autograd.grad(gx.square().sum() + gb.square().sum(), [x, bb])
plugin.chan_dot(x.detach(), y.detach())
torch.cuda.synchronize()
print(f'sanitize_driver: done, worst max-rel-err {worst:.2e}, {custom_ops.launch_count()} library launches', flush=True)
