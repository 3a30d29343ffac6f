#!/usr/bin/env python
"""conv_march.cu vs the exact FFMA kernel (correctness over pads / flips / scales / ragged shapes) and vs conv_tc.cu (speed).
    python tools/march_check.py [quick]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gagan_b200
gagan_b200.install(None)
from torch_utils import custom_ops
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
torch.manual_seed(0)
bad = 0
cases = [  # N, I, O, H, W, pad, transposed, flip, scaled
    (1, 32, 32, 16, 128, 1, False, False, False), (2, 32, 32, 40, 260, 1, False, False, True), (1, 64, 64, 24, 192, 1, True, False, True),
    (2, 48, 40, 33, 132, 1, False, True, False), (1, 16, 16, 9, 64, 0, False, False, False), (1, 32, 24, 20, 100, 2, False, False, True),
    (1, 96, 64, 12, 128, 1, False, False, False), (2, 32, 32, 130, 512, 1, True, True, True)]
for (N, I, O, H, W, pad, tr, flip, sc) in cases:
    x = torch.randn(N, I, H, W, device=dev)
    w = torch.randn((I, O, 3, 3) if tr else (O, I, 3, 3), device=dev) / np.sqrt(9 * I)
    a = (torch.rand(N, I, device=dev) + 0.5) if sc else None
    b = (torch.rand(N, O, device=dev) + 0.5) if sc else None
    kw = dict(padding=(pad, pad), transposed=tr, flip_w=flip, in_scale=a, out_scale=b)
    y0 = plugin.conv2d(x, w, prec=custom_ops.PREC_FP32_SIMT, **kw)
    custom_ops.set_conv_kernel_family(1)
    y1 = plugin.conv2d(x, w, prec=custom_ops.PREC_TF32X3, **kw)
    custom_ops.set_conv_kernel_family(0)
    y2 = plugin.conv2d(x, w, prec=custom_ops.PREC_TF32X3, **kw)
    custom_ops.set_conv_kernel_family(1)
    torch.cuda.synchronize()
    e1 = float((y1 - y0).abs().max() / y0.abs().max()); e2 = float((y2 - y0).abs().max() / y0.abs().max())
    ok = e1 < 2e-5
    bad += (not ok)
    print(f'{(N, I, O, H, W, pad, tr, flip, sc)}: march {e1:.2e}  tile {e2:.2e}  {"ok" if ok else "MISMATCH"}', flush=True)
    if not ok:
        d = (y1 - y0).abs()
        idx = torch.nonzero(d > 1e-3 * y0.abs().max())
        print('   first bad (n,o,y,x):', idx[:5].tolist(), ' count', idx.shape[0], 'of', d.numel())
print('BAD', bad)
QUICK = len(sys.argv) > 1 and sys.argv[1] == 'quick'
for (N, I, O, R) in ([] if QUICK else [(8, 32, 32, 1024), (8, 64, 64, 512), (32, 32, 32, 1024), (32, 64, 64, 512), (4, 64, 64, 256)]):
    x = torch.randn(N, I, R, R, device=dev); w = torch.randn(O, I, 3, 3, device=dev) / np.sqrt(9 * I)
    a = torch.rand(N, I, device=dev) + 0.5; b = torch.rand(N, O, device=dev) + 0.5
    for fam in (1, 0):
        custom_ops.set_conv_kernel_family(fam)
        for scaled in (False, True):
            kw = dict(padding=(1, 1), in_scale=a if scaled else None, out_scale=b if scaled else None)
            for _ in range(2):
                plugin.conv2d(x, w, **kw)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                plugin.conv2d(x, w, **kw)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f'N{N} {I}->{O} @{R}^2 family {"march" if fam else "tile "} scaled={scaled}: {ms:.3f} ms  {2.0 * N * O * I * 9 * R * R / ms / 1e9:.1f} TFLOP/s', flush=True)
custom_ops.set_conv_kernel_family(1)

# phase-major stride-2 layers (2x2 kernels with structurally dead taps) through conv2d_resample: K = 2 instantiations
from torch_utils.ops import conv2d_resample as cr, upfirdn2d
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
print('--- conv2d_resample up / down (phase-major 2x2): forward + gradients, march vs FFMA')
for (N, I, O, R, up, down) in [(2, 64, 32, 128, 2, 1), (2, 32, 64, 256, 1, 2), (1, 48, 24, 96, 2, 1), (2, 128, 64, 64, 2, 1), (2, 64, 128, 128, 1, 2)]:
    x = torch.randn(N, I, R, R, device=dev, requires_grad=True); w = (torch.randn(O, I, 3, 3, device=dev) / np.sqrt(9 * I)).requires_grad_(True)
    outs = []
    for prec, fam in ((custom_ops.PREC_AUTO, 1), (custom_ops.PREC_AUTO, 0), (custom_ops.PREC_FP32_SIMT, 0)):
        custom_ops.conv_precision = prec; custom_ops.set_conv_kernel_family(fam)
        y = cr.conv2d_resample(x, w, f=f, up=up, down=down, padding=1, flip_weight=(up == 1))
        if not outs:
            dy = torch.randn_like(y)
        outs.append([y.detach()] + list(torch.autograd.grad(y, [x, w], dy)))
    custom_ops.conv_precision = custom_ops.PREC_AUTO; custom_ops.set_conv_kernel_family(1)
    errs = [[float((a - b).abs().max() / b.abs().max()) for a, b in zip(o, outs[2])] for o in outs[:2]]
    ok = max(errs[0]) < 2e-5
    bad += (not ok)
    print(f'{(N, I, O, R, up, down)}: march y/dx/dw ' + ' '.join(f'{e:.1e}' for e in errs[0]) + '   tile ' + ' '.join(f'{e:.1e}' for e in errs[1]) + ('  ok' if ok else '  MISMATCH'), flush=True)
print('BAD', bad)
if len(sys.argv) > 1 and sys.argv[1] == 'quick':
    sys.exit(0)
custom_ops.set_conv_kernel_family(1)
