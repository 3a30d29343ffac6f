#!/usr/bin/env python
"""tcgen05 conv against the exact FFMA kernel over a sweep of small shapes (development tool)."""
import os, sys, itertools
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import numpy as np
import torch
from torch_utils import custom_ops
from torch_utils.ops import conv2d_resample as cr
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
torch.manual_seed(0)
worst = 0
for (N, I, O, R, k), tr, sc in itertools.product([(4, 32, 32, 64, 3), (4, 32, 32, 16, 3), (2, 48, 40, 33, 3), (4, 32, 16, 32, 3), (4, 64, 64, 32, 3), (4, 32, 32, 64, 2), (2, 128, 32, 20, 3), (4, 32, 3, 64, 1), (4, 16, 32, 64, 3), (3, 32, 32, 8, 3)],
                                          [False, True], [False, True]):
    x = torch.randn(N, I, R, R, device=dev)
    w = torch.randn((I, O, k, k) if tr else (O, I, k, k), device=dev) / np.sqrt(I * k * k)
    a = torch.rand(N, I, device=dev) + 0.5 if sc else None
    b = torch.rand(N, O, device=dev) + 0.5 if sc else None
    kw = dict(stride=1, padding=(k // 2, k // 2), transposed=tr, in_scale=a, out_scale=b)
    try:
        y_tc = plugin.conv2d(x, w, prec=custom_ops.PREC_TF32X3, **kw)
    except RuntimeError as e:
        print('skip', (N, I, O, R, k), tr, str(e)[:60]); continue
    y_ref = plugin.conv2d(x, w, prec=custom_ops.PREC_FP32_SIMT, **kw)
    e = float((y_tc - y_ref).abs().max() / y_ref.abs().max())
    worst = max(worst, e)
    flag = '  <<<<<<' if e > 1e-5 else ''
    print(f'N={N} {I}->{O} @{R} k{k} transposed={tr} scales={sc}: max-rel-err {e:.2e}{flag}')
print('worst', worst)
# phase-major stride-2 forms (structurally dead taps, free output extents)
for (N, I, O, R), sc in itertools.product([(4, 32, 32, 32), (4, 32, 16, 64), (2, 64, 64, 16), (4, 32, 32, 8), (4, 32, 32, 4)], [False, True]):
    w = torch.randn(O, I, 3, 3, device=dev) / np.sqrt(I * 9)
    for kind in ('up', 'down', 'up_T', 'down_T'):
        if kind.startswith('up'):
            w2 = cr.phase_major_weight_up(w); x = torch.randn(N, I, R, R, device=dev)
            pad, hw, ci, co = (1, 1), (R + 1, (R + 1 + 3) // 4 * 4), I, 4 * O
        else:
            w2 = cr.phase_major_weight_down(w); x = torch.randn(N, 4 * I, R // 2 + 1, (R // 2 + 1 + 3) // 4 * 4, device=dev)
            pad, hw, ci, co = (0, 0), (R // 2, R // 2), 4 * I, O
        tr = kind.endswith('_T')
        if tr:      # the data-gradient form: transposed, input = a gradient of the forward output
            x = torch.randn(N, co, hw[0], hw[1], device=dev)
            kw = dict(stride=1, padding=(1 - pad[0], 1 - pad[1]), transposed=True, flip_w=True)
            ohw = ((R, R) if kind == 'up_T' else (R // 2 + 1, (R // 2 + 1 + 3) // 4 * 4))
            cin, cout = co, ci
        else:
            kw = dict(stride=1, padding=pad, transposed=False); ohw = hw; cin, cout = ci, co
        a = torch.rand(N, cin, device=dev) + 0.5 if sc else None
        b = torch.rand(N, cout, device=dev) + 0.5 if sc else None
        try:
            y_tc = plugin.conv2d(x, w2, prec=custom_ops.PREC_TF32X3, out_hw=ohw, in_scale=a, out_scale=b, **kw)
        except RuntimeError as e:
            print('skip', kind, (N, I, O, R), str(e)[:80]); continue
        y_ref = plugin.conv2d(x, w2, prec=custom_ops.PREC_FP32_SIMT, out_hw=ohw, in_scale=a, out_scale=b, **kw)
        e = float((y_tc - y_ref).abs().max() / y_ref.abs().max())
        worst = max(worst, e)
        print(f'pm {kind} N={N} {I}->{O} @{R} scales={sc}: max-rel-err {e:.2e}' + ('  <<<<<<' if e > 1e-5 else ''))
print('worst', worst)
