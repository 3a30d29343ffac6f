#!/usr/bin/env python
"""One phase-major stride-2 conv (the G up-conv form: [N,I,H,W] x W2[4O,I,2,2] -> [N,4O,H+1,W+4]) a few times: ncu target."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import numpy as np
import torch
from torch_utils import custom_ops
from torch_utils.ops import conv2d_resample as cr
N, I, O, R = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (4, 64, 32, 512))]
kind = sys.argv[5] if len(sys.argv) > 5 else 'up'
reps = int(sys.argv[6]) if len(sys.argv) > 6 else 3
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
w = torch.randn(O, I, 3, 3, device=dev) / np.sqrt(I * 9)
if kind == 'up':
    x = torch.randn(N, I, R, R, device=dev)
    w2 = cr.phase_major_weight_up(w)
    run = lambda: plugin.conv2d(x, w2, padding=(1, 1), out_hw=(R + 1, (R + 1 + 3) // 4 * 4))
    fl = 2.0 * N * O * I * 9 * R * R
else:
    x = torch.randn(N, 4 * I, R // 2 + 1, (R // 2 + 1 + 3) // 4 * 4, device=dev)
    w2 = cr.phase_major_weight_down(w)
    run = lambda: plugin.conv2d(x, w2, padding=(0, 0), out_hw=(R // 2, R // 2))
    fl = 2.0 * N * O * I * 9 * (R // 2) ** 2
for _ in range(reps):
    y = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    y = run()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'pm-{kind} N={N} {I}->{O} @{R}: x{tuple(x.shape)} w{tuple(w2.shape)} -> {tuple(y.shape)}  {ms:.3f} ms  {fl/ms/1e9:.1f} TFLOP/s')
