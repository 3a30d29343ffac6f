#!/usr/bin/env python
"""Run one conv configuration a few times (target of `ncu -k regex:conv_tc_kernel`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import numpy as np
import torch
from torch_utils import custom_ops
N, I, O, R, k = [int(v) for v in (sys.argv[1:6] if len(sys.argv) > 5 else (8, 512, 512, 64, 3))]
mode = sys.argv[6] if len(sys.argv) > 6 else 'fwd'
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 3
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
x = torch.randn(N, I, R, R, device=dev); w = torch.randn(O, I, k, k, device=dev) / np.sqrt(I * k * k)
dy = torch.randn(N, O, R, R, device=dev)
for _ in range(reps):
    if mode == 'fwd':
        y = plugin.conv2d(x, w, padding=(k // 2, k // 2))
    else:
        y = plugin.conv2d_wgrad(x, dy, (k, k), padding=(k // 2, k // 2))
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    y = plugin.conv2d(x, w, padding=(k // 2, k // 2)) if mode == 'fwd' else plugin.conv2d_wgrad(x, dy, (k, k), padding=(k // 2, k // 2))
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f'{mode} N={N} {I}->{O} k{k} @{R}: {ms:.3f} ms  {2.0*N*O*I*k*k*R*R/ms/1e9:.1f} TFLOP/s')
