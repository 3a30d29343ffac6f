#!/usr/bin/env python
"""Forward conv timing over the config-f layer shapes incl. the phase-major stride-2 forms (development tool)."""
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
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
N = int(sys.argv[1]) if len(sys.argv) > 1 else 8
def t(run, reps=3):
    for _ in range(2): run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): run()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for I, O, R in [(32, 32, 1024), (64, 64, 512), (128, 128, 256), (256, 256, 128), (512, 512, 64), (512, 512, 32), (512, 512, 16)]:
    x = torch.randn(N, I, R, R, device=dev); w = torch.randn(O, I, 3, 3, device=dev) / np.sqrt(I * 9)
    ms = t(lambda: plugin.conv2d(x, w, padding=(1, 1)))
    print(f'conv3x3 N={N} {I}->{O} @{R}: {ms:.3f} ms {2.0*N*O*I*9*R*R/ms/1e9:.1f} TF', flush=True)
for I, O, R in [(64, 32, 512), (128, 64, 256), (256, 128, 128), (512, 256, 64), (512, 512, 32)]:
    w = torch.randn(O, I, 3, 3, device=dev) / np.sqrt(I * 9)
    x = torch.randn(N, I, R, R, device=dev); w2 = cr.phase_major_weight_up(w)
    ms = t(lambda: plugin.conv2d(x, w2, padding=(1, 1), out_hw=(R + 1, (R + 1 + 3) // 4 * 4)))
    print(f'pm-up   N={N} {I}->{O} @{R}->{2*R}: {ms:.3f} ms {2.0*N*O*I*9*R*R/ms/1e9:.1f} TF', flush=True)
    # D down: in channels O (at 2R) -> out channels I (at R): reuse names: x2 [N,4*O,R+1,..] w [I,O,3,3]
    wd = torch.randn(I, O, 3, 3, device=dev) / np.sqrt(O * 9)
    x2 = torch.randn(N, 4 * O, R + 1, (R + 1 + 3) // 4 * 4, device=dev); w3 = cr.phase_major_weight_down(wd)
    ms = t(lambda: plugin.conv2d(x2, w3, padding=(0, 0), out_hw=(R, R)))
    print(f'pm-down N={N} {O}->{I} @{2*R}->{R}: {ms:.3f} ms {2.0*N*O*I*9*R*R/ms/1e9:.1f} TF', flush=True)
