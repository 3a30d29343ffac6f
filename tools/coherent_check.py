#!/usr/bin/env python
"""How large is the COHERENT part of the tensor-core accumulation error at full layer size?  A cancellation-heavy functional of
a 1024^2 layer -- sum over all pixels and channels of (conv output x per-pixel noise), i.e. the noise-strength gradient of
SynthesisLayer -- computed from the tcgen05 conv, the exact FFMA conv and an fp64 CPU reference (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import numpy as np, torch
from torch_utils import custom_ops
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
torch.manual_seed(0)
for (N, C, R) in [(2, 32, 1024), (2, 64, 512), (2, 512, 64)]:
    x = torch.nn.functional.leaky_relu(torch.randn(N, C, R, R), 0.2)          # post-activation statistics (non-zero mean)
    w = torch.randn(C, C, 3, 3) / np.sqrt(9 * C)
    noise = torch.randn(N, 1, R, R)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
    f_ref = float((ref * noise.double()).sum()); typ = float(ref.abs().mean()) * np.sqrt(ref.numel())
    out = []
    for name, prec in (('tf32x3', custom_ops.PREC_TF32X3), ('ffma', custom_ops.PREC_FP32_SIMT)):
        y = plugin.conv2d(x.to(dev), w.to(dev), padding=(1, 1), prec=prec).double().cpu()
        f = float((y * noise.double()).sum())
        out.append(f'{name}: F err / |F| = {abs(f - f_ref) / abs(f_ref):.2e}, err / typical random-walk size = {abs(f - f_ref) / typ:.2e}, max elem err {float((y - ref).abs().max() / ref.abs().max()):.1e}')
    print(f'N={N} C={C} R={R}: F = {f_ref:.4e} (typical |F| ~ {typ:.2e})  ' + ' | '.join(out), flush=True)
