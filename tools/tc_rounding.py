#!/usr/bin/env python
"""Which way does the tcgen05 fp32 accumulator round?  Signed error of the 3xTF32 conv against an fp64 reference for
all-positive, all-negative and random-sign outputs.  RZ (toward zero) gives errors of opposite sign for positive and
negative outputs; floor (toward -inf) gives negative errors for both."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import torch
from torch_utils import custom_ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
g = torch.Generator().manual_seed(3)
for I in (32, 128, 512):
    N, O, R, k = 2, 64, 32, 3
    for name, fx, fw in (('x>0,w>0', torch.abs, torch.abs), ('x>0,w<0', torch.abs, lambda t: -t.abs()), ('random', lambda t: t, lambda t: t)):
        x = fx(torch.randn(N, I, R, R, generator=g)).to(dev)
        w = fw(torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)).to(dev)
        ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=1)
        scale = ref.abs().mean().item()
        for pname, prec in (('tf32x3', custom_ops.PREC_TF32X3), ('tf32x1', custom_ops.PREC_TF32X1), ('fp32_simt', custom_ops.PREC_FP32_SIMT)):
            y = plugin.conv2d(x, w, padding=(1, 1), prec=prec).double()
            e = y - ref
            print(f'I={I:4d} {name:8s} {pname:9s} mean signed err/mean|y| {e.mean().item() / scale:+.3e}   rms err/mean|y| {e.square().mean().sqrt().item() / scale:.3e}'
                  f'   sum-rel-err {abs(e.sum().item()) / max(abs(ref.sum().item()), 1e-30):.3e}', flush=True)
