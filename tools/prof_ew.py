#!/usr/bin/env python
"""Run the HBM-bound kernels once each on the 1024^2 / 512^2 layer shapes (target of `ncu -k regex:...`)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import numpy as np
import torch
from torch_utils import custom_ops
from torch_utils.ops import upfirdn2d, bias_act
dev = torch.device('cuda:0')
N = 4
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
bp = custom_ops.get_plugin('bias_act_plugin')
null = torch.empty(0, device=dev)
x = torch.randn(N, 32, 1024, 1024, device=dev); b = torch.randn(32, device=dev)
z = torch.randn(N, 4 * 32, 513, 516, device=dev)
xs = torch.randn(N, 32, 1024, 1024, device=dev)
xd = torch.randn(N, 64, 512, 512, device=dev)
for rep in range(3):
    y = bias_act.bias_act(x, b, act='lrelu')                                                   # bias_act_vec4<3,0>
    db = torch.zeros(32, device=dev)
    dx = bp.bias_act(x, b, null, y, null, 1, 1, 3, 0.2, float(np.sqrt(2)), -1.0, dbias=db)     # bias_act_vec4<3,1> + fused db
    up = upfirdn2d.fir_from_pm(z, f, [1, 1, 1, 1], False, 4, (1025, 1025))                       # fir_stream<1,1>
    dn = upfirdn2d.fir_to_pm(xs, f, [2, 2, 2, 2], False, 1, 513, 516)                            # fir_stream<2,2>
    d2 = upfirdn2d.upfirdn2d(xd, f, down=2, padding=[1, 1, 1, 1])                                # fir4_tile<1,2>
    u2 = upfirdn2d.upfirdn2d(xd, f, up=2, padding=[2, 1, 2, 1], gain=4)                          # fir4_tile<2,1>
torch.cuda.synchronize()
print('ok')
