#!/usr/bin/env python
"""Race hunt: the deterministic kernels (conv_tc forward / data gradient, the FIR kernels) must reproduce their own first
result bit for bit over many launches; the atomically flushed weight gradient to 1e-6.  (development tool)

    python tools/stress_race.py [reps]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import numpy as np, torch
from torch_utils import custom_ops
from torch_utils.ops import upfirdn2d, conv2d_resample as cr
dev = torch.device('cuda:0')
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
torch.manual_seed(0)
bad_total = 0
# (N, I, O, R, up, down)
for case in [(2, 64, 32, 512, 2, 1), (2, 32, 32, 1024, 1, 1), (2, 32, 64, 1024, 1, 2), (4, 512, 512, 64, 1, 1), (4, 512, 512, 32, 2, 1), (2, 128, 256, 256, 1, 2),
             (4, 32, 32, 64, 1, 1), (4, 32, 32, 32, 2, 1), (4, 64, 64, 128, 1, 2)]:
    N, I, O, R, up, down = case
    x = torch.randn(N, I, R, R, device=dev); w = torch.randn(O, I, 3, 3, device=dev) / np.sqrt(9 * I)
    ref = None
    bad = [0, 0, 0]
    for rep in range(reps):
        xr = x.clone().requires_grad_(True); wr = w.clone().requires_grad_(True)
        y = cr.conv2d_resample(xr, wr, f=f, up=up, down=down, padding=1, flip_weight=(up == 1))
        if rep == 0:
            dy = torch.randn_like(y)
        dx, dw = torch.autograd.grad(y, [xr, wr], dy)
        if ref is None:
            ref = (y.detach().clone(), dx.clone(), dw.clone())
            continue
        bad[0] += int(not torch.equal(y.detach(), ref[0]))
        bad[1] += int(not torch.equal(dx, ref[1]))
        bad[2] += int(float((dw - ref[2]).abs().max()) > 1e-5 * float(ref[2].abs().max()))
        if bad[0] + bad[1] + bad[2] and rep % 10 == 0:
            pass
    torch.cuda.synchronize()
    bad_total += sum(bad)
    print(f'{case}: {reps} launches: forward mismatches {bad[0]}, data-gradient mismatches {bad[1]}, weight-gradient outliers {bad[2]}', flush=True)
print('TOTAL', bad_total)
