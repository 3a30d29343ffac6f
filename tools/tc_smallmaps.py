#!/usr/bin/env python
"""Sweep the tcgen05 conv over small / odd map sizes against an fp64 reference (debug tool)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators
import torch
from torch_utils import custom_ops
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')
g = torch.Generator().manual_seed(0)
for (N, I, O, H, W, k, pad) in [(2, 32, 32, 4, 4, 3, 1), (2, 33, 32, 4, 4, 3, 1), (4, 128, 32, 5, 8, 2, 0), (3, 32, 32, 8, 8, 3, 1), (2, 32, 128, 4, 4, 2, 1),
                                (2, 32, 32, 4, 8, 3, 1), (2, 32, 32, 8, 4, 3, 1), (2, 32, 32, 16, 4, 3, 1), (2, 32, 32, 4, 16, 3, 1), (1, 16, 16, 4, 4, 1, 0),
                                (2, 32, 32, 12, 12, 3, 1), (2, 32, 32, 20, 20, 3, 1)]:
    x = torch.randn(N, I, H, W, generator=g); w = torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)
    ref = torch.nn.functional.conv2d(x.double(), w.double(), padding=pad)
    for tr in (False, True):
        if tr:
            wt = w.transpose(0, 1).flip([2, 3]).contiguous()     # conv_transpose2d(x, wt, padding=k-1-pad) == conv2d(x, w, padding=pad)
            y = plugin.conv2d(x.to(dev), wt.to(dev), padding=(k - 1 - pad, k - 1 - pad), transposed=True, prec=custom_ops.PREC_AUTO)
        else:
            y = plugin.conv2d(x.to(dev), w.to(dev), padding=(pad, pad), prec=custom_ops.PREC_AUTO)
        e = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
        print(f'N={N} I={I} O={O} {H}x{W} k={k} pad={pad} transposed={tr}: prec={plugin.last_conv_prec} max-rel-err {e:.2e}', flush=True)

print('--- free output extents (crop / extend into the zero region)')
import torch.nn.functional as F
for (N, I, O, H, W, k, pad, OH, OW) in [(2, 32, 128, 4, 4, 2, 1, 5, 8), (2, 128, 32, 5, 8, 2, 0, 4, 4), (2, 128, 32, 5, 8, 2, 1, 4, 4), (2, 32, 32, 16, 16, 2, 1, 17, 20),
                                        (2, 128, 32, 17, 20, 2, 0, 16, 16), (2, 64, 64, 8, 8, 3, 1, 8, 8), (2, 32, 128, 8, 8, 2, 1, 9, 12), (2, 128, 32, 9, 12, 2, 0, 8, 8)]:
    x = torch.randn(N, I, H, W, generator=g); w = torch.randn(O, I, k, k, generator=g) / np.sqrt(I * k * k)
    xp = F.pad(x.double(), (pad, max(OW + k - 1 - W - pad, 0), pad, max(OH + k - 1 - H - pad, 0)))[:, :, :OH + k - 1, :OW + k - 1]
    ref = F.conv2d(xp, w.double())
    for prec in (custom_ops.PREC_AUTO, custom_ops.PREC_FP32_SIMT):
        y = plugin.conv2d(x.to(dev), w.to(dev), padding=(pad, pad), out_hw=(OH, OW), prec=prec)
        e = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
        print(f'N={N} I={I} O={O} {H}x{W} k={k} pad={pad} -> {OH}x{OW}: prec={plugin.last_conv_prec} max-rel-err {e:.2e}', flush=True)
        wt = w.transpose(0, 1).flip([2, 3]).contiguous()
        y = plugin.conv2d(x.to(dev), wt.to(dev), padding=(k - 1 - pad, k - 1 - pad), transposed=True, out_hw=(OH, OW), prec=prec)
        e = float((y.double().cpu() - ref).abs().max() / ref.abs().max())
        print(f'      transposed layout: prec={plugin.last_conv_prec} max-rel-err {e:.2e}', flush=True)
