#!/usr/bin/env python
"""Debug harness for the tcgen05 conv path: compares it with the exact-fp32 FFMA kernel and torch on small shapes
and prints where (which pixels / channels) any mismatch sits."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'ga-gan_b200')):
    sys.path.insert(0, p)
import torch
from torch_utils import custom_ops

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device('cuda:0')
plugin = custom_ops.get_plugin('conv2d_plugin')


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def check(N, I, O, H, W, k, scales=False, transposed=False, pattern=None, verbose=True, pad=None):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(N, I, H, W, generator=g).to(dev)
    wshape = (I, O, k, k) if transposed else (O, I, k, k)
    w = (torch.randn(*wshape, generator=g) / np.sqrt(I * k * k)).to(dev)
    if pattern == 'delta':
        x.zero_(); x[:, 0, H // 2, W // 2] = 1.0
    a = torch.randn(N, I, generator=g).to(dev) if scales else None
    b = torch.randn(N, O, generator=g).to(dev) if scales else None
    pad = (k // 2, k // 2) if pad is None else (pad, pad)
    ref = plugin.conv2d(x, w, padding=pad, transposed=transposed, in_scale=a, out_scale=b, prec=custom_ops.PREC_FP32_SIMT)
    out = {}
    for name, prec in (('x3', custom_ops.PREC_TF32X3), ('x1', custom_ops.PREC_TF32X1)):
        y = plugin.conv2d(x, w, padding=pad, transposed=transposed, in_scale=a, out_scale=b, prec=prec)
        torch.cuda.synchronize()
        out[name] = rel(y, ref)
        if name == 'x3' and out[name] > 1e-4 and verbose:
            e = (y - ref).abs() / ref.abs().max()
            bad = e > 1e-4
            print(f'   bad fraction {bad.float().mean().item():.3f}; by channel(first 16): '
                  f'{[round(v, 2) for v in bad.float().mean(dim=(0, 2, 3))[:16].tolist()]}')
            print(f'   by column (ox, first 24): {[round(v, 2) for v in bad.float().mean(dim=(0, 1, 2))[:24].tolist()]}')
            print(f'   by row    (oy, first 24): {[round(v, 2) for v in bad.float().mean(dim=(0, 1, 3))[:24].tolist()]}')
            print(f'   y[0,0,:2,:8]   = {y[0, 0, :2, :8].tolist()}')
            print(f'   ref[0,0,:2,:8] = {ref[0, 0, :2, :8].tolist()}')
    print(f'N={N} I={I} O={O} {H}x{W} k={k} scales={scales} T={transposed}: rel err 3xTF32 {out["x3"]:.2e}   1xTF32 {out["x1"]:.2e}', flush=True)
    return out


if __name__ == '__main__':
    check(1, 16, 16, 16, 16, 1)
    check(1, 16, 16, 16, 16, 3)
    check(1, 32, 32, 16, 16, 3)
    check(2, 64, 128, 32, 32, 3)
    check(1, 48, 40, 24, 40, 3, scales=True)
    check(2, 513, 256, 16, 16, 3)
    check(2, 256, 512, 32, 32, 3, scales=True)
    check(1, 32, 32, 64, 64, 3, transposed=True)
    check(2, 64, 64, 128, 128, 1)
    check(1, 32, 32, 1024, 1024, 3, verbose=False)
