#!/usr/bin/env python
"""wgrad kernels side by side: the TMA-fed kernel (default) against the global-load kernel (GG_WG_LDG=1), timing and
agreement, including the phase-major hint.  Development tool."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'ga-gan_b200')):
    sys.path.insert(0, p)
import torch
from torch_utils import custom_ops
plugin = custom_ops.get_plugin('conv2d_plugin')
dev = torch.device('cuda:0')
DOWN = (2, sum(1 << ((py * 2 + px) * 4 + a * 2 + b) for py in range(2) for px in range(2) for a in range(2) for b in range(2) if 2 * a + py >= 3 or 2 * b + px >= 3))
# (N, A, HA, WA, B, HB, WB, k, pad, pm)
shapes = [(32, 32, 1024, 1024, 32, 1024, 1024, 3, 1, None), (32, 64, 512, 512, 64, 512, 512, 3, 1, None),
          (32, 128, 256, 256, 128, 256, 256, 3, 1, None), (32, 512, 64, 64, 512, 64, 64, 3, 1, None),
          (32, 512, 129, 132, 256, 128, 128, 2, 0, DOWN), (32, 128, 513, 516, 64, 512, 512, 2, 0, DOWN),
          (32, 512, 32, 32, 512, 32, 32, 3, 1, None), (32, 512, 16, 16, 512, 16, 16, 3, 1, None), (32, 512, 8, 8, 512, 8, 8, 3, 1, None)]
for N, A, HA, WA, B, HB, WB, k, pad, pm in shapes:
    x = torch.randn(N, A, HA, WA, device=dev); dy = torch.randn(N, B, HB, WB, device=dev)
    fl = 2.0 * N * A * B * k * k * HB * WB * (9 / 16 if pm else 1)
    line = f'wgrad a{N}x{A}x{HA}x{WA} b{B}x{HB}x{WB} k{k}:'
    res = {}
    for name, env, hint in (('ldg', '1', None), ('tma', '0', None), ('tma+pm', '0', pm)):
        if name == 'tma+pm' and pm is None:
            continue
        os.environ['GG_WG_LDG'] = env
        for _ in range(2):
            dw = plugin.conv2d_wgrad(x, dy, (k, k), padding=(pad, pad), pm=hint)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            dw = plugin.conv2d_wgrad(x, dy, (k, k), padding=(pad, pad), pm=hint)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        res[name] = dw
        line += f'  {name}: {ms:.2f} ms ({fl / ms / 1e9:.0f} TF)'
    err = float((res['tma'] - res['ldg']).abs().max() / res['ldg'].abs().max())
    line += f'  |tma-ldg|/max = {err:.1e}'
    if pm:
        live = res['tma+pm'] != 0
        err2 = float(((res['tma+pm'] - res['tma']) * live).abs().max() / res['tma'].abs().max())
        line += f'  pm: live-entry err {err2:.1e}, zero fraction {1 - float(live.float().mean()):.3f}'
    print(line, flush=True)
