#!/usr/bin/env python
"""Uninitialised-read hunt (compute-sanitizer initcheck is closed on this pool): poison the allocator caches with NaN, then run every
operator family forward + backward + double backward and compare with the results of the same calls made BEFORE the poisoning
(fresh, zero-filled memory).  Any read of memory the library did not write shows up as NaN or as a changed value.
    python tools/poison_check.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import gagan_b200
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)
from torch_utils import custom_ops
from torch_utils.ops import upfirdn2d, bias_act, conv2d_resample
from gagan_b200.training.networks import modulated_conv2d
dev = torch.device('cuda:0')
f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)


def poison(gb=40):
    """Leave `gb` GiB of NaN-filled blocks of many sizes in torch's caching allocator."""
    blocks = []
    for size in [2 ** k for k in range(10, 31)]:
        n = max(1, min(64, int(gb * 2 ** 30 / 21 / size)))
        for _ in range(n):
            try:
                blocks.append(torch.full((size // 4,), float('nan'), device=dev))
            except RuntimeError:
                break
    torch.cuda.synchronize()
    del blocks


def battery():
    out = {}
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s: torch.randn(*s, device=dev, generator=g)
    for (N, I, O, R, up, down, k) in [(2, 64, 32, 64, 2, 1, 3), (2, 32, 64, 64, 1, 2, 3), (2, 32, 32, 128, 1, 1, 3), (2, 128, 256, 32, 1, 2, 3),
                                       (2, 512, 512, 16, 2, 1, 3), (2, 64, 64, 32, 1, 2, 1), (1, 512, 512, 8, 1, 1, 3), (2, 512, 512, 4, 1, 1, 3),
                                       (2, 48, 24, 20, 1, 1, 3), (2, 3, 32, 64, 1, 1, 1), (2, 32, 3, 64, 1, 1, 1), (3, 513, 512, 4, 1, 1, 3)]:
        x = rn(N, I, R, R).requires_grad_(True); w = (rn(O, I, k, k) / np.sqrt(I * k * k)).requires_grad_(True)
        s = (rn(N, I) * 0.5 + 1).requires_grad_(True)
        for mod in (False, True):
            if mod and down > 1:
                continue
            if mod:
                y = modulated_conv2d(x=x, weight=w, styles=s, up=up, padding=k // 2, resample_filter=f, flip_weight=(up == 1), demodulate=(O != 3))
            else:
                y = conv2d_resample.conv2d_resample(x, w, f=f, up=up, down=down, padding=k // 2, flip_weight=(up == 1))
            dy = rn(*y.shape)
            ins = [x, w] + ([s] if mod else [])
            gr = torch.autograd.grad(y, ins, dy, create_graph=True)
            gg = torch.autograd.grad(sum((t.square().sum() for t in gr)), ins, allow_unused=True)
            key = f'conv N{N} {I}->{O} r{R} up{up} down{down} k{k} mod{int(mod)}'
            out[key] = [y.detach()] + [t.detach() for t in gr] + [t.detach() for t in gg if t is not None]
    x = rn(2, 16, 33, 36).requires_grad_(True); b = rn(16).requires_grad_(True); nz = rn(2, 1, 33, 36)
    y = bias_act.bias_act(x, b, act='lrelu', noise=nz)
    gx, gb_ = torch.autograd.grad(y, [x, b], rn(*y.shape), create_graph=True)
    out['bias_act'] = [y.detach(), gx.detach(), gb_.detach()] + [t.detach() for t in torch.autograd.grad(gx.square().sum(), [x], allow_unused=True) if t is not None]
    img = rn(2, 3, 32, 32).requires_grad_(True)
    u = upfirdn2d.upsample2d(img, f); d = upfirdn2d.downsample2d(u, f); fl = upfirdn2d.filter2d(img, f)
    out['upfirdn2d'] = [u.detach(), d.detach(), fl.detach()] + [t.detach() for t in torch.autograd.grad(u.sum() + d.square().sum() + fl.sum(), [img])]
    odd = rn(1, 5, 19, 23)
    out['upfirdn2d odd'] = [upfirdn2d.upfirdn2d(odd, f, up=2, padding=[2, 1, 2, 1], gain=4), upfirdn2d.upfirdn2d(odd, f, down=2, padding=[1, 1, 1, 1])]
    torch.cuda.synchronize()
    return out


clean = battery()
bad = 0
for rnd in range(3):
    poison()
    dirty = battery()
    for key in clean:
        for i, (a, b) in enumerate(zip(clean[key], dirty[key])):
            nan = bool(torch.isnan(b).any())
            diff = float((a - b).abs().max() / a.abs().max().clamp_min(1e-30)) if not nan else float('nan')
            if nan or diff > 1e-5:
                bad += 1
                print(f'round {rnd}: {key} output {i}: NaN={nan} max-rel-diff vs clean memory {diff:.2e}', flush=True)
print('poison_check: mismatching outputs:', bad)
