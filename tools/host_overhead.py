#!/usr/bin/env python
"""Host-side cost per call of the operator layer on tiny inputs (the launch-bound end of BASELINE.json configs[4]).

    python tools/host_overhead.py [--profile]

Prints wall-clock microseconds per call (GPU idle-bound: the tensors are 4x4, so this is python + ctypes + autograd time) for this
build and for the reference's own modules on the same device, and with --profile the cProfile top list of this build's calls.
"""
import os
import sys
import time
import argparse
import cProfile
import pstats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)

import numpy as np  # noqa: E402
import torch        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--profile', action='store_true')
    ap.add_argument('--calls', type=int, default=2000)
    args = ap.parse_args()
    from torch_utils.ops import bias_act, upfirdn2d
    from gagan_b200.training.networks import modulated_conv2d
    dev = torch.device('cuda:0')
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import sweep_cfg5                        # the reference's modules on the same GPU with its SIMT plugins built (comparison column)
    L, how = sweep_cfg5.reference_ops(dev)
    print('reference elementwise ops:', how)
    torch.backends.cudnn.benchmark = True
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
    n, C, r = 2, 512, 4
    x = torch.randn(n, C, r, r, device=dev)
    b = torch.randn(C, device=dev)
    s = torch.rand(n, C, device=dev) + 0.5
    w3 = torch.randn(C, C, 3, 3, device=dev) / 68
    w1 = torch.randn(3, C, 1, 1, device=dev) / 23

    def wall(fn):
        for _ in range(20):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.calls):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / args.calls * 1e6

    cases = []
    cases.append(('bias_act fwd (no grad)', lambda: bias_act.bias_act(x, b, act='lrelu'), (lambda: L.bias_act.bias_act(x, b, act='lrelu')) if L else None, True))
    cases.append(('upfirdn2d up2 (no grad)', lambda: upfirdn2d.upsample2d(x, f), (lambda: L.upfirdn2d.upsample2d(x, f)) if L else None, True))
    kw3 = dict(padding=1, resample_filter=f, flip_weight=True, demodulate=True)
    kw1 = dict(padding=0, resample_filter=f, flip_weight=True, demodulate=False)
    cases.append(('modconv 3x3 fwd (no grad)', lambda: modulated_conv2d(x=x, weight=w3, styles=s, **kw3),
                  (lambda: L.networks.modulated_conv2d(x=x, weight=w3, styles=s, fused_modconv=False, **kw3)) if L else None, True))
    cases.append(('modconv 1x1 fwd (no grad)', lambda: modulated_conv2d(x=x, weight=w1, styles=s, **kw1),
                  (lambda: L.networks.modulated_conv2d(x=x, weight=w1, styles=s, fused_modconv=False, **kw1)) if L else None, True))
    xg = x.clone().requires_grad_(True); wg = w3.clone().requires_grad_(True)

    def fb(mc, extra):
        def run():
            y = mc(x=xg, weight=wg, styles=s, **kw3, **extra)
            torch.autograd.grad(y.sum(), [xg, wg])
        return run
    cases.append(('modconv 3x3 fwd+bwd', fb(modulated_conv2d, {}), fb(L.networks.modulated_conv2d, dict(fused_modconv=False)) if L else None, False))

    def fb_ba(mod):
        def run():
            y = mod.bias_act(xg, b, act='lrelu')
            torch.autograd.grad(y.sum(), [xg])
        return run
    cases.append(('bias_act fwd+bwd', fb_ba(bias_act), fb_ba(L.bias_act) if L else None, False))

    for name, ours, ref, nograd in cases:
        ctx = torch.no_grad() if nograd else torch.enable_grad()
        with ctx:
            t = wall(ours)
            tr = wall(ref) if ref is not None else float('nan')
        print(f'{name:28s} this build {t:8.1f} us/call   reference {tr:8.1f} us/call', flush=True)
        if args.profile:
            pr = cProfile.Profile()
            with ctx:
                pr.enable()
                for _ in range(500):
                    ours()
                pr.disable()
            torch.cuda.synchronize()
            st = pstats.Stats(pr, stream=sys.stdout)
            st.sort_stats('tottime').print_stats(14)


if __name__ == '__main__':
    main()
