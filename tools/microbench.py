#!/usr/bin/env python
"""Kernel microbenchmarks vs roofline (BASELINE.json configs[4]): bias_act / upfirdn2d (HBM-bound, GB/s of
algorithmic bytes) and conv2d fwd/dgrad/wgrad (tensor-bound, algorithmic TFLOP/s) at the config-f layer shapes.

    python tools/microbench.py [--quick] [--out gpurun_out/microbench.json]

Timing: CUDA events on the launching stream, 3 warm-ups, then `reps` launches; between launches a >126 MB buffer
is rewritten so that every launch starts with a cold L2 unless the tensors themselves exceed L2.
"""
import os
import sys
import json
import argparse

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators

import numpy as np  # noqa: E402
import torch        # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--out', default='')
    ap.add_argument('--reps', type=int, default=10)
    ap.add_argument('--only', default='')
    args = ap.parse_args()
    from torch_utils import custom_ops
    from torch_utils.ops import upfirdn2d, bias_act
    dev = torch.device('cuda:0')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    plugin = custom_ops.get_plugin('conv2d_plugin')
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) \
        else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
    hbm, tf32 = peaks['hbm_gbs'], peaks['bf16_tflops'] / 2
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)
    rows = []

    def timeit(fn, reps=args.reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(reps):
            flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    def report(kind, name, ms, nbytes=None, flops=None, extra=''):
        if nbytes is not None:
            gbs = nbytes / ms / 1e6
            rows.append(dict(kind=kind, name=name, ms=ms, gbs=gbs, frac_hbm=gbs / hbm))
            print(f'{kind:10s} {name:44s} {ms:9.3f} ms  {gbs:8.1f} GB/s  {100 * gbs / hbm:5.1f}% of measured HBM {extra}', flush=True)
        else:
            tf = flops / ms / 1e9
            rows.append(dict(kind=kind, name=name, ms=ms, tflops=tf, frac_tf32=tf / tf32))
            print(f'{kind:10s} {name:44s} {ms:9.3f} ms  {tf:8.2f} TFLOP/s  {100 * tf / tf32:5.1f}% of TF32 peak {extra}', flush=True)

    chan = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}
    N = 4 if args.quick else 8
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)

    if args.only in ('', 'bias_act'):
        for r in ([64, 256, 1024] if args.quick else [16, 64, 128, 256, 512, 1024]):
            C = chan[r]
            x = torch.randn(N, C, r, r, device=dev); b = torch.randn(C, device=dev)
            ms = timeit(lambda: bias_act.bias_act(x, b, act='lrelu'))
            report('bias_act', f'fwd lrelu [{N},{C},{r},{r}]', ms, nbytes=8 * x.numel())
            y = bias_act.bias_act(x, b, act='lrelu')
            dy = torch.randn_like(y)
            null = torch.empty(0, device=dev)
            bp = custom_ops.get_plugin('bias_act_plugin')
            db = torch.zeros(C, device=dev)
            ms = timeit(lambda: bp.bias_act(dy, b, null, y, null, 1, 1, 3, 0.2, float(np.sqrt(2)), -1.0, dbias=db))
            report('bias_act', f'grad1+db lrelu [{N},{C},{r},{r}]', ms, nbytes=12 * x.numel())
        x = torch.randn(64, 512, device=dev); b = torch.randn(512, device=dev)
        report('bias_act', 'fwd lrelu [64,512] (mapping FC)', timeit(lambda: bias_act.bias_act(x, b, act='lrelu')), nbytes=8 * x.numel())

    if args.only in ('', 'upfirdn2d'):
        for r in ([128, 512] if args.quick else [32, 128, 256, 512]):
            C = chan[2 * r] if 2 * r in chan else 32
            x = torch.randn(N, C, 2 * r + 1, 2 * r + 1, device=dev)
            ms = timeit(lambda: upfirdn2d.upfirdn2d(x, f, padding=[1, 1, 1, 1], gain=4))
            report('upfirdn2d', f'filter p1 g4 [{N},{C},{2*r+1},{2*r+1}] (after up-conv)', ms, nbytes=4 * (x.numel() + N * C * 4 * r * r))
            x = torch.randn(N, chan[2 * r], 2 * r, 2 * r, device=dev)
            ms = timeit(lambda: upfirdn2d.upfirdn2d(x, f, padding=[2, 2, 2, 2]))
            report('upfirdn2d', f'filter p2 [{N},{chan[2*r]},{2*r},{2*r}] (before D conv1)', ms, nbytes=4 * (x.numel() + N * chan[2 * r] * (2 * r + 1) ** 2))
            ms = timeit(lambda: upfirdn2d.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1]))
            report('upfirdn2d', f'down2 [{N},{chan[2*r]},{2*r},{2*r}] (D skip)', ms, nbytes=4 * (x.numel() + x.numel() // 4))
            x = torch.randn(N, chan[2 * r], r, r, device=dev)
            ms = timeit(lambda: upfirdn2d.upfirdn2d(x, f, up=2, padding=[2, 1, 2, 1], gain=4))
            report('upfirdn2d', f'up2 [{N},{chan[2*r]},{r},{r}] (bwd of D skip)', ms, nbytes=4 * (x.numel() + 4 * x.numel()))

    if args.only in ('', 'upfirdn2d', 'fir_pm'):
        # the unit-rate FIRs of the stride-2 layers as the networks run them now: fused with the phase-major re-layout
        for r in ([128, 512] if args.quick else [32, 128, 256, 512]):
            C = chan[2 * r]
            xs_w = (r + 1 + 3) // 4 * 4
            z = torch.randn(N, 4 * C, r + 1, xs_w, device=dev)
            ms = timeit(lambda: upfirdn2d.fir_from_pm(z, f, [1, 1, 1, 1], False, 4, (2 * r + 1, 2 * r + 1)))
            report('upfirdn2d', f'fir_from_pm [{N},4x{C},{r+1},{xs_w}] -> [{N},{C},{2*r},{2*r}] (G up-conv)', ms,
                   nbytes=4 * (N * C * (2 * r + 1) ** 2 + N * C * 4 * r * r))
            x = torch.randn(N, C, 2 * r, 2 * r, device=dev)
            ys, xs = r + 1, (r + 1 + 3) // 4 * 4
            ms = timeit(lambda: upfirdn2d.fir_to_pm(x, f, [2, 2, 2, 2], False, 1, ys, xs))
            report('upfirdn2d', f'fir_to_pm [{N},{C},{2*r},{2*r}] -> [{N},4x{C},{ys},{xs}] (D conv1)', ms,
                   nbytes=4 * (x.numel() + N * C * (2 * r + 1) ** 2))

    if args.only in ('', 'conv'):
        precs = [('auto', custom_ops.PREC_AUTO), ('simt', custom_ops.PREC_FP32_SIMT)]
        shapes = [(64, 512, 512, 3), (256, 128, 128, 3), (1024, 32, 32, 3)] if args.quick else \
                 [(16, 512, 512, 3), (32, 512, 512, 3), (64, 512, 512, 3), (128, 256, 256, 3), (256, 128, 128, 3), (512, 64, 64, 3),
                  (1024, 32, 32, 3), (1024, 32, 3, 1)]
        for r, I, O, k in shapes:
            n = max(1, min(N, (1 << 28) // (max(I, O) * r * r)))
            x = torch.randn(n, I, r, r, device=dev); w = torch.randn(O, I, k, k, device=dev) / np.sqrt(I * k * k)
            dy = torch.randn(n, O, r, r, device=dev)
            fl = 2.0 * n * O * I * k * k * r * r
            for pname, prec in precs:
                if pname == 'simt' and fl > 3e11 and not args.quick:
                    continue
                ms = timeit(lambda: plugin.conv2d(x, w, padding=(k // 2, k // 2), prec=prec), reps=5)
                report('conv', f'fwd {I}->{O} k{k} @{r} N={n} [{pname}]', ms, flops=fl, extra=f'prec={plugin.last_conv_prec}')
                ms = timeit(lambda: plugin.conv2d_wgrad(x, dy, (k, k), padding=(k // 2, k // 2), prec=prec), reps=5)
                report('conv', f'wgrad {I}->{O} k{k} @{r} N={n} [{pname}]', ms, flops=fl, extra=f'prec={plugin.last_wgrad_prec}')
            # on-box library baseline for context (the reference's contraction is cuDNN through ATen, TF32 off)
            ms = timeit(lambda: torch.nn.functional.conv2d(x, w, padding=k // 2), reps=5)
            report('cudnn', f'fwd {I}->{O} k{k} @{r} N={n} [F.conv2d fp32, TF32 off]', ms, flops=fl)

    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        json.dump(dict(peaks=peaks, rows=rows), open(args.out, 'w'), indent=1)


if __name__ == '__main__':
    main()
