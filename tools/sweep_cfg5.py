#!/usr/bin/env python
"""BASELINE configs[4] as SURVEY.md section 8(d) defines it: upfirdn2d / bias_act / modconv microbenchmark sweep over
r in {4, 8, ..., 1024} (C = the config-f channel count at r) and N in {1, 2, 4, 8, 16, 32, 64}, fp32, inputs randn seed 0.

  bias_act    lrelu + bias, gain sqrt(2): forward and grad=1 (with the fused bias gradient)
  upfirdn2d   up2 (pad [2,1,2,1], gain 4), down2 (pad [1,1,1,1]), filter-only (pad [1,1,1,1] on (2r+1)^2 and pad [2,2,2,2] on r^2)
  modconv     3x3 up=1, 3x3 up=2, 1x1 ToRGB; forward / data gradient / weight gradient; styles randn*0.5+1

Next to every point: the REFERENCE on the same GPU -- its own modulated_conv2d / conv2d_resample on cuDNN fp32 (TF32 off) and, for the
elementwise ops, its SIMT plugins (torch_utils/ops/{bias_act,upfirdn2d}.cu JIT-built for this device and pre-seeded into
custom_ops._cached_plugins) or, if the build fails, its impl='ref' torch ops.

Timing: CUDA events on the launching stream, 2 warm-ups, median of `reps`; a 160 MB buffer is rewritten between launches whenever the
operands fit the 126 MB L2.  Rooflines: measured HBM copy bandwidth and measured bf16/2 (MEASURED_PEAKS.json).

    python tools/sweep_cfg5.py [--out gpurun_out/r2_cfg5_sweep] [--reps 3] [--quick]
"""
import os
import sys
import json
import time
import argparse
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np    # noqa: E402
import torch          # noqa: E402
import gagan_b200     # noqa: E402

CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(CHECKOUT if os.path.isdir(CHECKOUT) else None)
from torch_utils import custom_ops                               # noqa: E402
from torch_utils.ops import upfirdn2d, bias_act                  # noqa: E402
from gagan_b200.training.networks import modulated_conv2d        # noqa: E402

warnings.filterwarnings('ignore')
CHAN = {4: 512, 8: 512, 16: 512, 32: 512, 64: 512, 128: 256, 256: 128, 512: 64, 1024: 32}


def reference_ops(dev):
    """The reference's operator modules (private import) with its SIMT plugins built for this GPU if possible."""
    from oracle import live_ref
    if not live_ref.available():
        return None, 'absent'
    L = live_ref.load()
    how = "impl='ref' torch ops"
    try:
        import torch.utils.cpp_extension as ext
        src = os.path.join(live_ref.REF_ROOT, 'torch_utils', 'ops')
        os.environ.setdefault('TORCH_CUDA_ARCH_LIST', '10.0')
        for name in ('bias_act', 'upfirdn2d'):
            mod = ext.load(name=f'ref_{name}_plugin', sources=[os.path.join(src, f'{name}.cpp'), os.path.join(src, f'{name}.cu')],
                           extra_cuda_cflags=['--use_fast_math'], verbose=False)
            getattr(L, name).custom_ops._cached_plugins[f'{name}_plugin'] = mod
        how = 'SIMT plugins (bias_act.cu, upfirdn2d.cu built for sm_100)'
    except Exception as e:      # no ninja / compile error: fall back to the torch ops the reference itself falls back to
        how += f' (plugin build failed: {str(e)[:80]})'
    return L, how


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'r2_cfg5_sweep'))
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--quick', action='store_true')
    ap.add_argument('--no-reference', action='store_true')
    args = ap.parse_args()
    dev = torch.device('cuda:0')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cudnn.benchmark = True                        # training_loop.py:206: the reference's own setting
    custom_ops.verbosity = 'none'
    peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json'))) if os.path.isfile(os.path.join(ROOT, 'MEASURED_PEAKS.json')) \
        else dict(hbm_gbs=6650.0, bf16_tflops=1590.0)
    hbm, tf32 = peaks['hbm_gbs'], peaks['bf16_tflops'] / 2
    L, ref_how = (None, 'skipped') if args.no_reference else reference_ops(dev)
    flush = torch.empty(160 * 1024 * 1024 // 4, device=dev)
    f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
    rows, t_start = [], time.time()

    def timeit(fn, nbytes):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.reps):
            if nbytes < 126e6:
                flush.add_(1.0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        return float(np.median(ts))

    def point(op, r, n, ours, ref, nbytes=None, flops=None):
        row = dict(op=op, r=r, N=n)
        work = nbytes if nbytes is not None else flops
        try:
            row['ms'] = timeit(ours, nbytes or 0)
        except Exception as e:
            row['error'] = str(e)[:100]
        if ref is not None:
            try:
                row['ref_ms'] = timeit(ref, nbytes or 0)
            except Exception as e:
                row['ref_error'] = str(e)[:100]
        if 'ms' in row:
            if nbytes is not None:
                row['gbs'] = nbytes / row['ms'] / 1e6; row['frac'] = row['gbs'] / hbm
            else:
                row['tflops'] = flops / row['ms'] / 1e9; row['frac'] = row['tflops'] / tf32
            if 'ref_ms' in row:
                row['speedup'] = row['ref_ms'] / row['ms']
        rows.append(row)
        unit = f"{row.get('gbs', 0):8.0f} GB/s" if nbytes is not None else f"{row.get('tflops', 0):8.1f} TFLOP/s"
        print(f"{op:24s} r={r:<5d} N={n:<3d} {row.get('ms', float('nan')):9.4f} ms {unit} {100 * row.get('frac', 0):5.1f}%   reference "
              f"{row.get('ref_ms', float('nan')):9.4f} ms  x{row.get('speedup', float('nan')):.2f}", flush=True)

    res = [64, 1024] if args.quick else [4, 8, 16, 32, 64, 128, 256, 512, 1024]
    batches = [1, 8] if args.quick else [1, 2, 4, 8, 16, 32, 64]
    g = torch.Generator(device=dev).manual_seed(0)
    for r in res:
        C = CHAN[r]
        for n in batches:
            if n * C * r * r > 2 ** 31 - 1 or n * C * r * r * 4 > 12e9:
                continue
            x = torch.randn(n, C, r, r, device=dev, generator=g)
            b = torch.randn(C, device=dev, generator=g)
            nb = 4 * x.numel()
            # ---- bias_act
            point('bias_act fwd', r, n, lambda: bias_act.bias_act(x, b, act='lrelu'),
                  (lambda: L.bias_act.bias_act(x, b, act='lrelu')) if L else None, nbytes=2 * nb)
            xr = x.clone().requires_grad_(True)
            y = bias_act.bias_act(xr, b, act='lrelu'); dy = torch.randn_like(y)
            ref = None
            if L:
                xr2 = x.clone().requires_grad_(True); br2 = b.clone().requires_grad_(True)
                y2 = L.bias_act.bias_act(xr2, br2, act='lrelu')
                ref = lambda: torch.autograd.grad(y2, [xr2, br2], dy, retain_graph=True)
            br = b.clone().requires_grad_(True)
            y = bias_act.bias_act(xr, br, act='lrelu')
            point('bias_act grad1+db', r, n, lambda: torch.autograd.grad(y, [xr, br], dy, retain_graph=True), ref, nbytes=3 * nb)
            del y, xr, dy
            # ---- upfirdn2d
            point('upfirdn2d up2', r, n, lambda: upfirdn2d.upfirdn2d(x, f, up=2, padding=[2, 1, 2, 1], gain=4),
                  (lambda: L.upfirdn2d.upfirdn2d(x, f, up=2, padding=[2, 1, 2, 1], gain=4)) if L and n * C * r * r * 16 < 8e9 else None, nbytes=5 * nb)
            if r >= 8:
                point('upfirdn2d down2', r, n, lambda: upfirdn2d.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1]),
                      (lambda: L.upfirdn2d.upfirdn2d(x, f, down=2, padding=[1, 1, 1, 1])) if L else None, nbytes=nb + nb // 4)
            point('upfirdn2d filter p2', r, n, lambda: upfirdn2d.upfirdn2d(x, f, padding=[2, 2, 2, 2]),
                  (lambda: L.upfirdn2d.upfirdn2d(x, f, padding=[2, 2, 2, 2])) if L else None, nbytes=nb + 4 * n * C * (r + 1) ** 2)
            if r <= 512:
                x1 = torch.randn(n, CHAN[2 * r], 2 * r + 1, 2 * r + 1, device=dev, generator=g)
                point('upfirdn2d filter p1', 2 * r + 1, n, lambda: upfirdn2d.upfirdn2d(x1, f, padding=[1, 1, 1, 1], gain=4),
                      (lambda: L.upfirdn2d.upfirdn2d(x1, f, padding=[1, 1, 1, 1], gain=4)) if L else None,
                      nbytes=4 * (x1.numel() + n * CHAN[2 * r] * 4 * r * r))
                del x1
            # ---- modconv: 3x3 same resolution, 3x3 up=2 (from r/2), 1x1 ToRGB
            s = (torch.randn(n, C, device=dev, generator=g) * 0.5 + 1)
            for name, up, k, O in (('modconv 3x3', 1, 3, C), ('modconv 3x3 up2', 2, 3, C), ('modconv 1x1 torgb', 1, 1, 3)):
                if up == 2 and r < 8:
                    continue
                Cin = CHAN[r // 2] if up == 2 else C
                rin = r // up
                xi = torch.randn(n, Cin, rin, rin, device=dev, generator=g).requires_grad_(True)
                w = (torch.randn(O, Cin, k, k, device=dev, generator=g) / np.sqrt(Cin * k * k)).requires_grad_(True)
                si = s if Cin == C else (torch.randn(n, Cin, device=dev, generator=g) * 0.5 + 1)
                kw = dict(up=up, padding=k // 2, resample_filter=f, flip_weight=(up == 1), demodulate=(k == 3))
                fl = 2.0 * n * O * Cin * k * k * (rin * rin if up == 2 else r * r)
                ours_f = lambda: modulated_conv2d(x=xi, weight=w, styles=si, **kw)
                ref_f = (lambda: L.networks.modulated_conv2d(x=xi, weight=w, styles=si, fused_modconv=False, **kw)) if L else None
                with torch.no_grad():
                    point(name + ' fwd', r, n, ours_f, ref_f, flops=fl)
                # one graph per gradient: only the tensor in question requires grad, so both sides compute exactly that gradient
                dyo = None
                for what, tx, tw in (('dgrad', True, False), ('wgrad', False, True)):
                    xi.requires_grad_(tx); w.requires_grad_(tw)
                    yo = ours_f()
                    if dyo is None:
                        dyo = torch.randn_like(yo)
                    yr = ref_f() if L else None
                    tgt = xi if tx else w
                    point(f'{name} {what}', r, n, lambda: torch.autograd.grad(yo, tgt, dyo, retain_graph=True),
                          (lambda: torch.autograd.grad(yr, tgt, dyo, retain_graph=True)) if L else None, flops=fl)
                    del yo, yr
                del dyo, xi, w
            del x
            torch.cuda.empty_cache()
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    meta = dict(peaks=peaks, reference_elementwise=ref_how, reference_conv='cuDNN fp32 through the reference modulated_conv2d (non-fused), TF32 off, cudnn.benchmark on',
                gpu=torch.cuda.get_device_name(0), seconds=time.time() - t_start)
    json.dump(dict(meta=meta, rows=rows), open(args.out + '.json', 'w'))
    # compact table: one line per (op, r) with the fraction of roofline at each N and the geometric-mean speed-up over the reference
    with open(args.out + '.txt', 'w') as fh:
        fh.write(f'cfg 5 sweep on {meta["gpu"]}: percent of roofline (HBM {hbm:.0f} GB/s measured copy; TF32 {tf32:.0f} TFLOP/s = measured bf16 / 2; an fp32-faithful\n'
                 f'conv needs 3 TF32 products per MAC, so 33 % is its ceiling) at N = {batches}, then the speed-up over the reference on the same GPU\n'
                 f'(convolutions: {meta["reference_conv"]}; elementwise: {ref_how})\n\n')
        ops = []
        for row in rows:
            if row['op'] not in ops:
                ops.append(row['op'])
        for op in ops:
            fh.write(op + '\n')
            for r in sorted({row['r'] for row in rows if row['op'] == op}):
                pts = {row['N']: row for row in rows if row['op'] == op and row['r'] == r}
                fr = ' '.join(f"{100 * pts[n]['frac']:5.1f}" if n in pts and 'frac' in pts[n] else '    -' for n in batches)
                sp = ' '.join(f"{pts[n]['speedup']:5.1f}" if n in pts and 'speedup' in pts[n] else '    -' for n in batches)
                fh.write(f'   r={r:<5d} %roofline: {fr}   | x reference: {sp}\n')
    print(open(args.out + '.txt').read())


if __name__ == '__main__':
    main()
