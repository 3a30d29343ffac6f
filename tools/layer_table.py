#!/usr/bin/env python
"""Per-call table of one training iteration: every entry point of libgagan_b200.so is wrapped with CUDA events and
the calls are grouped by (entry point, shapes).  Answers "which layer shape costs what" -- the launch list of ncu only
has kernel names and grids.  Development tool.

    python tools/layer_table.py [--res 1024] [--batch 32] [--batch-gpu 32] [--main-only] [--out gpurun_out/layer_table.txt]
"""
import os
import sys
import argparse
import collections

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200  # noqa: E402
_CHECKOUT = os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains')
gagan_b200.install(_CHECKOUT if os.path.isdir(_CHECKOUT) else None)   # the reference checkout on this build's operators

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=1024)
    ap.add_argument('--cfg', default='stylegan2')
    ap.add_argument('--batch', type=int, default=32)
    ap.add_argument('--batch-gpu', type=int, default=32)
    ap.add_argument('--main-only', action='store_true')
    ap.add_argument('--out', default='')
    args = ap.parse_args()
    from torch_utils import custom_ops
    from gagan_b200.training import training_loop
    dev = torch.device('cuda:0')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    custom_ops.verbosity = 'none'
    spec = training_loop.CONFIGS[args.cfg]
    torch.manual_seed(0)
    G, D = training_loop.build_networks(args.res, args.cfg, device=dev)
    step = training_loop.TrainingStep(G, D, batch_size=args.batch, batch_gpu=min(args.batch_gpu, args.batch), device=dev,
                                      lrate=spec['lrate'], r1_gamma=spec['gamma'], ema_kimg=spec['ema'])
    real = torch.rand(args.batch, 3, args.res, args.res, device=dev) * 2 - 1
    for _ in range(2):
        step.cur_it = 0
        step.run(real)
    torch.cuda.synchronize()

    plugins = {n: custom_ops.get_plugin(n) for n in ('conv2d_plugin', 'upfirdn2d_plugin', 'bias_act_plugin')}
    plugin = plugins['conv2d_plugin']
    records = []

    def shp(t):
        return 'x'.join(str(int(s)) for s in t.shape) if isinstance(t, torch.Tensor) else str(t)

    def wrap(pname, name, describe):
        orig = getattr(plugins[pname], name)

        def fn(*a, **kw):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig(*a, **kw)
            e1.record()
            key, work = describe(out, *a, **kw)
            records.append((name, key, work, e0, e1))
            return out
        setattr(plugins[pname], name, fn)
        return orig

    def d_conv(y, x, w, **kw):
        tr = bool(kw.get('transposed', False))
        k = w.shape[2]
        N, I = x.shape[0], x.shape[1]
        O = w.shape[1] if tr else w.shape[0]
        fl = kw.get('flop_scale', 1.0) * 2.0 * N * O * I * k * k * y.shape[2] * y.shape[3]
        sc = ('s' if kw.get('in_scale') is not None else '-') + ('d' if kw.get('out_scale') is not None else '-')
        return f'{"T" if tr else " "} x{shp(x)} w{shp(w)} -> {shp(y)} {sc} prec{plugin.last_conv_prec}', ('flop', fl)

    def d_wgrad(dw, a, b, ks, **kw):
        fl = kw.get('flop_scale', 1.0) * 2.0 * a.shape[0] * a.shape[1] * b.shape[1] * ks[0] * ks[1] * b.shape[2] * b.shape[3]
        return f'a{shp(a)} b{shp(b)} k{ks[0]} prec{plugin.last_wgrad_prec}', ('flop', fl)

    def d_fir(y, x, f, *a, **kw):
        return f'x{shp(x)} -> {shp(y)} in_pm={kw.get("in_pm") is not None} out_pm={kw.get("out_pm") is not None}', ('byte', 4.0 * (x.numel() + y.numel()))

    def d_up(y, x, f, upx, upy, downx, downy, *a):
        return f'x{shp(x)} -> {shp(y)} up{upx} down{downx} f{shp(f)}', ('byte', 4.0 * (x.numel() + y.numel()))

    def d_ba(y, x, b, xref, yref, dy, grad, *a, **kw):
        nb = 8.0 if grad == 0 else 12.0
        return f'x{shp(x)} grad{grad} act{a[1] if len(a) > 1 else "?"} db={kw.get("dbias") is not None}', ('byte', nb * x.numel())

    def d_ban(y, x, b, noise, *a):
        return f'x{shp(x)} noise{shp(noise)}', ('byte', 8.0 * x.numel())

    def d_dot(out, a, b):
        return f'a{shp(a)}', ('byte', 8.0 * a.numel())

    table = (('conv2d_plugin', 'conv2d', d_conv), ('conv2d_plugin', 'conv2d_wgrad', d_wgrad), ('conv2d_plugin', 'chan_dot', d_dot),
             ('upfirdn2d_plugin', 'fir4_pm', d_fir), ('upfirdn2d_plugin', 'upfirdn2d', d_up),
             ('bias_act_plugin', 'bias_act', d_ba), ('bias_act_plugin', 'bias_act_noise', d_ban))
    origs = {(pn, n): wrap(pn, n, d) for pn, n, d in table}
    step.cur_it = 1 if args.main_only else 0
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    step.run(real)
    t1.record()
    torch.cuda.synchronize()
    for (pn, n), o in origs.items():
        setattr(plugins[pn], n, o)
    total_ms = t0.elapsed_time(t1)

    agg = collections.OrderedDict()
    for name, key, work, e0, e1 in records:
        a = agg.setdefault((name, key), [0, 0.0, 0.0, work[0]])
        a[0] += 1
        a[1] += e0.elapsed_time(e1)
        a[2] += work[1]
    lines = [f'one iteration ({"Gmain+Dmain" if args.main_only else "all four phases"}), res {args.res}, batch {args.batch} in rounds of '
             f'{min(args.batch_gpu, args.batch)}: {total_ms:.1f} ms with per-call events; {len(records)} library calls, '
             f'{sum(a[1] for a in agg.values()):.1f} ms inside them']
    by_name = collections.defaultdict(float)
    for (name, key), a in agg.items():
        by_name[name] += a[1]
    for name, ms in sorted(by_name.items(), key=lambda kv: -kv[1]):
        lines.append(f'  {name:16s} {ms:9.2f} ms  {100 * ms / total_ms:5.1f}%')
    lines.append('')
    for (name, key), a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        rate = a[2] / a[1] / 1e9 if a[3] == 'flop' else a[2] / a[1] / 1e6
        unit = 'TFLOP/s' if a[3] == 'flop' else 'GB/s'
        lines.append(f'{a[1]:9.2f} ms {100 * a[1] / total_ms:5.1f}% {a[0]:4d}x {a[1] / a[0]:8.3f} ms/call {rate:9.1f} {unit:7s} {name:14s} {key}')
    text = '\n'.join(lines)
    print(text)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        open(args.out, 'w').write(text + '\n')


if __name__ == '__main__':
    main()
