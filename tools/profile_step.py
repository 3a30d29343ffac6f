#!/usr/bin/env python
"""Where does one training iteration spend its GPU time?  torch.profiler (CUPTI) kernel table of one bench.py step
(config-f, --res, batch/batch_gpu as given), grouped by kernel name.  Development tool; the judged numbers come from
bench.py and the ncu captures under profiles/.

    python tools/profile_step.py [--res 1024] [--batch 8] [--batch-gpu 4] [--out gpurun_out/step_profile.txt]
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

import numpy as np  # noqa: E402
import torch  # noqa: E402
from torch.profiler import profile, ProfilerActivity  # noqa: E402


def profile_ga(args):
    from torch_utils import custom_ops
    from gagan_b200.training import training_loop, ga_eval
    dev = torch.device('cuda:0')
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    custom_ops.verbosity = 'none'
    torch.manual_seed(0)
    G, D = training_loop.build_networks(256, 'paper256', device=dev, use_domain_modulation=True, domain_modulation_parametrization='additive')
    pop = ga_eval.init_population(G, 4, seed=0)
    z = torch.randn(8, 512, generator=torch.Generator().manual_seed(1)).to(dev)
    for _ in range(2):
        ga_eval.evaluate_population(G, D, pop, z, cuda_graph=False)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        ga_eval.evaluate_population(G, D, pop, z, cuda_graph=False)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '')
            agg[name[:110]][0] += 1
            agg[name[:110]][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
    total = sum(v[1] for v in agg.values())
    lines = [f'GA fitness evaluation, 4 individuals x 8 latents at 256^2 (paper256): {total / 1000:.2f} ms of kernel time in {sum(v[0] for v in agg.values())} launches']
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        lines.append(f'{us / 1000:10.3f} ms {100 * us / total:5.1f}%  {n:6d}x  {name}')
    text = '\n'.join(lines)
    print(text)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        open(args.out, 'w').write(text + '\n')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--res', type=int, default=1024)
    ap.add_argument('--cfg', default='stylegan2')
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--batch-gpu', type=int, default=4)
    ap.add_argument('--out', default='')
    ap.add_argument('--main-only', action='store_true', help='profile an iteration without the lazy regularisation phases')
    ap.add_argument('--ga', action='store_true', help='profile the GA population fitness evaluation (BASELINE configs[3]) instead: paper256, 4 individuals x 8 latents')
    args = ap.parse_args()
    if args.ga:
        return profile_ga(args)
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
    # which conv shapes miss the tensor-core path?
    plugin = custom_ops.get_plugin('conv2d_plugin')
    shapes = collections.Counter()
    orig_conv, orig_wgrad = plugin.conv2d, plugin.conv2d_wgrad

    def conv2d(x, w, **kw):
        y = orig_conv(x, w, **kw)
        shapes[('conv', tuple(x.shape), tuple(w.shape), kw.get('stride', 1), tuple(kw.get('padding', (0, 0))), bool(kw.get('transposed', False)),
                tuple(y.shape[2:]), plugin.last_conv_prec)] += 1
        return y

    def conv2d_wgrad(a, b, ks, **kw):
        y = orig_wgrad(a, b, ks, **kw)
        shapes[('wgrad', tuple(a.shape), tuple(b.shape), kw.get('stride', 1), tuple(kw.get('padding', (0, 0))), False, tuple(ks),
                plugin.last_wgrad_prec)] += 1
        return y
    plugin.conv2d, plugin.conv2d_wgrad = conv2d, conv2d_wgrad
    step.cur_it = 0
    step.run(real)
    plugin.conv2d, plugin.conv2d_wgrad = orig_conv, orig_wgrad
    print('conv shapes NOT on the tcgen05 path (kind, x, w, stride, pad, transposed, out/k, prec) x count:')
    for k, n in sorted(shapes.items(), key=lambda kv: -int(np.prod(kv[0][1])) * kv[1]):
        if k[-1] == 0:
            print('   ', k, 'x', n)
    torch.cuda.synchronize()
    step.cur_it = 1 if args.main_only else 0      # 0: all four phases fire
    import time
    t0 = time.perf_counter()
    step.run(real); torch.cuda.synchronize()
    wall_ms = (time.perf_counter() - t0) * 1000
    step.cur_it = 1 if args.main_only else 0
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        step.run(real)
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            name = ev.name.replace('(anonymous namespace)::', '').replace('void ', '')
            agg[name[:110]][0] += 1
            agg[name[:110]][1] += ev.device_time if hasattr(ev, 'device_time') else ev.cuda_time
    total = sum(v[1] for v in agg.values())
    lines = [f'one iteration ({"Gmain+Dmain" if args.main_only else "all four phases"}), res {args.res}, batch {args.batch} in rounds of {args.batch_gpu}: '
             f'{total / 1000:.1f} ms of kernel time in {sum(v[0] for v in agg.values())} launches; wall clock of the same iteration without the profiler {wall_ms:.1f} ms']
    for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
        lines.append(f'{us / 1000:10.2f} ms {100 * us / total:5.1f}%  {n:6d}x  {name}')
    # where does the GPU wait for the host?  gaps between consecutive kernels on the timeline, attributed to the kernel that follows
    evs = sorted([(ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events() if ev.device_type == torch.autograd.DeviceType.CUDA],
                 key=lambda t: t[0])
    gaps = collections.defaultdict(lambda: [0, 0.0])
    idle, busy_end = 0.0, None
    big = []
    for st, en, name in evs:
        if busy_end is not None and st > busy_end:
            g = st - busy_end
            idle += g
            if g > 15:
                key = name.replace('(anonymous namespace)::', '').replace('void ', '')[:70]
                gaps[key][0] += 1; gaps[key][1] += g
                big.append((g, key))
        busy_end = en if busy_end is None else max(busy_end, en)
    span = evs[-1][1] - evs[0][0]
    lines.append(f'timeline: first kernel start -> last kernel end {span / 1000:.1f} ms, of which no kernel is running for {idle / 1000:.1f} ms; gaps > 15 us by the kernel that ends them:')
    for key, (n, us) in sorted(gaps.items(), key=lambda kv: -kv[1][1])[:14]:
        lines.append(f'{us / 1000:10.2f} ms  {n:5d}x  before {key}')
    lines.append('largest single gaps (us): ' + ', '.join(f'{g:.0f} [{k[:40]}]' for g, k in sorted(big, reverse=True)[:8]))
    text = '\n'.join(lines)
    print(text)
    if args.out:
        os.makedirs(os.path.dirname(args.out) or '.', exist_ok=True)
        open(args.out, 'w').write(text + '\n')


if __name__ == '__main__':
    main()
