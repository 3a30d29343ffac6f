#!/usr/bin/env python
"""CUDA-event time of each loss phase (+ optimiser, EMA) of the config-f 1024^2 training iteration, batch 32.
    python tools/phase_times.py [--res 1024] [--batch 32]"""
import os, sys, argparse
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gagan_b200
gagan_b200.install(os.path.join(ROOT, 'baseline', '_ref', 'DissimilarDomains'))
import torch
from torch_utils import custom_ops
from gagan_b200.training import training_loop
ap = argparse.ArgumentParser(); ap.add_argument('--res', type=int, default=1024); ap.add_argument('--batch', type=int, default=32)
ap.add_argument('--cfg', default='stylegan2'); args = ap.parse_args()
dev = torch.device('cuda:0')
custom_ops.verbosity = 'none'
spec = training_loop.CONFIGS[args.cfg]
torch.manual_seed(0)
G, D = training_loop.build_networks(args.res, args.cfg, device=dev)
step = training_loop.TrainingStep(G, D, batch_size=args.batch, batch_gpu=args.batch, device=dev, lrate=spec['lrate'], r1_gamma=spec['gamma'], ema_kimg=spec['ema'])
real = torch.rand(args.batch, 3, args.res, args.res, device=dev) * 2 - 1
orig = step.loss.accumulate_gradients
times = {}
def timed(phase, **kw):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = custom_ops.launch_count()
    e0.record(); orig(phase=phase, **kw); e1.record()
    times.setdefault(phase, []).append((e0, e1, custom_ops.launch_count() - l0))
step.loss.accumulate_gradients = timed
for _ in range(2):
    step.cur_it = 0; step.run(real)
times.clear()
tot0, tot1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
tot0.record()
for _ in range(3):
    step.cur_it = 0; step.run(real)
tot1.record(); torch.cuda.synchronize()
ms = {p: sum(a.elapsed_time(b) for a, b, _ in v) / len(v) for p, v in times.items()}
ln = {p: v[0][2] for p, v in times.items()}
total = tot0.elapsed_time(tot1) / 3
print(f'all four phases + optimiser + EMA: {total:.1f} ms; phases: ' + ', '.join(f'{p} {m:.1f} ms ({ln[p]} library launches)' for p, m in ms.items()))
am = ms['Gmain'] + ms['Dmain'] + ms['Greg'] / 4 + ms['Dreg'] / 16
print(f'amortised iteration (Gmain + Dmain + Greg/4 + Dreg/16): {am:.1f} ms -> {args.batch / am * 1000:.1f} img/s; outside the phases: {total - sum(ms.values()):.1f} ms')
