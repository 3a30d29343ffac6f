#!/usr/bin/env python
"""Gmain parameter gradients against the golden fixture, per parameter (development tool)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'ga-gan_b200')):
    sys.path.insert(0, p)
import numpy as np, torch
from tests.util import load_golden, t, patched_randn, max_rel_err
from training import networks
from training.loss import StyleGAN2Loss
device = torch.device('cuda:0')
g = load_golden('networks')
cfg = {kv.split('=')[0]: int(kv.split('=')[1]) for kv in (str(m) for m in g['meta'])}
G = networks.Generator(z_dim=cfg['z_dim'], c_dim=0, w_dim=cfg['w_dim'], img_resolution=cfg['res'], img_channels=3,
                       mapping_kwargs=dict(num_layers=cfg['num_layers']),
                       synthesis_kwargs=dict(channel_base=cfg['channel_base'], channel_max=cfg['channel_max']))
D = networks.Discriminator(c_dim=0, img_resolution=cfg['res'], img_channels=3, channel_base=cfg['channel_base'],
                           channel_max=cfg['channel_max'], epilogue_kwargs=dict(mbstd_group_size=cfg['mbstd']))
G.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('G.')}, strict=False)
D.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('D.')}, strict=False)
G, D = G.to(device).train(), D.to(device).train()
G.mapping.w_avg_beta = None
phase = sys.argv[1] if len(sys.argv) > 1 else 'Gmain'
for p in list(G.parameters()) + list(D.parameters()):
    p.requires_grad_(True); p.grad = None
loss = StyleGAN2Loss(device=device, G_mapping=G.mapping, G_synthesis=G.synthesis, D=D, style_mixing_prob=0, r1_gamma=10, pl_batch_shrink=2, pl_decay=0.01, pl_weight=2)
z = t(g['z'], device); real = t(g['real'], device); c = torch.zeros(z.shape[0], 0, device=device)
with patched_randn(11):
    loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, sync=True, gain=1.0)
net = G if phase[0] == 'G' else D
rows = []
for k, p in net.named_parameters():
    want = g[f'{phase}.grad.{k}']
    if np.abs(want).max() == 0: continue
    got = p.grad.detach().cpu().numpy() if p.grad is not None else np.zeros_like(want)
    rows.append((float(np.abs(got - want).max() / np.abs(want).max()), k, float(np.abs(want).max()), want.size))
for e, k, m, n in sorted(rows, reverse=True)[:12]:
    print(f'{e:.3e}  {k:45s} max|want| {m:.3e}  numel {n}')
