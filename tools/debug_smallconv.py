#!/usr/bin/env python
"""Debug: inside the Gmain phase of the golden test network, compare every small-map conv call (tcgen05 vs FFMA) on the
actual data."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'ga-gan_b200')):
    sys.path.insert(0, p)
import torch
from tests.util import load_golden, t, patched_randn
from torch_utils import custom_ops
from training import networks
from training.loss import StyleGAN2Loss
dev = torch.device('cuda:0')
g = load_golden('networks')
cfg = {kv.split('=')[0]: int(kv.split('=')[1]) for kv in (str(m) for m in g['meta'])}
G = networks.Generator(z_dim=cfg['z_dim'], c_dim=0, w_dim=cfg['w_dim'], img_resolution=cfg['res'], img_channels=3,
                       mapping_kwargs=dict(num_layers=cfg['num_layers']),
                       synthesis_kwargs=dict(channel_base=cfg['channel_base'], channel_max=cfg['channel_max']))
D = networks.Discriminator(c_dim=0, img_resolution=cfg['res'], img_channels=3, channel_base=cfg['channel_base'],
                           channel_max=cfg['channel_max'], epilogue_kwargs=dict(mbstd_group_size=cfg['mbstd']))
G.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('G.')}, strict=False)
D.load_state_dict({k[2:]: t(v) for k, v in g.items() if k.startswith('D.')}, strict=False)
G.to(dev).train(); D.to(dev).train(); G.mapping.w_avg_beta = None
plugin = custom_ops.get_plugin('conv2d_plugin')
orig = plugin.conv2d

def conv2d(x, w, **kw):
    y = orig(x, w, **kw)
    if min(y.shape[2:]) < 8 and plugin.last_conv_prec == 3:
        kw2 = dict(kw); kw2['prec'] = custom_ops.PREC_FP32_SIMT
        y2 = orig(x, w, **kw2)
        xd, wd = x.double(), w.double()
        d = (y - y2).abs().max().item(); s = y2.abs().max().item()
        print(f'x{tuple(x.shape)} w{tuple(w.shape)} kw={ {k: v for k, v in kw.items() if k != "flop_scale"} } -> {tuple(y.shape)}: '
              f'max|tc-simt| {d:.3e} / max|y| {s:.3e} = {d / max(s, 1e-30):.2e}; x absmax {x.abs().max().item():.2e} w absmax {w.abs().max().item():.2e} '
              f'nan={bool(torch.isnan(y).any())}', flush=True)
    return y
plugin.conv2d = conv2d
for p in list(G.parameters()) + list(D.parameters()):
    p.requires_grad_(True)
loss = StyleGAN2Loss(device=dev, G_mapping=G.mapping, G_synthesis=G.synthesis, D=D, style_mixing_prob=0, r1_gamma=10, pl_batch_shrink=2,
                     pl_decay=0.01, pl_weight=2)
z = t(g['z'], dev); real = t(g['real'], dev); c = torch.zeros(z.shape[0], 0, device=dev)
with patched_randn(11):
    loss.accumulate_gradients(phase='Gmain', real_img=real, real_c=c, gen_z=z, gen_c=c, sync=True, gain=1.0)
