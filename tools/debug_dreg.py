#!/usr/bin/env python
"""Debug harness: Dreg parameter gradients on the GPU under several toggles (conv precision, un-fused db,
oracle upfirdn2d on the device) to localise a deviation from the golden gradients.  Test tooling only."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'ga-gan_b200')):
    sys.path.insert(0, p)
import torch
from tests.util import load_golden, t, max_rel_err, patched_randn
from torch_utils import custom_ops
from torch_utils.ops import bias_act as ba_mod, upfirdn2d as up_mod, conv2d_gradfix
from training import networks
from training.loss import StyleGAN2Loss
from oracle import ops_ref as R

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
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
G.to(dev).train(); D.to(dev).train()
G.mapping.w_avg_beta = None
plugin = custom_ops.get_plugin('bias_act_plugin')


def run(phase, tag):
    for p in list(G.parameters()) + list(D.parameters()):
        p.requires_grad_(True); p.grad = None
    loss = StyleGAN2Loss(device=dev, G_mapping=G.mapping, G_synthesis=G.synthesis, D=D, style_mixing_prob=0,
                         r1_gamma=10, pl_batch_shrink=2, pl_decay=0.01, pl_weight=2)
    z = t(g['z'], dev); real = t(g['real'], dev); c = torch.zeros(z.shape[0], 0, device=dev)
    with patched_randn(11):
        loss.accumulate_gradients(phase=phase, real_img=real, real_c=c, gen_z=z, gen_c=c, sync=True, gain=1.0)
    net = G if phase[0] == 'G' else D
    errs = []
    for k, p in net.named_parameters():
        want = g[f'{phase}.grad.{k}']
        if np.abs(want).max() == 0:
            continue
        errs.append((max_rel_err(p.grad, want), k))
    errs.sort(reverse=True)
    print(f'[{tag}] {phase}: ' + '  '.join(f'{k}={e:.2e}' for e, k in errs[:4]), flush=True)
    return {k: p.grad.detach().clone() for k, p in net.named_parameters() if p.grad is not None}


phases = sys.argv[1:] or ['Dreg']
for ph in phases:
    ga = run(ph, 'default')
    custom_ops.conv_precision = custom_ops.PREC_FP32_SIMT
    gb = run(ph, 'conv=fp32_simt')
    custom_ops.conv_precision = custom_ops.PREC_AUTO
    key = os.environ.get('GG_DEBUG_PARAM', '')
    if key and key in ga:
        print(f'   {key}: default {ga[key].flatten()[:4].tolist()}  simt {gb[key].flatten()[:4].tolist()}  golden {g[ph + ".grad." + key].flatten()[:4].tolist()}')

    conv2d_gradfix.fuse_scales = False
    run(ph, 'scales unfused')
    conv2d_gradfix.fuse_scales = True

    orig_ba = plugin.bias_act.__func__

    def ba_unfused(self, x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp, dbias=None):
        y = orig_ba(self, x, b, xref, yref, dy, grad, dim, act, alpha, gain, clamp, dbias=None)
        if dbias is not None:
            dbias += y.sum([i for i in range(y.ndim) if i != dim])
        return y
    type(plugin).bias_act = ba_unfused
    run(ph, 'db=torch.sum')
    type(plugin).bias_act = orig_ba

    orig_up = type(plugin).upfirdn2d

    def up_ref(self, x, f, upx, upy, downx, downy, padx0, padx1, pady0, pady1, flip, gain):
        return R.upfirdn2d(x, f, up=[upx, upy], down=[downx, downy], padding=[padx0, padx1, pady0, pady1], flip_filter=flip, gain=gain)
    type(plugin).upfirdn2d = up_ref
    run(ph, 'upfirdn2d=oracle-on-gpu')
    type(plugin).upfirdn2d = orig_up
