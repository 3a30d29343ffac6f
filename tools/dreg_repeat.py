#!/usr/bin/env python
"""Run-to-run variation of the Dreg parameter gradients (same weights, same data, same process): is the distance to the fp64
truth stable, and which parameters move?      python tools/dreg_repeat.py [reps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tests.util as U
from oracle import live_ref, networks_ref as NR
from torch_utils import custom_ops
from training import networks, loss as loss_mod
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 6
family = int(sys.argv[2]) if len(sys.argv) > 2 else 1
custom_ops.set_conv_kernel_family(family)
if len(sys.argv) > 3:
    custom_ops.conv_precision = int(sys.argv[3])
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
L = live_ref.load()
kw = dict(c_dim=0, img_resolution=256, img_channels=3, channel_base=16384, channel_max=512, num_fp16_res=0, conv_clamp=None,
          epilogue_kwargs=dict(mbstd_group_size=4))
torch.manual_seed(1)
D_cpu = U.quiet(L.networks.Discriminator, **kw).train()
with torch.no_grad():
    for p in D_cpu.parameters():
        if float(p.abs().max()) == 0:
            p.copy_(torch.randn(p.shape) * 0.1)
D = U.quiet(networks.Discriminator, **kw).train()
D.load_state_dict(D_cpu.state_dict()); D = D.to(dev); D.requires_grad_(True)
real = torch.rand(4, 3, 256, 256) * 2 - 1
c = torch.zeros(4, 0); z = torch.zeros(4, 512)
names = [n for n, _ in D_cpu.named_parameters()]
PD = {k: v.detach().double().requires_grad_(v.dtype.is_floating_point) for k, v in D_cpu.state_dict().items()}
truth = {n: g for n, g in zip(names, torch.autograd.grad(NR.loss_Dr1(PD, real.double(), 256, 1.0, 4) * 16, [PD[n] for n in names], allow_unused=True)) if g is not None}


def rel(a, b):
    d = float(b.abs().max())
    return float((a.detach().double().cpu() - b).abs().max()) / d if d > 0 else float('nan')


def one():
    for p in D.parameters():
        p.grad = None
    Lg = loss_mod.StyleGAN2Loss(device=dev, G_mapping=None, G_synthesis=None, D=D, r1_gamma=1.0)
    Lg.accumulate_gradients(phase='Dreg', real_img=real.to(dev), real_c=c.to(dev), gen_z=z.to(dev), gen_c=c.to(dev), sync=True, gain=16)
    torch.cuda.synchronize()
    return {n: p.grad.detach().clone() for n, p in D.named_parameters() if p.grad is not None and n in truth}


first = None
for r in range(reps):
    if r == reps // 2:       # churn the allocator and the caches with unrelated heavy work, as a long test session does
        from torch_utils.ops import conv2d_resample, upfirdn2d
        f = upfirdn2d.setup_filter([1, 3, 3, 1]).to(dev)
        for (N, I, O, R, up, down) in [(1, 32, 32, 1024, 1, 1), (2, 128, 64, 256, 2, 1), (2, 64, 128, 512, 1, 2), (2, 512, 512, 64, 1, 1)]:
            x = torch.randn(N, I, R, R, device=dev, requires_grad=True); w = torch.randn(O, I, 3, 3, device=dev, requires_grad=True)
            y = conv2d_resample.conv2d_resample(x, w, f=f, up=up, down=down, padding=1, flip_weight=(up == 1))
            torch.autograd.grad(y, [x, w], torch.randn_like(y))
        print('-- churned', flush=True)
    g = one()
    rows = sorted(((rel(v, truth[n]), n) for n, v in g.items()), reverse=True)
    if first is None:
        first = g
    moved = sorted(((rel(v, first[n].double().cpu()), n) for n, v in g.items()), reverse=True)
    print(f'rep {r}: worst vs truth {rows[0][0]:.2e} {rows[0][1]} | {rows[1][0]:.2e} {rows[1][1]} | {rows[2][0]:.2e} {rows[2][1]}   '
          f'|| moved vs rep 0: {moved[0][0]:.2e} {moved[0][1]} | {moved[1][0]:.2e} {moved[1][1]}', flush=True)
