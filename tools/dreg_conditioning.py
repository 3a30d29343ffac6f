#!/usr/bin/env python
"""How well-conditioned are the Dreg (R1) parameter gradients at config size?  For several data seeds: distance to the fp64 truth of
  (a) this build (marching + tile kernels), (b) this build with the tile kernel everywhere, (c) this build's exact FFMA kernels,
  (d) the REFERENCE itself on the GPU (its impl='ref' torch ops = cuDNN fp32, TF32 off), (e) the reference on the CPU,
and the change of the fp64 truth itself when the input image is perturbed by 1e-7 relative (one fp32 rounding).
    python tools/dreg_conditioning.py [seeds]"""
import os, sys, time, warnings
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tests.util as U
from oracle import live_ref, networks_ref as NR
from torch_utils import custom_ops
from training import networks, loss as loss_mod
nseeds = int(sys.argv[1]) if len(sys.argv) > 1 else 3
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
warnings.filterwarnings('ignore')
L = live_ref.load()
L.conv2d_gradfix.enabled = True
kw = dict(c_dim=0, img_resolution=256, img_channels=3, channel_base=16384, channel_max=512, num_fp16_res=0, conv_clamp=None,
          epilogue_kwargs=dict(mbstd_group_size=4))


def rel(a, b):
    d = float(b.abs().max())
    return float((a.detach().double().cpu() - b).abs().max()) / d if d > 0 else float('nan')


def run(loss_cls, D, real, device):
    for p in D.parameters():
        p.grad = None
    D.requires_grad_(True)
    c = torch.zeros(4, 0, device=device); z = torch.zeros(4, 512, device=device)
    loss_cls(device=device, G_mapping=None, G_synthesis=None, D=D, r1_gamma=1.0).accumulate_gradients(
        phase='Dreg', real_img=real.to(device), real_c=c, gen_z=z, gen_c=c, sync=True, gain=16)
    return {n: p.grad.detach().double().cpu() for n, p in D.named_parameters() if p.grad is not None}


print('worst / median over parameters of max|g - truth| / max|truth|   (bias gradients are the worst: they only flow through the minibatch-std layer)')
for seed in range(1, nseeds + 1):
    torch.manual_seed(seed)
    D_cpu = U.quiet(L.networks.Discriminator, **kw).train()
    with torch.no_grad():
        for p in D_cpu.parameters():
            if float(p.abs().max()) == 0:
                p.copy_(torch.randn(p.shape) * 0.1)
    real = torch.rand(4, 3, 256, 256) * 2 - 1
    names = [n for n, _ in D_cpu.named_parameters()]

    def truth_of(img):
        PD = {k: v.detach().double().requires_grad_(v.dtype.is_floating_point) for k, v in D_cpu.state_dict().items()}
        gs = torch.autograd.grad(NR.loss_Dr1(PD, img, 256, 1.0, 4) * 16, [PD[n] for n in names], allow_unused=True)
        return {n: g for n, g in zip(names, gs) if g is not None and float(g.abs().max()) > 0}
    truth = truth_of(real.double())
    pert = truth_of(real.double() * (1 + 1e-7 * torch.randn(real.shape, dtype=torch.float64)))
    res = {'fp64 truth, image perturbed by 1e-7': pert}
    for label, fam, prec in (('this build (default)', 1, custom_ops.PREC_AUTO), ('this build, tile kernel only', 0, custom_ops.PREC_AUTO),
                             ('this build, FFMA fp32', 1, custom_ops.PREC_FP32_SIMT)):
        custom_ops.set_conv_kernel_family(fam); custom_ops.conv_precision = prec
        D = U.quiet(networks.Discriminator, **kw).train(); D.load_state_dict(D_cpu.state_dict())
        res[label] = run(loss_mod.StyleGAN2Loss, D.to(dev), real, dev)
    custom_ops.set_conv_kernel_family(1); custom_ops.conv_precision = custom_ops.PREC_AUTO
    D_ref_gpu = U.quiet(L.networks.Discriminator, **kw).train(); D_ref_gpu.load_state_dict(D_cpu.state_dict())
    try:
        res['reference on the GPU (cuDNN fp32, impl=ref ops)'] = run(L.loss.StyleGAN2Loss, D_ref_gpu.to(dev), real, dev)
    except Exception as e:
        print('reference on the GPU failed:', str(e)[:200])
    res['reference on the CPU (fp32)'] = run(L.loss.StyleGAN2Loss, D_cpu, real, torch.device('cpu'))
    print(f'seed {seed}:')
    for label, g in res.items():
        rows = sorted(((rel(g[n], truth[n]), n) for n in truth if n in g), reverse=True)
        wts = [r[0] for r in rows if r[1].endswith('weight')]
        print(f'   {label:50s} worst {rows[0][0]:.2e} ({rows[0][1]})  2nd {rows[1][0]:.2e}  median {np.median([r[0] for r in rows]):.2e}  worst weight {max(wts):.2e}', flush=True)
