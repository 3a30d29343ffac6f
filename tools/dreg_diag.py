#!/usr/bin/env python
"""Where does the Dreg (R1, double backward) gradient error at config size come from?  paper256 256^2 discriminator, batch 4:
parameter gradients of the reference's StyleGAN2Loss('Dreg') on the library in several arithmetic modes vs the fp64 oracle.
    python tools/dreg_diag.py [res] [channel_max]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import tests.util as U
from oracle import live_ref, networks_ref as NR
from torch_utils import custom_ops
from training import networks, loss as loss_mod
res = int(sys.argv[1]) if len(sys.argv) > 1 else 256
cmax = int(sys.argv[2]) if len(sys.argv) > 2 else 512
dev = torch.device('cuda:0')
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
L = live_ref.load()
kw = dict(c_dim=0, img_resolution=res, img_channels=3, channel_base=16384, channel_max=cmax, num_fp16_res=0, conv_clamp=None,
          epilogue_kwargs=dict(mbstd_group_size=4))
torch.manual_seed(1)
D_cpu = U.quiet(L.networks.Discriminator, **kw).train()
with torch.no_grad():
    for p in D_cpu.parameters():
        if float(p.abs().max()) == 0:
            p.copy_(torch.randn(p.shape) * 0.1)
real = torch.rand(4, 3, res, res) * 2 - 1
c = torch.zeros(4, 0); z = torch.zeros(4, 512)
names = [n for n, _ in D_cpu.named_parameters()]
t0 = time.time()
PD = {k: v.detach().double().requires_grad_(v.dtype.is_floating_point) for k, v in D_cpu.state_dict().items()}
img64 = real.double().requires_grad_(True)
logits64 = NR.discriminator(PD, img64, res, mbstd_group_size=4)
g64, = torch.autograd.grad(logits64.sum(), img64, create_graph=True)
pen64 = g64.square().sum([1, 2, 3])
truth = {n: g for n, g in zip(names, torch.autograd.grad((pen64 * 0.5).mean() * 16, [PD[n] for n in names], allow_unused=True)) if g is not None}
print(f'fp64 oracle: {time.time() - t0:.1f} s', flush=True)
L.conv2d_gradfix.enabled = True
D_cpu.requires_grad_(True)
L.loss.StyleGAN2Loss(device=torch.device('cpu'), G_mapping=None, G_synthesis=None, D=D_cpu, r1_gamma=1.0).accumulate_gradients(
    phase='Dreg', real_img=real, real_c=c, gen_z=z, gen_c=c, sync=True, gain=16)
ref = {n: p.grad.double() for n, p in D_cpu.named_parameters() if p.grad is not None}


def rel(a, b):
    d = float(b.abs().max())
    return float((a.double().cpu() - b).abs().max()) / d if d > 0 else float('nan')


for label, prec, fam in (('3xTF32 tile', custom_ops.PREC_AUTO, 0), ('3xTF32 march', custom_ops.PREC_AUTO, 1), ('FFMA fp32', custom_ops.PREC_FP32_SIMT, 0)):
    custom_ops.conv_precision = prec
    custom_ops.set_conv_kernel_family(fam)
    D = U.quiet(networks.Discriminator, **kw).train()
    D.load_state_dict(D_cpu.state_dict()); D = D.to(dev); D.requires_grad_(True)
    img = real.to(dev).requires_grad_(True)
    logits = D(img, c.to(dev))
    g, = torch.autograd.grad(logits.sum(), img, create_graph=True)
    pen = g.square().sum([1, 2, 3])
    grads = torch.autograd.grad((pen * 0.5).mean() * 16, list(D.parameters()), allow_unused=True)
    torch.cuda.synchronize()
    rows = sorted(((rel(gr, truth[n]), rel(ref[n], truth[n]), n) for n, gr in zip(names, grads) if gr is not None and n in truth and n in ref), reverse=True)
    print(f'== {label}: logits {rel(logits, logits64.detach()):.2e}  dlogits/dimg {rel(g, g64.detach()):.2e}  r1 penalty {rel(pen, pen64.detach()):.2e}')
    for e, er, n in rows[:6] + rows[len(rows) // 2:len(rows) // 2 + 3] + rows[-3:]:
        print(f'   {e:.2e} (cpu fp32 reference {er:.2e})  {n}')
    med = np.median([r[0] for r in rows]); print(f'   median {med:.2e}', flush=True)
custom_ops.conv_precision = custom_ops.PREC_AUTO

# the same phase through the reference's loss class (what the parity test and the training step run)
for label, nwg in (('StyleGAN2Loss.accumulate_gradients(Dreg)', None),):
    D = U.quiet(networks.Discriminator, **kw).train()
    D.load_state_dict(D_cpu.state_dict()); D = D.to(dev); D.requires_grad_(True)
    Lg = loss_mod.StyleGAN2Loss(device=dev, G_mapping=None, G_synthesis=None, D=D, r1_gamma=1.0)
    Lg.accumulate_gradients(phase='Dreg', real_img=real.to(dev), real_c=c.to(dev), gen_z=z.to(dev), gen_c=c.to(dev), sync=True, gain=16)
    torch.cuda.synchronize()
    rows = sorted(((rel(p.grad, truth[n]), rel(ref[n], truth[n]), n) for n, p in D.named_parameters() if p.grad is not None and n in truth and n in ref), reverse=True)
    print(f'== {label}')
    for e, er, n in rows[:8]:
        print(f'   {e:.2e} (cpu fp32 reference {er:.2e})  {n}')
# and by hand, with / without the no_weight_gradients() context and with / without the zero-weighted logits term
from torch_utils.ops import conv2d_gradfix
import contextlib
for use_ctx in (False, True):
    for zero_term in (False, True):
        D = U.quiet(networks.Discriminator, **kw).train()
        D.load_state_dict(D_cpu.state_dict()); D = D.to(dev); D.requires_grad_(True)
        img = real.to(dev).requires_grad_(True)
        logits = D(img, c.to(dev))
        with (conv2d_gradfix.no_weight_gradients() if use_ctx else contextlib.nullcontext()):
            g, = torch.autograd.grad(logits.sum(), img, create_graph=True, only_inputs=True)
        pen = g.square().sum([1, 2, 3]) * 0.5
        tot = (logits * 0 + pen) if zero_term else pen
        tot.mean().mul(16).backward()
        torch.cuda.synchronize()
        rows = sorted(((rel(p.grad, truth[n]), n) for n, p in D.named_parameters() if p.grad is not None and n in truth), reverse=True)
        print(f'== by hand: no_weight_gradients={use_ctx} zero-weighted logits term={zero_term}: worst {rows[0][0]:.2e} {rows[0][1]}, {rows[1][0]:.2e} {rows[1][1]}', flush=True)
