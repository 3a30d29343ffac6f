#!/usr/bin/env python
"""Per-call comparison inside the Dreg chain: every march-eligible conv2d call is evaluated by the tile kernel, the marching kernel
and the FFMA kernel and compared with an fp64 convolution of the same operands (max, rms and SIGNED mean error)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import torch.nn.functional as F
import tests.util as U
from oracle import live_ref
from torch_utils import custom_ops
from training import networks, loss as loss_mod
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
plugin = custom_ops.get_plugin('conv2d_plugin')
orig = plugin.conv2d
calls = []


def wrapped(x, w, stride=1, padding=(0, 0), transposed=False, output_padding=(0, 0), flip_w=False, in_scale=None, out_scale=None,
            prec=None, out_hw=None, flop_scale=1.0):
    kwargs = dict(stride=stride, padding=padding, transposed=transposed, output_padding=output_padding, flip_w=flip_w, in_scale=in_scale,
                  out_scale=out_scale, out_hw=out_hw, flop_scale=flop_scale)
    O = w.shape[1] if transposed else w.shape[0]
    if w.shape[2] == 3 and O <= 64 and x.shape[3] >= 64 and stride == 1 and in_scale is None and out_hw is not None:
        custom_ops.set_conv_kernel_family(0); yt = orig(x, w, prec=custom_ops.PREC_TF32X3, **kwargs)
        custom_ops.set_conv_kernel_family(1); ym = orig(x, w, prec=custom_ops.PREC_TF32X3, **kwargs)
        ys = orig(x, w, prec=custom_ops.PREC_FP32_SIMT, **kwargs)
        # fp64 truth of the same op: stride-1 correlation, free output extent == the layer's extent here
        wd = w.double()
        if transposed:
            wd = wd.transpose(0, 1)
            pad = (w.shape[2] - 1 - padding[0], w.shape[3] - 1 - padding[1])
            wd = wd if flip_w else wd.flip([2, 3])
        else:
            pad = padding
            wd = wd.flip([2, 3]) if flip_w else wd
        y64 = F.conv2d(x.double(), wd, padding=pad)
        if y64.shape == yt.shape:
            row = [tuple(x.shape), tuple(w.shape), transposed, flip_w]
            for y in (yt, ym, ys):
                d = (y.double() - y64)
                row += [float(d.abs().max() / y64.abs().max()), float(d.square().mean().sqrt() / y64.square().mean().sqrt()),
                        float(d.mean() / y64.abs().mean()), float((d * y64.sign()).mean() / y64.abs().mean())]
            row.append(float(x.abs().max())); row.append(float((x == 0).float().mean()))
            calls.append(row)
            d = (yt.double() - ym.double()); e_t = yt.double() - y64; e_m = ym.double() - y64
            sc = float(y64.abs().mean())
            print('call', len(calls), 'tile-march: rms', float(d.square().mean().sqrt()) / sc, ' <d,y>/<y,y>', float((d * y64).sum() / (y64 * y64).sum()),
                  ' <e_t,y>/<y,y>', float((e_t * y64).sum() / (y64 * y64).sum()), ' <e_m,y>/<y,y>', float((e_m * y64).sum() / (y64 * y64).sum()))
            for name, e in (('tile', e_t), ('march', e_m)):
                es = (e * y64.sign()) / sc
                byrow = es.reshape(es.shape[0], es.shape[1], -1, 16, es.shape[3]).mean(dim=(0, 1, 2, 4))
                bycol = es.reshape(es.shape[0], es.shape[1], es.shape[2], -1, 16).mean(dim=(0, 1, 2, 3))
                bych = es.mean(dim=(0, 2, 3))
                edge = es[:, :, [0, -1], :].mean(), es[:, :, :, [0, -1]].mean(), es[:, :, 1:-1, 1:-1].mean()
                print(f'   {name}: signed err by row%16 min {float(byrow.min()):+.2e} max {float(byrow.max()):+.2e}; by col%16 min {float(bycol.min()):+.2e} max {float(bycol.max()):+.2e}; '
                      f'by channel min {float(bych.min()):+.2e} max {float(bych.max()):+.2e}; top/bottom rows {float(edge[0]):+.2e} left/right cols {float(edge[1]):+.2e} interior {float(edge[2]):+.2e}')
        return ym
    return orig(x, w, prec=prec, **kwargs)


plugin.conv2d = wrapped
Lg = loss_mod.StyleGAN2Loss(device=dev, G_mapping=None, G_synthesis=None, D=D, r1_gamma=1.0)
Lg.accumulate_gradients(phase='Dreg', real_img=real.to(dev), real_c=c.to(dev), gen_z=z.to(dev), gen_c=c.to(dev), sync=True, gain=16)
torch.cuda.synchronize()
print('columns per kernel (tile | march | ffma): max-rel, rms-rel, signed mean / mean|y|, mean of err*sign(y) / mean|y|')
for r in calls:
    print(r[0], r[1], 'T' if r[2] else '-', 'F' if r[3] else '-', ' | '.join(' '.join(f'{v:+.1e}' for v in r[4 + 4 * i: 8 + 4 * i]) for i in range(3)),
          f' max|x| {r[16]:.2e} zeros {r[17]:.2f}')
