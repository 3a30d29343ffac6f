"""StyleGAN2 loss phases over the B200 ops: Gmain, Greg (path length), Dmain, Dreg (R1).

Mirrors DissimilarDomains/training/loss.py:26-152 (`StyleGAN2Loss.run_G`, `run_D`,
`accumulate_gradients(phase, real_img, real_c, gen_z, gen_c, sync, gain)`) without the
`training_stats` reporting; the scalar losses are kept in `self.last` (device tensors, no sync) for
the caller to read.  Greg and Dreg differentiate through first-order gradients, i.e. they exercise
the double-backward closure of conv2d_gradfix / upfirdn2d / bias_act.
"""
import numpy as np
import torch

from torch_utils import misc
from torch_utils.ops import conv2d_gradfix


class Loss:
    def accumulate_gradients(self, phase, real_img, real_c, gen_z, gen_c, sync, gain):  # to be overridden by subclass
        raise NotImplementedError()


class StyleGAN2Loss(Loss):
    def __init__(self, device, G_mapping, G_synthesis, D, augment_pipe=None, style_mixing_prob=0.9, r1_gamma=10,
                 pl_batch_shrink=2, pl_decay=0.01, pl_weight=2):
        super().__init__()
        self.device = device
        self.G_mapping = G_mapping
        self.G_synthesis = G_synthesis
        self.D = D
        self.augment_pipe = augment_pipe
        self.style_mixing_prob = style_mixing_prob
        self.r1_gamma = r1_gamma
        self.pl_batch_shrink = pl_batch_shrink
        self.pl_decay = pl_decay
        self.pl_weight = pl_weight
        self.pl_mean = torch.zeros([], device=device)
        self.last = {}

    def run_G(self, z, c, sync, set_w_requires_grad=False):
        with misc.ddp_sync(self.G_mapping, sync):
            ws = self.G_mapping(z, c)
            if self.style_mixing_prob > 0:
                cutoff = torch.empty([], dtype=torch.int64, device=ws.device).random_(1, ws.shape[1])
                cutoff = torch.where(torch.rand([], device=ws.device) < self.style_mixing_prob, cutoff,
                                     torch.full_like(cutoff, ws.shape[1]))
                ws[:, cutoff:] = self.G_mapping(torch.randn_like(z), c, skip_w_avg_update=True)[:, cutoff:]
        if set_w_requires_grad:
            ws.requires_grad_(True)
        with misc.ddp_sync(self.G_synthesis, sync):
            img = self.G_synthesis(ws)
        return img, ws

    def run_D(self, img, c, sync):
        if self.augment_pipe is not None:
            img = self.augment_pipe(img)
        with misc.ddp_sync(self.D, sync):
            logits = self.D(img, c)
        return logits

    def accumulate_gradients(self, phase, real_img, real_c, gen_z, gen_c, sync, gain):
        assert phase in ['Gmain', 'Greg', 'Gboth', 'Dmain', 'Dreg', 'Dboth']
        do_Gmain = (phase in ['Gmain', 'Gboth'])
        do_Dmain = (phase in ['Dmain', 'Dboth'])
        do_Gpl = (phase in ['Greg', 'Gboth']) and (self.pl_weight != 0)
        do_Dr1 = (phase in ['Dreg', 'Dboth']) and (self.r1_gamma != 0)

        # Gmain: maximize logits for generated images (loss.py:77-86).
        if do_Gmain:
            gen_img, _gen_ws = self.run_G(gen_z, gen_c, sync=(sync and not do_Gpl))
            gen_logits = self.run_D(gen_img, gen_c, sync=False)
            loss_Gmain = torch.nn.functional.softplus(-gen_logits)
            self.last['Loss/G/loss'] = loss_Gmain.detach().mean()
            loss_Gmain.mean().mul(gain).backward()

        # Gpl: path length regularization (loss.py:89-111).
        if do_Gpl:
            batch_size = gen_z.shape[0] // self.pl_batch_shrink
            gen_img, gen_ws = self.run_G(gen_z[:batch_size], gen_c[:batch_size], sync=sync, set_w_requires_grad=True)
            pl_noise = torch.randn_like(gen_img) / np.sqrt(gen_img.shape[2] * gen_img.shape[3])
            with conv2d_gradfix.no_weight_gradients():
                pl_grads = torch.autograd.grad(outputs=[(gen_img * pl_noise).sum()], inputs=[gen_ws], create_graph=True,
                                               only_inputs=True, allow_unused=True)[0]
            pl_lengths = pl_grads.square().sum(2).mean(1).sqrt()
            pl_mean = self.pl_mean.lerp(pl_lengths.mean(), self.pl_decay)
            self.pl_mean.copy_(pl_mean.detach())
            pl_penalty = (pl_lengths - pl_mean).square()
            loss_Gpl = pl_penalty * self.pl_weight
            self.last['Loss/G/reg'] = loss_Gpl.detach().mean()
            (gen_img[:, 0, 0, 0] * 0 + loss_Gpl).mean().mul(gain).backward()

        # Dmain: minimize logits for generated images (loss.py:114-123).
        loss_Dgen = 0
        if do_Dmain:
            gen_img, _gen_ws = self.run_G(gen_z, gen_c, sync=False)
            gen_logits = self.run_D(gen_img, gen_c, sync=False)
            loss_Dgen = torch.nn.functional.softplus(gen_logits)
            loss_Dgen.mean().mul(gain).backward()

        # Dmain: maximize logits for real images.  Dr1: R1 regularization (loss.py:127-152).
        if do_Dmain or do_Dr1:
            real_img_tmp = real_img.detach().requires_grad_(do_Dr1)
            real_logits = self.run_D(real_img_tmp, real_c, sync=sync)
            loss_Dreal = 0
            if do_Dmain:
                loss_Dreal = torch.nn.functional.softplus(-real_logits)
                self.last['Loss/D/loss'] = (loss_Dgen + loss_Dreal).detach().mean()
            loss_Dr1 = 0
            if do_Dr1:
                with conv2d_gradfix.no_weight_gradients():
                    r1_grads = torch.autograd.grad(outputs=[real_logits.sum()], inputs=[real_img_tmp], create_graph=True,
                                                   only_inputs=True)[0]
                r1_penalty = r1_grads.square().sum([1, 2, 3])
                loss_Dr1 = r1_penalty * (self.r1_gamma / 2)
                self.last['Loss/D/reg'] = loss_Dr1.detach().mean()
            (real_logits * 0 + loss_Dreal + loss_Dr1).mean().mul(gain).backward()
