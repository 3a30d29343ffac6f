"""The G+D training iteration the images/sec metric is defined over.

The reference's `training_loop()` (DissimilarDomains/training/training_loop.py:163-666) is a 500-line
driver whose optimiser step / EMA code was de-indented out of its loop in this fork (SURVEY.md section
0.2); the intended behaviour is upstream stylegan2-ada-pytorch.  `TrainingStep` is that inner iteration
and nothing else:

    phases Gmain, Greg (every G_reg_interval), Dmain, Dreg (every D_reg_interval) with the lazy-
    regularisation rescaling of lr and betas                                  (:293-318)
    per phase: zero_grad, requires_grad on the phase's module only, `batch_size // (batch_gpu*num_gpus)`
    accumulation rounds of `loss.accumulate_gradients`, nan_to_num on grads, Adam step  (:459-512)
    G_ema <- lerp(G, G_ema, 0.5 ** (batch_size / ema_nimg)), buffers copied    (:515-523)

Data loading, ADA, snapshots, metrics, logging and the GA hooks are out of scope.  With
torch.distributed initialised, the modules are wrapped in DistributedDataParallel exactly as
:270-285 does (one process per GPU, NCCL all-reduce of the phase's gradients on its last round).
"""
import copy
import numpy as np
import torch

from torch_utils import misc
from torch_utils.ops import conv2d_gradfix
from . import networks
from .loss import StyleGAN2Loss

CONFIGS = {
    # train.py:219-228 (+ --fp32: num_fp16_res=0, conv_clamp=None, :418-423)
    'stylegan2': dict(fmaps=1.0, lrate=0.002, gamma=10.0, ema=10, mbstd=4, map=8, mb=32),   # config-f
    'paper256':  dict(fmaps=0.5, lrate=0.0025, gamma=1.0, ema=20, mbstd=8, map=8, mb=64),
    'paper512':  dict(fmaps=1.0, lrate=0.0025, gamma=0.5, ema=20, mbstd=8, map=8, mb=64),
    'paper1024': dict(fmaps=1.0, lrate=0.002, gamma=2.0, ema=10, mbstd=4, map=8, mb=32),
}


def build_networks(resolution, cfg='stylegan2', z_dim=512, w_dim=512, channel_max=512, device='cuda', **synthesis_extra):
    """G and D as train.py:264-273 builds them for `--cfg=<cfg> --fp32=1`."""
    spec = CONFIGS[cfg]
    channel_base = int(spec['fmaps'] * 32768)
    G = networks.Generator(z_dim=z_dim, c_dim=0, w_dim=w_dim, img_resolution=resolution, img_channels=3,
                           mapping_kwargs=dict(num_layers=spec['map']),
                           synthesis_kwargs=dict(channel_base=channel_base, channel_max=channel_max, num_fp16_res=0,
                                                 conv_clamp=None, **synthesis_extra))
    D = networks.Discriminator(c_dim=0, img_resolution=resolution, img_channels=3, channel_base=channel_base,
                               channel_max=channel_max, num_fp16_res=0, conv_clamp=None,
                               epilogue_kwargs=dict(mbstd_group_size=spec['mbstd']))
    return G.to(device), D.to(device)


class TrainingStep:
    def __init__(self, G, D, batch_size, batch_gpu, device, lrate=0.002, r1_gamma=10.0, ema_kimg=10.0, G_reg_interval=4,
                 D_reg_interval=16, style_mixing_prob=0.9, pl_weight=2.0, rank=0, num_gpus=1, g_trainable=None):
        assert batch_size % (batch_gpu * num_gpus) == 0
        self.device = torch.device(device)
        self.batch_size, self.batch_gpu, self.num_gpus, self.rank = batch_size, batch_gpu, num_gpus, rank
        self.G = G.train().requires_grad_(False).to(self.device)
        self.D = D.train().requires_grad_(False).to(self.device)
        self.G_ema = copy.deepcopy(self.G).eval()
        self.ema_nimg = ema_kimg * 1000
        conv2d_gradfix.enabled = True                          # training_loop.py:209
        torch.backends.cuda.matmul.allow_tf32 = False          # :207-208
        torch.backends.cudnn.allow_tf32 = False
        self.g_trainable = g_trainable                         # optional name filter (Affine+/StyleSpace parts)

        # DDP wrap (:270-285).  G_ema is not wrapped.
        ddp = dict(G_mapping=self.G.mapping, G_synthesis=self.G.synthesis, D=self.D)
        if num_gpus > 1:
            for name, module in list(ddp.items()):
                if len(list(module.parameters())) != 0:
                    module.requires_grad_(True)
                    module = torch.nn.parallel.DistributedDataParallel(
                        module, device_ids=[self.device] if self.device.type == 'cuda' else None, broadcast_buffers=False)
                    module.requires_grad_(False)
                ddp[name] = module
        self.loss = StyleGAN2Loss(device=self.device, **ddp, style_mixing_prob=style_mixing_prob, r1_gamma=r1_gamma,
                                  pl_weight=pl_weight)

        # Phases with lazy regularisation (:293-318).
        self.phases = []
        for name, module, reg_interval in [('G', self.G, G_reg_interval), ('D', self.D, D_reg_interval)]:
            params = [p for n, p in module.named_parameters() if self._trainable(name, n)]
            if reg_interval is None:
                opt = torch.optim.Adam(params, lr=lrate, betas=(0.0, 0.99), eps=1e-8)
                self.phases.append(misc.EasyDict(name=name + 'both', module=module, opt=opt, interval=1, params=params))
            else:
                mb_ratio = reg_interval / (reg_interval + 1)
                opt = torch.optim.Adam(params, lr=lrate * mb_ratio, betas=(float(0 ** mb_ratio), 0.99 ** mb_ratio), eps=1e-8)
                self.phases.append(misc.EasyDict(name=name + 'main', module=module, opt=opt, interval=1, params=params))
                self.phases.append(misc.EasyDict(name=name + 'reg', module=module, opt=opt, interval=reg_interval, params=params))
        self.cur_it = 0

    def _trainable(self, net, pname):
        if net == 'D' or self.g_trainable is None:
            return True
        return any(key in pname for key in self.g_trainable)

    def run(self, real_img, gen_z_all=None):
        """One iteration.  real_img: [batch_size // num_gpus, 3, R, R] float32 in [-1, 1] on the device.
        gen_z_all: optional [len(phases), batch_size // num_gpus, z_dim]; drawn on the device when None (:446-451)."""
        per_gpu = self.batch_size // self.num_gpus
        assert real_img.shape[0] == per_gpu
        real_c = torch.zeros([per_gpu, 0], device=self.device)
        if gen_z_all is None:
            gen_z_all = torch.randn([len(self.phases), per_gpu, self.G.z_dim], device=self.device)
        real_rounds = real_img.split(self.batch_gpu)
        c_rounds = real_c.split(self.batch_gpu)
        for phase, phase_z in zip(self.phases, gen_z_all):
            if self.cur_it % phase.interval != 0:
                continue
            phase.opt.zero_grad(set_to_none=True)
            for p in phase.params:
                p.requires_grad_(True)
            z_rounds = phase_z.split(self.batch_gpu)
            for round_idx, (r_img, r_c, g_z) in enumerate(zip(real_rounds, c_rounds, z_rounds)):
                sync = (round_idx == len(real_rounds) - 1)
                self.loss.accumulate_gradients(phase=phase.name, real_img=r_img, real_c=r_c, gen_z=g_z, gen_c=r_c,
                                               sync=sync, gain=phase.interval)
            for p in phase.params:
                p.requires_grad_(False)
                if p.grad is not None:
                    misc.nan_to_num(p.grad, nan=0, posinf=1e5, neginf=-1e5, out=p.grad)
            phase.opt.step()

        # G_ema (:515-523)
        ema_beta = 0.5 ** (self.batch_size / max(self.ema_nimg, 1e-8))
        with torch.no_grad():
            p_ema, p_cur = list(self.G_ema.parameters()), list(self.G.parameters())
            torch._foreach_lerp_(p_ema, p_cur, 1.0 - ema_beta)       # p_ema = lerp(p, p_ema, beta)
            for b_ema, b in zip(self.G_ema.buffers(), self.G.buffers()):
                b_ema.copy_(b)
        self.cur_it += 1
        return self.loss.last
